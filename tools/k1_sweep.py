#!/usr/bin/env python
"""Sweeps the K1 tuning space (warps per tile x tile buffers per CTA x fused angles) over several shapes.

    python tools/k1_sweep.py > gpurun_out/k1_sweep.json
"""
import json
import statistics
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from protstruc_b200 import _cabi  # noqa: E402

DEV = "cuda"


def time_call(fn, iters=12, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), statistics.median(ts)


def main():
    lib = _cabi.load()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=DEV).manual_seed(0)
    results = []
    for (B, L) in ((16, 512), (28, 384), (64, 256), (80, 229), (5, 1000)):
        A = 15
        xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
        mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
        xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan"))).contiguous()
        dist = torch.empty(B, L, L, A, A, device=DEV)
        dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)
        om = torch.empty(B, L, L, device=DEV)
        th, ph = torch.empty_like(om), torch.empty_like(om)
        for fused in (0, 1):
            nbytes = B * (L * L * (A * A * 5 + (12 if fused else 0)) + L * A * 13)
            for wpt_bit in (0, 1):
                for slots in (0, 3, 5, 6):
                    variant = (wpt_bit << 9) | (slots << 4)

                    def run():
                        if fused:
                            rc = lib.ps_inter_residue_geometry_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(),
                                                                  dmask.data_ptr(), om.data_ptr(), th.data_ptr(), ph.data_ptr(),
                                                                  B, L, A, variant, s)
                        else:
                            rc = lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(),
                                                          B, L, A, variant, s)
                        _cabi.check(rc, "k1")
                    best, med = time_call(run)
                    results.append({"B": B, "L": L, "fused": fused, "non_default_wpt": wpt_bit, "slots": slots or "max",
                                    "best_ms": best, "median_ms": med, "GBps_best": nbytes / best / 1e6,
                                    "GBps_median": nbytes / med / 1e6})
        del dist, dmask, om, th, ph
    print(json.dumps(results, indent=1))


if __name__ == "__main__":
    main()
