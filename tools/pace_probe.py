#!/usr/bin/env python
"""How the achieved bandwidth of the linear-sweep K1 kernel depends on the pace at which tiles are produced.

    python tools/pace_probe.py > gpurun_out/pace_probe.json

Per (length, output kind): GB/s with the issuing lane sleeping 0 .. 700 ns after handing a tile to the TMA engine
(variant bits 28-30), with 3 or 4 tile buffers per SM, and with the dearer MUFU square root — the knobs that change how
fast a tile buffer comes back with its next tile.  CUDA events, best of 5 after 3 warm-ups.
"""
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from protstruc_b200 import _cabi  # noqa: E402

DEV = "cuda"
SWEEP = 1 << 27


def main():
    lib = _cabi.load()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=DEV).manual_seed(11)
    out = []
    for L in (512, 384, 229):
        B = max(2, int(4.5e9 / (L * L * 1137)))
        xyz = 10.0 * torch.randn(B, L, 15, 3, device=DEV, generator=g)
        mask = torch.rand(B, L, 15, device=DEV, generator=g) < 0.5
        xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan"))).contiguous()
        maskf = mask.float()
        d = torch.empty(B, L, L, 15, 15, device=DEV)
        dm = torch.empty(B, L, L, 15, 15, dtype=torch.bool, device=DEV)
        dmf = torch.empty(B, L, L, 15, 15, device=DEV)
        ang = torch.empty(3, B, L, L, device=DEV)
        kinds = {
            "dist+boolmask": (lambda v: lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, d.data_ptr(), dm.data_ptr(), B, L, 15, v, s),
                              B * L * L * 225 * 5),
            "dist only": (lambda v: lib.ps_pair_dist_mask_ex(xyz.data_ptr(), None, 0, d.data_ptr(), None, B, L, 15, v, s), B * L * L * 225 * 4),
            "fused": (lambda v: lib.ps_inter_residue_geometry_ex(xyz.data_ptr(), mask.data_ptr(), 0, d.data_ptr(), dm.data_ptr(), ang[0].data_ptr(),
                                                                 ang[1].data_ptr(), ang[2].data_ptr(), B, L, 15, v, s), B * L * L * (225 * 5 + 12)),
            "dist+f32mask": (lambda v: lib.ps_pair_dist_mask_ex(xyz.data_ptr(), maskf.data_ptr(), 1, d.data_ptr(), dmf.data_ptr(), B, L, 15, v, s),
                             B * L * L * 225 * 8),
        }
        for kind, (call, nbytes) in kinds.items():
            row = {"L": L, "B": B, "kind": kind}
            configs = [(f"pace{p}00ns", SWEEP | (p << 28)) for p in range(8)]
            configs += [("3 buffers", SWEEP | (3 << 4)), ("3 buffers pace200ns", SWEEP | (3 << 4) | (2 << 28)),
                        ("sqrt.approx (non-ftz)", SWEEP | 1), ("sqrt.rn", SWEEP | 2)]
            if kind == "dist+f32mask":
                configs = [c for c in configs if "sqrt" not in c[0]] + [("2 buffers", SWEEP | (2 << 4))]
            for label, variant in configs:
                for _ in range(3):
                    _cabi.check(call(variant), kind)
                torch.cuda.synchronize()
                best = 1e9
                for _ in range(5):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    _cabi.check(call(variant), kind)
                    e1.record()
                    torch.cuda.synchronize()
                    best = min(best, e0.elapsed_time(e1))
                row[label] = round(nbytes / (best / 1e3) / 1e9, 1)
            out.append(row)
        del d, dm, dmf, ang
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
