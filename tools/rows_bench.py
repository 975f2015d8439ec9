#!/usr/bin/env python
"""Any-shape distance kernel (pair_rows_kernel) timings on shapes the staged kernel does not cover.

    python tools/rows_bench.py > gpurun_out/rows_bench.json
"""
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tools"))

from kernel_bench import entry, peak_gbs, time_call  # noqa: E402
from protstruc_b200 import _cabi  # noqa: E402

DEV = "cuda"


def main():
    lib = _cabi.load()
    peak = peak_gbs()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=DEV).manual_seed(0)
    out = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak, "results": []}
    F, R = 1 << 8, 1 << 12  # keep the staged kernel out / row kernel only
    shapes = [(16, 512, 15, F), (16, 512, 15, F | R), (64, 128, 25, 0), (64, 128, 25, R), (32, 128, 37, 0),
              (64, 256, 4, 0), (64, 256, 4, F), (256, 256, 3, 0), (256, 256, 3, F), (1024, 24, 15, 0), (4, 1024, 27, 0), (2, 300, 15, F),
              (64, 256, 8, 0), (32, 256, 20, 0)]
    for B, L, A, variant in shapes:
        xyz = (10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)).contiguous()
        mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
        dist = torch.empty(B, L, L, A, A, device=DEV)
        dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)

        def run(dp=dist.data_ptr(), mp=dmask.data_ptr(), am=mask.data_ptr()):
            _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), am, 0, dp, mp, B, L, A, variant, s), "k1")

        n = B * L * L * A * A
        tag = " [rows]" if variant & R else (" [any-A tiles]" if variant & F else "")
        best, med = time_call(run)
        out["results"].append(entry(f"dist+bool mask B{B} L{L} A{A}" + tag, best, med, n * 5, peak))
        best, med = time_call(lambda: run(mp=0, am=0))
        out["results"].append(entry(f"dist only B{B} L{L} A{A}" + tag, best, med, n * 4, peak))
        best, med = time_call(lambda: run(dp=0))
        out["results"].append(entry(f"bool mask only B{B} L{L} A{A}" + tag, best, med, n, peak))
        del dist, dmask
    print(json.dumps(out, indent=1))
    for r in out["results"]:
        print(f'{r["kernel"]:48s} {r["best_ms"]:8.3f} ms {r["GBps_best"]:8.0f} GB/s', file=sys.stderr)


if __name__ == "__main__":
    main()
