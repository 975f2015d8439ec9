#!/usr/bin/env python
"""Tile-size / CTA-size sweep of the any-A tile kernel (pair_cols_kernel).

    python tools/cols_sweep.py > gpurun_out/cols_sweep.json
"""
import json
import math
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tools"))

from kernel_bench import peak_gbs, time_call  # noqa: E402
from protstruc_b200 import _cabi  # noqa: E402

DEV = "cuda"


def main():
    lib = _cabi.load()
    peak = peak_gbs()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=DEV).manual_seed(0)
    out = {"hbm_peak_gbs": peak, "results": []}
    F = 1 << 8
    for B, L, A in [(16, 512, 15), (24, 256, 25), (12, 256, 37), (40, 256, 20), (48, 256, 14), (64, 256, 8), (20, 256, 29)]:
        xyz = (10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)).contiguous()
        mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
        dist = torch.empty(B, L, L, A, A, device=DEV)
        dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)
        n = B * L * L * A * A
        AA = A * A
        for kind, bpe in (("dist+mask", 5), ("dist", 4)):
            quantum = 4 // math.gcd(AA, 4)
            rows = []
            seen = set()
            for kb in (0, 6, 8, 12, 16, 24, 32, 48, 64):
                for warps in ((0,) if kb == 0 else (2, 3, 4, 5, 6, 7, 8, 10, 12, 16)):
                    units = 0 if kb == 0 else max(1, min(255, round(kb * 1024 / (AA * bpe) / quantum)))
                    if (units, warps) in seen:
                        continue
                    seen.add((units, warps))
                    variant = F | (units << 16) | (warps << 24)

                    def run():
                        am, mp = (mask.data_ptr(), dmask.data_ptr()) if kind == "dist+mask" else (0, 0)
                        _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), am, 0, dist.data_ptr(), mp, B, L, A, variant, s), "k1")

                    try:
                        best, med = time_call(run, iters=5, warmup=2)
                    except Exception as e:  # noqa: BLE001
                        rows.append({"kb": kb, "error": str(e)[:80]})
                        continue
                    gbs = n * bpe / (best / 1e3) / 1e9
                    plan = _cabi.last_pair_dist_plan()
                    rows.append({"kb": kb, "pairs": plan["tile_pairs"], "threads": 32 * warps, "best_ms": best, "GBps": gbs})
                    print(f"B{B} L{L} A{A} {kind:9s} tile~{kb:3d}KB pairs={rows[-1]['pairs']:4d} thr={rows[-1]['threads']:3d} "
                          f"{best:8.3f} ms {gbs:7.0f} GB/s", file=sys.stderr)
            out["results"].append({"shape": [B, L, A], "kind": kind, "rows": rows})
        del dist, dmask
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
