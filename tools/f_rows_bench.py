#!/usr/bin/env python
"""Timings of the neighbouring rows (rigid-frame family, Kabsch alignment, top-k mask) and of the small
per-residue kernels, through the public API, at B = 256, L = 512, A = 15 (CUDA events, inputs resident).

    python tools/f_rows_bench.py > gpurun_out/f_rows_bench.json
"""
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tools"))

import protstruc_b200 as ps  # noqa: E402
from kernel_bench import entry, peak_gbs, time_call  # noqa: E402

DEV = "cuda"


def main():
    peak = peak_gbs()
    B, L, A = 256, 512, 15
    g = torch.Generator(device=DEV).manual_seed(0)
    xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
    mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.7
    mask[:, :, :4] = True
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    tgt = ps.StructureBatch.from_xyz(xyz + 0.1 * torch.randn_like(xyz), mask)
    state = B * L * A * 12
    out = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak, "shape": [B, L, A], "results": []}

    def add(name, fn, nbytes):
        best, med = time_call(fn, iters=10, warmup=3)
        out["results"].append(entry(name, best, med, nbytes, peak))
        print(f"{name:70s} {best * 1e3:9.1f} us {nbytes / best / 1e6:8.0f} GB/s", file=sys.stderr)

    rot = torch.linalg.qr(torch.randn(B, 3, 3, device=DEV, generator=g))[0]
    frames = sb.backbone_orientations()
    trans = sb.backbone_translations().contiguous()
    add("K3 backbone_dihedrals (+mask)", sb.backbone_dihedrals, B * L * (A * 12 + 12 + 3 + 5))
    add("K3 backbone_orientations", sb.backbone_orientations, B * L * (36 + 36))
    add("K4 center_of_mass", sb.center_of_mass, B * L * 12)
    add("K4 center_at (COM + translate, in place)", lambda: sb.center_at(torch.zeros(B, 3, device=DEV)), B * L * 12 + 2 * state)
    add("f1 get_local_xyz", sb.get_local_xyz, 2 * state + B * L * 36)
    add("f1 rotate (B,3,3)", lambda: sb.rotate(rot), 2 * state)
    add("f1 translate (B,1,3), in place", lambda: sb.translate(torch.ones(B, 1, 3, device=DEV)), 2 * state)
    add("f1 from_backbone_orientations_translations",
        lambda: ps.StructureBatch.from_backbone_orientations_translations(frames, trans), B * L * (48 + A * 16))
    add("f2 align (Kabsch + rotate + translate)", lambda: sb.align(tgt), 2 * state + B * L * A + 4 * state)

    def standardize_roundtrip():
        sb.standardize()
        sb.unstandardize()

    add("K4 standardize + unstandardize", standardize_roundtrip, 5 * state + B * L * A)
    one = ps.StructureBatch.from_xyz(xyz[:1], mask[:1])
    query = xyz[0, :16, 1].contiguous()
    add("f4 get_topk_nearest_residue_mask (B=1, 16 query points, k=128)",
        lambda: one.get_topk_nearest_residue_mask(query, k=128), L * A * 12 + L)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
