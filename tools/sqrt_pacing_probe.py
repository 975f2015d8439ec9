#!/usr/bin/env python
"""Staged K1 at A = 15 with the flush-to-zero MUFU square root (variant 0) and the non-ftz flavour (variant 1, three
more issue slots per element), by kernel kind and structure length.  The slower flavour is FASTER for the kernels
without the angle triple: the memory system takes the tiles better when they are produced a little more slowly.

    python tools/sqrt_pacing_probe.py
"""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tools"))
from kernel_bench import time_call  # noqa: E402
from protstruc_b200 import _cabi  # noqa: E402

lib = _cabi.load()
g = torch.Generator(device="cuda").manual_seed(0)
s = torch.cuda.current_stream().cuda_stream
A = 15
for B, L in ((256, 128), (70, 229), (64, 256), (28, 384), (16, 500), (16, 512), (4, 1024)):
    xyz = (10 * torch.randn(B, L, A, 3, device="cuda", generator=g)).contiguous()
    mask = torch.rand(B, L, A, device="cuda", generator=g) < 0.5
    xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan"))).contiguous()
    dist = torch.empty(B, L, L, A, A, device="cuda")
    dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device="cuda")
    om, th, ph = (torch.empty(B, L, L, device="cuda") for _ in range(3))
    n = B * L * L * A * A
    kinds = {
        "dist+mask": (5 * n, lambda v: lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dm.data_ptr(), B, L, A, v, s)),
        "dist only": (4 * n, lambda v: lib.ps_pair_dist_mask_ex(xyz.data_ptr(), 0, 0, dist.data_ptr(), 0, B, L, A, v, s)),
        "fused": (5 * n + 12 * B * L * L, lambda v: lib.ps_inter_residue_geometry_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dm.data_ptr(), om.data_ptr(), th.data_ptr(), ph.data_ptr(), B, L, A, v, s)),
    }
    for name, (nbytes, call) in kinds.items():
        row = []
        for variant in (0, 1, 0, 1):
            best, med = time_call(lambda: _cabi.check(call(variant), name), iters=10, warmup=3)
            row.append(f"{'ftz' if variant == 0 else 'non-ftz'} {nbytes / best / 1e6:6.0f}")
        print(f"B={B:4d} L={L:5d} {name:10s}: " + " | ".join(row) + " GB/s")
    del dist, dm
