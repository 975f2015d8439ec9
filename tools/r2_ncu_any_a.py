#!/usr/bin/env python
"""Launches the any-A tile kernel (pair_cols_kernel) at A = 25 and A = 37, distances + bool mask and distances only —
the command profiled with `ncu --set full -k regex:pair_cols -c 8` (profiles/r2*_any_a_ncu_summary.txt)."""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from protstruc_b200 import _cabi  # noqa: E402

DEV = "cuda"
lib = _cabi.load()
s = torch.cuda.current_stream().cuda_stream
g = torch.Generator(device=DEV).manual_seed(0)
for (B, L, A) in ((24, 256, 25), (12, 256, 37)):
    xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
    mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
    xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan"))).contiguous()
    dist = torch.empty(B, L, L, A, A, device=DEV)
    dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)
    for with_mask in (True, False):
        for _ in range(2):
            _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr() if with_mask else None, 0, dist.data_ptr(),
                                                 dmask.data_ptr() if with_mask else None, B, L, A, 0, s), "any-A")
        torch.cuda.synchronize()
    del dist, dmask
print("ok")
