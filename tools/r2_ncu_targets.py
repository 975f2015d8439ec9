#!/usr/bin/env python
"""Launches each round-2 target kernel a few times at its BASELINE shape — the command profiled with
`ncu --set full -k regex:'trrosetta_fast|masked_stats|kabsch' -c 12` (profiles/r2*_ncu_summary.txt)."""
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


def inputs(B, L, A, nan_masked):
    xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
    mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
    if nan_masked:
        xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
    return xyz.contiguous(), mask.contiguous()


# K2f packed, BASELINE config 3 (all atoms valid), then NaN-masked
for nan_masked in (False, True):
    B, L, A = 256, 512, 5
    xyz, _ = inputs(B, L, A, nan_masked)
    om, th, ph = (torch.empty(B, L, L, device=DEV) for _ in range(3))
    for _ in range(2):
        _cabi.check(lib.ps_trrosetta_angles_ex(xyz.data_ptr(), B, L, A, 0, om.data_ptr(), th.data_ptr(), ph.data_ptr(), 0, s), "k2f")
    torch.cuda.synchronize()
    del om, th, ph
# K4 register-resident and three-pass at config 4 and at 256 x 512
for (B, L) in ((1024, 128), (256, 512)):
    xyz, mask = inputs(B, L, 15, True)
    mu, sd, xo = torch.empty(B, 3, device=DEV), torch.empty(B, 3, device=DEV), torch.empty_like(xyz)
    for variant in (0, 1):
        for _ in range(2):
            _cabi.check(lib.ps_masked_stats_ex(xyz.data_ptr(), mask.data_ptr(), 0, B, L, 15, mu.data_ptr(), sd.data_ptr(),
                                               xo.data_ptr(), variant, s), "k4")
    torch.cuda.synchronize()
# Kabsch at 256 x 512 x 15
B, L = 256, 512
xyz, mask = inputs(B, L, 15, False)
tgt = xyz + 1.0
m8 = mask.reshape(B, L * 15).to(torch.uint8).contiguous()
rot, tr = torch.empty(B, 3, 3, device=DEV), torch.empty(B, 3, device=DEV)
for _ in range(2):
    _cabi.check(lib.ps_kabsch(xyz.data_ptr(), tgt.data_ptr(), m8.data_ptr(), B, B, L * 15, rot.data_ptr(), tr.data_ptr(), s), "kabsch")
torch.cuda.synchronize()
print("ok")
