#!/usr/bin/env python
"""Staged K1 (distances + mask, and fused) with the automatic schedule choice, the cell schedule (variant bit 13) and
the lock-step schedule (variant bit 11), by L.

    python tools/lockstep_probe.py
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
import os
SHAPES = ((256, 128), (100, 190), (70, 229), (60, 250), (64, 256), (40, 300), (30, 350), (28, 384), (20, 437),
          (16, 500), (16, 510), (16, 512), (4, 1024))
if os.environ.get("PROBE_SHAPES"):
    SHAPES = tuple(tuple(int(v) for v in item.split("x")) for item in os.environ["PROBE_SHAPES"].split(","))
NAN_MASKED = os.environ.get("PROBE_NAN", "0") == "1"
for B, L in SHAPES:
    A = 15
    xyz = (10 * torch.randn(B, L, A, 3, device="cuda", generator=g)).contiguous()
    mask = torch.rand(B, L, A, device="cuda", generator=g) < 0.5
    if NAN_MASKED:  # as the PDB ingest yields: missing atoms are NaN
        xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan"))).contiguous()
    dist = torch.empty(B, L, L, A, A, device="cuda")
    dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device="cuda")
    nbytes = B * L * L * A * A * 5
    row = []
    for rep in range(2):
        for variant in (0, 1 << 13, 1 << 11, 1 << 14):
            def run():
                _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dm.data_ptr(),
                                                     B, L, A, variant, s), "k1")
            best, med = time_call(run, iters=10, warmup=3)
            row.append(f"{ {0: 'auto', 1 << 13: 'cells', 1 << 11: 'lockstep', 1 << 14: 'relaxed'}[variant] } {nbytes / best / 1e6:6.0f}")
    print(f"B={B:4d} L={L:5d} ({nbytes / 1e9:.2f} GB): " + " | ".join(row) + " GB/s")
    om, th, ph = (torch.empty(B, L, L, device="cuda") for _ in range(3))
    fbytes = nbytes + B * L * L * 12
    row = []
    for rep in range(2):
        for variant in (0, 1 << 13, 1 << 11, 1 << 14):
            def run_fused():
                _cabi.check(lib.ps_inter_residue_geometry_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dm.data_ptr(),
                                                             om.data_ptr(), th.data_ptr(), ph.data_ptr(), B, L, A, variant, s), "k1f")
            best, med = time_call(run_fused, iters=10, warmup=3)
            row.append(f"{ {0: 'auto', 1 << 13: 'cells', 1 << 11: 'lockstep', 1 << 14: 'relaxed'}[variant] } {fbytes / best / 1e6:6.0f}")
    print(f"      fused            : " + " | ".join(row) + " GB/s")
    del dist, dm, om, th, ph
