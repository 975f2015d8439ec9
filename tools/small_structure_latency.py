#!/usr/bin/env python
"""Kernel latency of the fused feature set for ONE small structure (C-ABI call, CUDA events, best of 50).

    python tools/small_structure_latency.py
"""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from protstruc_b200 import _cabi  # noqa: E402

lib = _cabi.load()
g = torch.Generator(device="cuda").manual_seed(0)
for L in (32, 48, 64, 96, 128, 229, 384, 512):
    B, A = 1, 15
    xyz = (10 * torch.randn(B, L, A, 3, device="cuda", generator=g)).contiguous()
    mask = torch.rand(B, L, A, device="cuda", generator=g) < 0.5
    dist = torch.empty(B, L, L, A, A, device="cuda")
    dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device="cuda")
    om, th, ph = (torch.empty(B, L, L, device="cuda") for _ in range(3))

    def run():
        _cabi.check(lib.ps_inter_residue_geometry(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dm.data_ptr(),
                                                  om.data_ptr(), th.data_ptr(), ph.data_ptr(), B, L, A,
                                                  torch.cuda.current_stream().cuda_stream), "k1")

    graph = torch.cuda.CUDAGraph()
    run()
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        for _ in range(20):
            run()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 20)
    nbytes = L * L * (A * A * 5 + 12)
    print(f"L={L:4d}: {best * 1e3:7.2f} us per launch  {nbytes / best / 1e6:7.0f} GB/s")
