#!/usr/bin/env python
"""K4 `ps_masked_stats` over structure counts / sizes (C-ABI call, CUDA events): few large structures take the
thread-block-cluster path, many small ones the plain kernel.  Checked against an fp64 evaluation on the device.

    python tools/stats_cluster_bench.py
"""
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tools"))
import torch
import protstruc_b200 as ps
from protstruc_b200 import _cabi
from kernel_bench import time_call
lib = _cabi.load()
g = torch.Generator(device='cuda').manual_seed(0)
for B, L in ((1, 512), (4, 512), (16, 512), (64, 512), (256, 512), (1024, 128), (2, 4096)):
    A = 15
    xyz = (10 * torch.randn(B, L, A, 3, device='cuda', generator=g)).contiguous()
    mask = torch.rand(B, L, A, device='cuda', generator=g) < 0.6
    mu = torch.empty(B, 3, device='cuda'); sd = torch.empty(B, 3, device='cuda'); out = torch.empty_like(xyz)
    s = torch.cuda.current_stream().cuda_stream
    def run():
        _cabi.check(lib.ps_masked_stats(xyz.data_ptr(), mask.data_ptr(), 0, B, L, A, mu.data_ptr(), sd.data_ptr(), out.data_ptr(), s), 'k4')
    best, med = time_call(run, iters=20, warmup=3)
    # reference on device in fp64
    m = mask[..., None].double(); x = torch.nan_to_num(xyz.double())
    cnt = mask.sum((1, 2)).double()[:, None]
    mu_ref = (x * m).sum((1, 2)) / cnt
    sd_ref = (((x - mu_ref[:, None, None]) ** 2 * m).sum((1, 2)) / cnt).sqrt()
    print(f'B={B:5d} L={L:5d}: {best*1e3:7.1f} us  mu err {float((mu.double()-mu_ref).abs().max()):.2e}  sd err {float((sd.double()-sd_ref).abs().max()):.2e}  out err {float((out.double() - (xyz.double()-mu_ref[:,None,None])/sd_ref[:,None,None]).abs().max()):.2e}')
