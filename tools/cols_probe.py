#!/usr/bin/env python
"""One shape of the any-A tile kernel, a few launches (target for `ncu -k regex:pair_cols`).

    python tools/cols_probe.py [B L A]
"""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
from protstruc_b200 import _cabi  # noqa: E402

B, L, A = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (64, 128, 25)
lib = _cabi.load()
g = torch.Generator(device="cuda").manual_seed(0)
xyz = (10.0 * torch.randn(B, L, A, 3, device="cuda", generator=g)).contiguous()
mask = torch.rand(B, L, A, device="cuda", generator=g) < 0.5
dist = torch.empty(B, L, L, A, A, device="cuda")
dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for _ in range(4):
    _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(),
                                         B, L, A, 1 << 8, s), "k1")
torch.cuda.synchronize()
print("ok", float(dist[0, 0, 1, 0, 0]))
