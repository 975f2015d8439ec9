#!/usr/bin/env python
"""What a pass over config 4's coordinates costs at best: torch's own copy / add of the same 23.6 MB beside K4 and the
elementwise kernels, as `ncu --metrics gpu__time_duration.sum` targets (event timing of ONE small launch measures the
host's launch latency, ~10-15 us, not the kernel).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/floor.csv python tools/floor_probe.py
"""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from protstruc_b200 import _cabi  # noqa: E402

DEV = "cuda"
lib = _cabi.load()
B, L, A = 1024, 128, 15
x = torch.randn(B, L, A, 3, device=DEV)
out = torch.empty_like(x)
mask = torch.rand(B, L, A, device=DEV) < 0.5
mu, sd = torch.empty(B, 3, device=DEV), torch.empty(B, 3, device=DEV)
s = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    out.copy_(x)
    torch.add(x, 1.0, out=out)
    torch.mul(x, 2.0, out=out)
    _cabi.check(lib.ps_masked_stats(x.data_ptr(), mask.data_ptr(), 0, B, L, A, mu.data_ptr(), sd.data_ptr(), out.data_ptr(), s), "k4")
    _cabi.check(lib.ps_masked_stats(x.data_ptr(), mask.data_ptr(), 0, B, L, A, mu.data_ptr(), sd.data_ptr(), None, s), "k4 stats")
    _cabi.check(lib.ps_scale_shift(out.data_ptr(), sd.data_ptr(), mu.data_ptr(), B, L, A, out.data_ptr(), s), "scale_shift")
torch.cuda.synchronize()
print("ok")
