"""Small driver for profiling the config-3 angle kernel under ncu: python tools/c3_probe.py"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from protstruc_b200 import _cabi

lib = _cabi.load()
B, L, A = 256, 512, 5
g = torch.Generator(device="cuda").manual_seed(3)
xyz = 10.0 * torch.randn(B, L, A, 3, device="cuda", generator=g)
om = torch.empty(B, L, L, device="cuda")
th, ph = torch.empty_like(om), torch.empty_like(om)
s = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    _cabi.check(lib.ps_trrosetta_angles(xyz.data_ptr(), B, L, A, 0, om.data_ptr(), th.data_ptr(), ph.data_ptr(), s), "c3")
torch.cuda.synchronize()
print("ok")
