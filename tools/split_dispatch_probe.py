#!/usr/bin/env python
"""`inter_residue_geometry` on the staged atom counts below 15: ONE fused launch (variant bit 19) against distance tiles
+ the exact-sequence angle kernel (bit 20), by input kind.  The launcher's default splits 5 and 10 atoms.

    python tools/split_dispatch_probe.py > gpurun_out/split_dispatch_probe.json
"""
import json
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
rows = []
for B, L, A in ((256, 512, 5), (64, 384, 10), (32, 384, 14)):
    for kind in ("every atom present", "missing atoms as NaN (half of the slots)", "8 % of the residues without CB"):
        xyz = (10 * torch.randn(B, L, A, 3, device="cuda", generator=g)).contiguous()
        if kind.startswith("missing"):
            mask = torch.rand(B, L, A, device="cuda", generator=g) < 0.5
        elif kind.startswith("8 %"):
            mask = torch.ones(B, L, A, dtype=torch.bool, device="cuda")
            mask[:, :, 4] = torch.rand(B, L, device="cuda", generator=g) >= 0.08
        else:
            mask = torch.ones(B, L, A, dtype=torch.bool, device="cuda")
        xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan"))).contiguous()
        dist = torch.empty(B, L, L, A, A, device="cuda")
        dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device="cuda")
        om, th, ph = (torch.empty(B, L, L, device="cuda") for _ in range(3))
        nbytes = B * L * L * (A * A * 5 + 12)
        rec = {"B": B, "L": L, "A": A, "input": kind, "GB": nbytes / 1e9}
        for rep in range(2):
            for name, variant in (("fused", 1 << 19), ("split", 1 << 20), ("default", 0)):
                def run():
                    _cabi.check(lib.ps_inter_residue_geometry_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(),
                                                                 dm.data_ptr(), om.data_ptr(), th.data_ptr(), ph.data_ptr(),
                                                                 B, L, A, variant, s), name)
                best, med = time_call(run, iters=10, warmup=3)
                rec.setdefault(name + "_ms", []).append(round(best, 4))
        rows.append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
        del dist, dm, om, th, ph
print(json.dumps(rows, indent=1))
