#!/usr/bin/env python
"""L2 eviction policy of the linear-sweep kernel's tile stores (variant bits 16-18 of the `_ex` hooks: 0 = no hint,
1 = evict_first, 2 = evict_last, 3 / 4 = evict_first for the distance / the mask tile only, 7 = none; +n = without the pacing
defaults, variant bit 26) for distances + byte mask, distances + fp32 mask and the fused kernel, by shape.

    python tools/l2_hint_probe.py > gpurun_out/l2_hint_probe.json
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
A = 15
rows = []
NOPACE = 1 << 26
VARIANTS = (("hint0", 7 << 16), ("hint1", 1 << 16), ("hint2", 2 << 16), ("hint3", 3 << 16), ("hint4", 4 << 16),
            ("hint5", 5 << 16), ("hint6", 6 << 16), ("hint0n", (7 << 16) | NOPACE), ("hint1n", (1 << 16) | NOPACE), ("default", 0))
import os  # noqa: E402

SWEEP_SHAPES = () if os.environ.get("PROBE_PART") == "others" else ((32, 512), (64, 256), (28, 384), (70, 229), (16, 511))
for B, L in SWEEP_SHAPES:
    xyz = (10 * torch.randn(B, L, A, 3, device="cuda", generator=g)).contiguous()
    mask = torch.rand(B, L, A, device="cuda", generator=g) < 0.5
    xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan"))).contiguous()  # missing atoms are NaN
    maskf = mask.float().contiguous()
    dist = torch.empty(B, L, L, A, A, device="cuda")
    dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device="cuda")
    om, th, ph = (torch.empty(B, L, L, device="cuda") for _ in range(3))
    pairs = B * L * L
    for kind in ("dist+bool", "fused", "dist+f32"):
        dmf = torch.empty(B, L, L, A, A, device="cuda") if kind == "dist+f32" else None
        nbytes = pairs * A * A * (8 if kind == "dist+f32" else 5) + (pairs * 12 if kind == "fused" else 0)
        rec = {"B": B, "L": L, "kind": kind, "GB": nbytes / 1e9}
        for rep in range(2):
            for name, variant in VARIANTS:

                def run():
                    if kind == "fused":
                        rc = lib.ps_inter_residue_geometry_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(),
                                                              dm.data_ptr(), om.data_ptr(), th.data_ptr(), ph.data_ptr(),
                                                              B, L, A, variant, s)
                    elif kind == "dist+bool":
                        rc = lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dm.data_ptr(),
                                                      B, L, A, variant, s)
                    else:
                        rc = lib.ps_pair_dist_mask_ex(xyz.data_ptr(), maskf.data_ptr(), 1, dist.data_ptr(), dmf.data_ptr(),
                                                      B, L, A, variant, s)
                    _cabi.check(rc, kind)

                best, med = time_call(run, iters=12, warmup=3)
                rec.setdefault(f"{name}_gbs", []).append(round(nbytes / best / 1e6))
        rows.append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
        del dmf
    del dist, dm, om, th, ph

# ---- the other tile kernels: any-A (`pair_cols_kernel`, hint in variant bits 28-29) and the column-strip kernel of the
# staged atom counts 5 / 10 / 14 (`pair_tiles_kernel`, bits 16-17), distances + byte mask and distances only
others = []
for B, L, A in ((24, 256, 25), (12, 256, 37), (40, 256, 20), (256, 512, 5), (64, 384, 10), (32, 384, 14)):
    xyz = (10 * torch.randn(B, L, A, 3, device="cuda", generator=g)).contiguous()
    mask = torch.rand(B, L, A, device="cuda", generator=g) < 0.5
    dist = torch.empty(B, L, L, A, A, device="cuda")
    dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device="cuda")
    shift = 16 if A in (5, 10, 14) else 28
    om, th, ph = (torch.empty(B, L, L, device="cuda") for _ in range(3))
    for kind in ("dist+bool", "dist") + (("fused",) if shift == 16 else ()):
        nbytes = B * L * L * (A * A * (4 if kind == "dist" else 5) + (12 if kind == "fused" else 0))
        rec = {"B": B, "L": L, "A": A, "kind": kind, "GB": nbytes / 1e9}
        for rep in range(2):
            for hint in (3, 1, 2, 0):  # 3 = explicitly none, 0 = the launcher's default
                def run():
                    if kind == "fused":
                        _cabi.check(lib.ps_inter_residue_geometry_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(),
                                                                     dm.data_ptr(), om.data_ptr(), th.data_ptr(),
                                                                     ph.data_ptr(), B, L, A, hint << shift, s), kind)
                        return
                    with_mask = kind == "dist+bool"
                    rc = lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr() if with_mask else None, 0, dist.data_ptr(),
                                                  dm.data_ptr() if with_mask else None, B, L, A, hint << shift, s)
                    _cabi.check(rc, kind)

                best, med = time_call(run, iters=12, warmup=3)
                rec.setdefault({3: "hint0", 0: "default"}.get(hint, f"hint{hint}") + "_gbs", []).append(round(nbytes / best / 1e6))
        others.append(rec)
        print(json.dumps(rec), file=sys.stderr, flush=True)
    del dist, dm
print(json.dumps({"sweep_kernel": rows, "other_tile_kernels": others}, indent=1))
