#!/usr/bin/env python
"""K1 at A = 15: the linear-sweep kernel (pair_sweep.cu) against the column-strip kernel (pair_dist.cu).

    timeout 600 python tools/sweep_check.py [--no-timing] > gpurun_out/sweep_check.json

1. bit-equality of every output byte (distances, bool / fp32 mask, omega / theta / phi, compact planes) between the two
   kernels on ragged, NaN-masked batches of many lengths (tile wraps inside a structure, across structures, partial last
   tile, the manual-staging tiles at the end of the arrays), with guard bands around every output;
2. achieved GB/s of both kernels by structure length, distances + mask and fused (CUDA events, best of 5).
"""
import argparse
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from protstruc_b200 import _cabi  # noqa: E402
from tests import helpers as H  # noqa: E402

DEV = "cuda"
STRIP = 1 << 15
SWEEP = 1 << 27
GUARD = 64


def guarded(shape, dtype, fill):
    n = 1
    for d in shape:
        n *= d
    flat = torch.full((n + 2 * GUARD,), fill, dtype=dtype, device=DEV)
    return flat, flat[GUARD:GUARD + n].view(shape)


def run(lib, x, m, code, variant, fused, compact):
    B, L, A = x.shape[:3]
    s = torch.cuda.current_stream().cuda_stream
    mdt = torch.bool if code == 0 else torch.float32
    gd, d = guarded((B, L, L, A, A), torch.float32, -7.0)
    gm, dm = guarded((B, L, L, A, A), mdt, True if code == 0 else -7.0)
    outs = {"dist": (gd, d), "mask": (gm, dm)}
    if not fused:
        _cabi.check(lib.ps_pair_dist_mask_ex(x.data_ptr(), m.data_ptr(), code, d.data_ptr(), dm.data_ptr(), B, L, A, variant, s),
                    "ps_pair_dist_mask_ex")
    else:
        ga, ang = guarded((3, B, L, L), torch.float32, -7.0)
        outs["angles"] = (ga, ang)
        _cabi.check(lib.ps_inter_residue_geometry_ex(x.data_ptr(), m.data_ptr(), code, d.data_ptr(), dm.data_ptr(),
                                                     ang[0].data_ptr(), ang[1].data_ptr(), ang[2].data_ptr(), B, L, A,
                                                     variant, s), "ps_inter_residue_geometry_ex")
    torch.cuda.synchronize()
    plan = _cabi.last_pair_dist_plan()
    return outs, plan


def same_bits(a, b):
    return torch.equal(a.contiguous().view(torch.uint8), b.contiguous().view(torch.uint8))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-timing", action="store_true")
    args = ap.parse_args()
    lib = _cabi.load()
    report = {"equality": [], "timing": [], "ok": True}
    shapes = [(2, 32), (1, 33), (3, 37), (2, 45), (1, 229), (7, 40), (2, 250), (2, 256), (1, 511), (2, 512), (1, 1000), (5, 64),
              (4, 100), (3, 127), (1, 2048)]
    for idx, (B, L) in enumerate(shapes):
        for kind in ("bool", "float"):
            if kind == "float" and L > 600:
                continue
            xyz, mask, _ = H.synthetic_batch(7000 + idx, B, L, 15, kind)
            x, m = xyz.to(DEV), mask.to(DEV).contiguous()
            code = 0 if kind == "bool" else 1
            for fused in (False, True):
                a, plan_a = run(lib, x, m, code, STRIP, fused, False)
                b, plan_b = run(lib, x, m, code, SWEEP, fused, False)
                entry = {"B": B, "L": L, "mask": kind, "fused": fused, "sweep_plan": plan_b["sweep"], "strip_plan": plan_a["sweep"],
                         "launches": [plan_a["launches"], plan_b["launches"]]}
                ok = plan_b["sweep"] == 1 and plan_a["sweep"] == 0
                for name in a:
                    ga, ta = a[name]
                    gb, tb = b[name]
                    eq = same_bits(ta, tb)
                    fill_ok = same_bits(ga[:GUARD], gb[:GUARD]) and same_bits(ga[-GUARD:], gb[-GUARD:]) and \
                        bool((gb[:GUARD] == gb[0]).all()) and bool((gb[-GUARD:] == gb[-1]).all())
                    entry[name] = eq
                    entry[name + "_guards"] = fill_ok
                    ok = ok and eq and fill_ok
                entry["ok"] = ok
                report["ok"] = report["ok"] and ok
                report["equality"].append(entry)
                del a, b
    if not args.no_timing:
        s = torch.cuda.current_stream().cuda_stream
        g = torch.Generator(device=DEV).manual_seed(5)
        for L in (229, 250, 300, 384, 509, 511, 512, 1024):
            B = max(2, int(3.2e9 / (L * L * 1137)))
            xyz = 10.0 * torch.randn(B, L, 15, 3, device=DEV, generator=g)
            mask = torch.rand(B, L, 15, device=DEV, generator=g) < 0.5
            xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan"))).contiguous()
            d = torch.empty(B, L, L, 15, 15, device=DEV)
            dm = torch.empty(B, L, L, 15, 15, dtype=torch.bool, device=DEV)
            ang = torch.empty(3, B, L, L, device=DEV)
            for fused in (False, True):
                nbytes = B * (L * L * (225 * 5 + (12 if fused else 0)) + L * 15 * 13)
                row = {"L": L, "B": B, "fused": fused}
                for label, variant in (("strip_auto", STRIP), ("sweep", SWEEP)):
                    def call():
                        if fused:
                            _cabi.check(lib.ps_inter_residue_geometry_ex(xyz.data_ptr(), mask.data_ptr(), 0, d.data_ptr(), dm.data_ptr(),
                                                                         ang[0].data_ptr(), ang[1].data_ptr(), ang[2].data_ptr(), B, L,
                                                                         15, variant, s), "fused")
                        else:
                            _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, d.data_ptr(), dm.data_ptr(), B, L, 15,
                                                                 variant, s), "dist")
                    for _ in range(3):
                        call()
                    torch.cuda.synchronize()
                    best = 1e9
                    for _ in range(5):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        call()
                        e1.record()
                        torch.cuda.synchronize()
                        best = min(best, e0.elapsed_time(e1))
                    row[label + "_gbs"] = nbytes / (best / 1e3) / 1e9
                report["timing"].append(row)
            # fp32 mask: one launch (sweep) vs two (strip)
            if L in (256, 384, 512):
                mf = mask.float()
                dmf = torch.empty(B, L, L, 15, 15, device=DEV)
                row = {"L": L, "B": B, "fused": False, "mask": "fp32"}
                for label, variant in (("strip_auto", STRIP), ("sweep", SWEEP)):
                    def callf():
                        _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mf.data_ptr(), 1, d.data_ptr(), dmf.data_ptr(), B, L, 15,
                                                             variant, s), "dist f32 mask")
                    for _ in range(2):
                        callf()
                    torch.cuda.synchronize()
                    best = 1e9
                    for _ in range(5):
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        callf()
                        e1.record()
                        torch.cuda.synchronize()
                        best = min(best, e0.elapsed_time(e1))
                    row[label + "_gbs"] = B * L * L * 225 * 8 / (best / 1e3) / 1e9
                report["timing"].append(row)
                del dmf, mf
            del d, dm, ang
    print(json.dumps(report, indent=1))
    if not report["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
