#!/usr/bin/env python
"""Fused K1 (inter_residue_geometry, L = 512, A = 15) by input kind: does the angle part leave the HBM-bound regime on
everyday inputs (zero-padded ragged batches, missing atoms)?

    python tools/fused_input_kinds_probe.py > gpurun_out/fused_input_kinds.json
"""
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

from protstruc_b200 import _cabi  # noqa: E402

DEV = "cuda"


def main():
    lib = _cabi.load()
    s = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device=DEV).manual_seed(0)
    B, L, A = 16, 512, 15
    d = torch.empty(B, L, L, A, A, device=DEV)
    dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)
    ang = torch.empty(3, B, L, L, device=DEV)
    out = []
    for nan_masked in (True, False):
        for ragged in (False, True):
            xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
            mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
            if nan_masked:
                xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
            if ragged:
                lens = (L * (0.75 + 0.25 * torch.rand(B, device=DEV, generator=g))).long()
                pad = torch.arange(L, device=DEV)[None, :] >= lens[:, None]
                xyz = torch.where(pad[:, :, None, None], torch.zeros_like(xyz), xyz)
                mask = mask & ~pad[:, :, None]
            xyz, mask = xyz.contiguous(), mask.contiguous()

            def call():
                _cabi.check(lib.ps_inter_residue_geometry(xyz.data_ptr(), mask.data_ptr(), 0, d.data_ptr(), dm.data_ptr(),
                                                          ang[0].data_ptr(), ang[1].data_ptr(), ang[2].data_ptr(), B, L, A, s), "fused")
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(8):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                call()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            nbytes = B * (L * L * (225 * 5 + 12) + L * 15 * 13)
            out.append({"missing_atoms_are_nan": nan_masked, "ragged_zero_padded": ragged, "best_ms": best,
                        "GBps": nbytes / (best / 1e3) / 1e9})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
