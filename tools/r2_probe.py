#!/usr/bin/env python
"""Round-2 kernel probe: the kernels VERDICT r1 ranked furthest below their roof, old variant beside new.

    python tools/r2_probe.py > gpurun_out/r2_probe.json

CUDA events on the launching stream, inputs resident, 3 warm-ups, best and median of 10.  Between timed launches a
256 MB buffer is written so that the small kernels (whose state would otherwise sit in the 126 MB L2) read from HBM.
"""
import json
import statistics
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import protstruc_b200 as ps  # noqa: E402
from protstruc_b200 import _cabi  # noqa: E402

DEV = "cuda"


def peak_gbs():
    p = REPO / "MEASURED_PEAKS.json"
    return float(json.loads(p.read_text())["hbm_gbs"]) if p.exists() else 6650.0


FLUSH = None


def time_call(fn, iters=10, warmup=3, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        if flush:
            FLUSH.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return min(times), statistics.median(times)


def main():
    lib = _cabi.load()
    peak = peak_gbs()
    s = torch.cuda.current_stream().cuda_stream
    out = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak, "results": []}
    g = torch.Generator(device=DEV).manual_seed(0)

    def add(name, best, med, nbytes, **extra):
        gbs = nbytes / (best / 1e3) / 1e9
        d = {"kernel": name, "best_ms": best, "median_ms": med, "algorithmic_bytes": nbytes, "GBps_best": gbs,
             "frac_of_measured_hbm": gbs / peak}
        d.update(extra)
        out["results"].append(d)

    def inputs(B, L, A, nan_masked=True, ragged=False):
        xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
        mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
        if nan_masked:
            xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
        if ragged:  # last L - len_b residues zero-padded, len_b ~ U[0.75 L, L]
            lens = (L * (0.75 + 0.25 * torch.rand(B, device=DEV, generator=g))).long()
            pad = torch.arange(L, device=DEV)[None, :] >= lens[:, None]
            xyz = torch.where(pad[:, :, None, None], torch.zeros_like(xyz), xyz)
            mask = mask & ~pad[:, :, None]
        return xyz.contiguous(), mask.contiguous()

    # ---- K2f: BASELINE config 3 (256 x 512 backbone), packed kernel vs the exact-sequence kernel of round 1
    for (B, L, A, nan_masked, ragged, tag) in ((256, 512, 5, False, False, "config3 (all atoms valid)"),
                                                (256, 512, 5, "gly", False, "config3 shape, 8 % of the residues without CB (glycine)"),
                                                (256, 512, 5, False, True, "config3 shape, ragged (lengths U[0.75 L, L], zero padding)"),
                                                (256, 512, 5, True, True, "config3 shape, NaN-masked + ragged"),
                                                (64, 384, 15, True, True, "B64 L384 A15 NaN-masked + ragged"),
                                                (64, 511, 15, False, False, "B64 L511 (odd) A15")):
        if nan_masked == "gly":
            xyz, _ = inputs(B, L, A, False, False)
            gly = torch.rand(B, L, device=DEV, generator=g) < 0.08
            xyz[:, :, 4][gly] = float("nan")
        else:
            xyz, _ = inputs(B, L, A, nan_masked, ragged)
        om = torch.empty(B, L, L, device=DEV)
        th, ph = torch.empty_like(om), torch.empty_like(om)
        for variant, label in ((0, "packed FP32, 3 CTAs / SM (default)"), (4, "packed FP32, 4 CTAs / SM"), (3, "packed FP32, two rows per iteration, 2 CTAs / SM"),
                               (1, "exact sequence (round 1)")):
            def run(variant=variant):
                _cabi.check(lib.ps_trrosetta_angles_ex(xyz.data_ptr(), B, L, A, 0, om.data_ptr(), th.data_ptr(),
                                                       ph.data_ptr(), variant, s), "k2f")
            best, med = time_call(run)
            add(f"K2f omega+theta+phi {tag} [{label}]", best, med, B * (L * L * 12 + L * 5 * 12),
                pairs_per_s=B * L * L / (best / 1e3))
        if tag.startswith("config3 (all"):
            one = torch.empty(B, L, L, device=DEV)
            for (si, sj, kind, what) in (([1, 4], [1, 4], 0, "dihedral CA,CB | CA,CB"), ([0, 1, 4], [4], 0, "dihedral N,CA,CB | CB"),
                                         ([1, 4], [4], 1, "planar CA,CB | CB")):
                for variant, label in ((0, "packed (default)"), (1, "exact sequence (round 1)")):
                    def run_generic(si=si, sj=sj, kind=kind, variant=variant):
                        _cabi.check(lib.ps_pair_angles_ex(xyz.data_ptr(), B, L, A, _cabi.int_array(si), len(si), _cabi.int_array(sj),
                                                          len(sj), kind, one.data_ptr(), variant, s), "k2 generic")
                    best, med = time_call(run_generic)
                    add(f"K2 generic {what} {tag} [{label}]", best, med, B * (L * L * 4 + L * 5 * 12))
            del one
        if A >= 3:
            def run_virtual():
                _cabi.check(lib.ps_trrosetta_angles_ex(xyz.data_ptr(), B, L, A, 1, om.data_ptr(), th.data_ptr(),
                                                       ph.data_ptr(), 0, s), "k2f virtual")
            best, med = time_call(run_virtual)
            add(f"K2f omega+theta+phi {tag} [packed, virtual CB]", best, med, B * (L * L * 12 + L * 5 * 12))
        del om, th, ph

    # ---- any-A tile kernel (run-time atom count): the reference's own tests use A = 25, atom37 users A = 37
    for (B, L, A) in ((24, 256, 25), (12, 256, 37), (40, 256, 20), (16, 512, 15)):
        xyz, mask = inputs(B, L, A)
        dist = torch.empty(B, L, L, A, A, device=DEV)
        dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)
        force = (1 << 8) if A == 15 else 0

        def run_cols():
            _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(), B, L, A,
                                                 force, s), "cols")
        best, med = time_call(run_cols, flush=False)
        add(f"K1 any-A tile kernel dist+boolmask B{B} L{L} A{A}", best, med, B * L * L * A * A * 5,
            plan=_cabi.last_pair_dist_plan())

        def run_cols_r1():
            _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(), B, L, A,
                                                 force | (13 << 24), s), "cols generic A")
        best, med = time_call(run_cols_r1, flush=False)
        add(f"K1 any-A tile kernel dist+boolmask B{B} L{L} A{A} [run-time-A instantiation]", best, med, B * L * L * A * A * 5)

        def run_cols_dist():
            _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), None, 0, dist.data_ptr(), None, B, L, A, force, s), "cols d")
        best, med = time_call(run_cols_dist, flush=False)
        add(f"K1 any-A tile kernel dist only B{B} L{L} A{A}", best, med, B * L * L * A * A * 4)
        del dist, dmask

    # ---- K4 standardize: register-resident single-read kernel vs the three-pass kernel
    for (B, L, A, tag) in ((1024, 128, 15, "config4 B1024 L128"), (256, 512, 15, "B256 L512"), (16, 512, 15, "B16 L512"),
                           (2, 4096, 15, "B2 L4096"), (64, 229, 15, "B64 L229 (odd sizes)"), (4096, 64, 15, "B4096 L64")):
        xyz, mask = inputs(B, L, A)
        mu, sd = torch.empty(B, 3, device=DEV), torch.empty(B, 3, device=DEV)
        xo = torch.empty_like(xyz)
        for variant, label in ((0, "register-resident (default)"), (3, "register-resident, <= 2 quads per thread"),
                               (2, "register-resident, scalar mapping"), (1, "three-pass (round 1)")):
            def run(variant=variant):
                _cabi.check(lib.ps_masked_stats_ex(xyz.data_ptr(), mask.data_ptr(), 0, B, L, A, mu.data_ptr(), sd.data_ptr(),
                                                   xo.data_ptr(), variant, s), "k4")
            best, med = time_call(run)
            add(f"K4 standardize {tag} A{A} [{label}]", best, med, B * L * A * 25)
        maskf = mask.float()

        def run_f32():
            _cabi.check(lib.ps_masked_stats_ex(xyz.data_ptr(), maskf.data_ptr(), 1, B, L, A, mu.data_ptr(), sd.data_ptr(),
                                               xo.data_ptr(), 0, s), "k4f")
        best, med = time_call(run_f32)
        add(f"K4 standardize {tag} A{A} [register-resident, fp32 mask]", best, med, B * L * A * 28)

    # ---- K4 centre of mass, f2 Kabsch, K3 backbone at B256 x L512 x A15 (and config 4's shape)
    for (B, L, A) in ((256, 512, 15), (1024, 128, 15)):
        xyz, mask = inputs(B, L, A, nan_masked=False)
        com = torch.empty(B, 3, device=DEV)

        def run_com():
            _cabi.check(lib.ps_center_of_mass(xyz.data_ptr(), B, L, A, 1, com.data_ptr(), s), "com")
        best, med = time_call(run_com)
        add(f"K4 center_of_mass B{B} L{L} A{A}", best, med, B * L * 12, sector_bytes=B * L * 32)
        tgt = xyz + 1.0
        m8 = mask.reshape(B, L * A).to(torch.uint8).contiguous()
        rot, tr = torch.empty(B, 3, 3, device=DEV), torch.empty(B, 3, device=DEV)

        def run_kabsch():
            _cabi.check(lib.ps_kabsch(xyz.data_ptr(), tgt.data_ptr(), m8.data_ptr(), B, B, L * A, rot.data_ptr(),
                                      tr.data_ptr(), s), "kabsch")
        best, med = time_call(run_kabsch)
        add(f"f2 kabsch B{B} L{L} A{A}", best, med, B * L * A * 25)

        def run_kabsch_small():  # 64 atoms per structure: the launch and the serial 3x3 solve, next to no data
            _cabi.check(lib.ps_kabsch(xyz.data_ptr(), tgt.data_ptr(), m8.data_ptr(), B, B, 64, rot.data_ptr(),
                                      tr.data_ptr(), s), "kabsch small")
        best, med = time_call(run_kabsch_small)
        add(f"f2 kabsch B{B} x 64 atoms (solve only)", best, med, B * 64 * 25)
        sb = ps.StructureBatch.from_xyz(xyz, mask)
        rm = sb.residue_mask.to(torch.uint8).contiguous()
        ch = sb.chain_idx.float().contiguous()
        dih = torch.empty(B, L, 3, device=DEV)
        dm = torch.empty(B, L, 3, dtype=torch.bool, device=DEV)
        fr = torch.empty(B, L, 3, 3, device=DEV)

        def run_backbone():
            _cabi.check(lib.ps_backbone(xyz.data_ptr(), rm.data_ptr(), ch.data_ptr(), B, L, A, 0, 1, 2, dih.data_ptr(),
                                        dm.data_ptr(), fr.data_ptr(), s), "k3")
        best, med = time_call(run_backbone)
        add(f"K3 backbone dihedrals+mask+frames B{B} L{L} A{A}", best, med, B * L * (41 + 51))

    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
