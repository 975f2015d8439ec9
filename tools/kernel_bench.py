#!/usr/bin/env python
"""Per-kernel timings on the BASELINE.json configs (run on the GPU box; CUDA events, inputs resident).

    python tools/kernel_bench.py [--quick] > gpurun_out/kernel_bench.json

Reports, per kernel: best / median time, algorithmic bytes, achieved GB/s and the fraction of
the measured HBM peak.  Also sweeps the tuning variants of K1 (sqrt flavour, warps per CTA).
"""
import argparse
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


def time_call(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    return min(times), statistics.median(times)


def entry(name, best, med, nbytes, peak, **extra):
    gbs = nbytes / (best / 1e3) / 1e9
    d = {"kernel": name, "best_ms": best, "median_ms": med, "algorithmic_bytes": nbytes, "GBps_best": gbs,
         "frac_of_measured_hbm": gbs / peak}
    d.update(extra)
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    lib = _cabi.load()
    peak = peak_gbs()
    s = torch.cuda.current_stream().cuda_stream
    out = {"device": torch.cuda.get_device_name(0), "hbm_peak_gbs": peak, "results": []}
    g = torch.Generator(device=DEV).manual_seed(0)

    def inputs(B, L, A, nan_masked=True):
        """xyz ~ 10 N(0,1), Bernoulli(0.5) mask, masked slots NaN (as the PDB ingest yields)."""
        xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
        mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
        if nan_masked:
            xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
        return xyz.contiguous(), mask

    # ---- K1 at the metric shape (L=512, A=15), variants
    B, L, A = 16, 512, 15
    xyz, mask = inputs(B, L, A)
    dist = torch.empty(B, L, L, A, A, device=DEV)
    dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)
    om = torch.empty(B, L, L, device=DEV)
    th, ph = torch.empty_like(om), torch.empty_like(om)
    nbytes = B * (L * L * A * A * 5 + L * A * 13)
    variants = {"sqrt.approx.ftz (default)": 0, "sqrt.approx": 1, "sqrt.rn": 2, "any-A tile kernel": 1 << 8, "row kernel": (1 << 8) | (1 << 12)}
    variants["1 warp per tile (6 warps/SM)"] = 1 << 9
    variants["DIAGNOSTIC stores only (no arithmetic)"] = 1 << 10
    variants["DIAGNOSTIC stores only, 6 tile buffers"] = (1 << 10) | (6 << 4)
    if not args.quick:
        for w in (3, 5, 6):
            variants[f"default, {w} tile buffers/CTA"] = w << 4
    for label, v in variants.items():
        def run(v=v):
            _cabi.check(lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(),
                                                 B, L, A, v, s), "k1")
        best, med = time_call(run)
        out["results"].append(entry(f"K1 dist+boolmask B{B} L{L} A{A} [{label}]", best, med, nbytes, peak))

    def run_dist_only():
        _cabi.check(lib.ps_pair_dist_mask(xyz.data_ptr(), None, 0, dist.data_ptr(), None, B, L, A, s), "k1d")
    best, med = time_call(run_dist_only)
    out["results"].append(entry(f"K1 dist only B{B} L{L} A{A}", best, med, B * L * L * A * A * 4, peak))

    maskf = mask.float()
    dmaskf = torch.empty(B, L, L, A, A, device=DEV)

    def run_f32mask():
        _cabi.check(lib.ps_pair_dist_mask(xyz.data_ptr(), maskf.data_ptr(), 1, dist.data_ptr(), dmaskf.data_ptr(),
                                          B, L, A, s), "k1f")
    best, med = time_call(run_f32mask)
    out["results"].append(entry(f"K1 dist + fp32 mask (one launch since round 2) B{B} L{L} A{A}", best, med, B * L * L * A * A * 8, peak))
    del dmaskf, maskf

    def run_fused():  # noqa: E306
        _cabi.check(lib.ps_inter_residue_geometry(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(),
                                                  om.data_ptr(), th.data_ptr(), ph.data_ptr(), B, L, A, s), "fused")
    best, med = time_call(run_fused)
    out["results"].append(entry(f"K1+K2f fused inter_residue_geometry B{B} L{L} A{A}", best, med,
                                B * (L * L * (A * A * 5 + 12) + L * A * 13), peak,
                                structures_per_s=B / (best / 1e3)))

    def run_fill():
        dist.zero_()
        dmask.zero_()
    best, med = time_call(run_fill)
    out["results"].append(entry("reference point: torch fill of the same dist+mask buffers (library kernel)", best, med,
                                B * L * L * A * A * 5, peak))
    del dist, dmask, om, th, ph

    # ---- config 2: 64 x 256 x 15 dist + mask
    B, L, A = 64, 256, 15
    xyz, mask = inputs(B, L, A)
    dist = torch.empty(B, L, L, A, A, device=DEV)
    dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)

    def run_c2():
        _cabi.check(lib.ps_pair_dist_mask(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(), B, L, A, s), "c2")
    best, med = time_call(run_c2)
    out["results"].append(entry("config2: K1 dist+boolmask B64 L256 A15", best, med, B * (L * L * A * A * 5 + L * A * 13), peak,
                                structures_per_s=B / (best / 1e3)))
    del dist, dmask

    # ---- staged kernel at other atom counts (backbone + CB, atom14)
    for (B, L, A) in ((128, 512, 5), (16, 512, 14)):
        xyz, mask = inputs(B, L, A)
        dist = torch.empty(B, L, L, A, A, device=DEV)
        dmask = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)

        def run_other(xyz=xyz, mask=mask, dist=dist, dmask=dmask, B=B, L=L, A=A):
            _cabi.check(lib.ps_pair_dist_mask(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(), B, L, A, s), "kA")
        best, med = time_call(run_other)
        out["results"].append(entry(f"K1 dist+boolmask B{B} L{L} A{A} (staged kernel, other atom count)", best, med,
                                    B * (L * L * A * A * 5 + L * A * 13), peak))
        del dist, dmask

    # ---- config 3: 256 x 512 backbone (A=5), omega/theta/phi
    B, L, A = 256, 512, 5
    xyz, _ = inputs(B, L, A, nan_masked=False)  # config 3: backbone slots, all valid
    om = torch.empty(B, L, L, device=DEV)
    th, ph = torch.empty_like(om), torch.empty_like(om)

    def run_c3():
        _cabi.check(lib.ps_trrosetta_angles(xyz.data_ptr(), B, L, A, 0, om.data_ptr(), th.data_ptr(), ph.data_ptr(), s), "c3")
    best, med = time_call(run_c3)
    out["results"].append(entry("config3: K2f omega+theta+phi B256 L512 A5", best, med, B * (L * L * 12 + L * A * 12), peak,
                                structures_per_s=B / (best / 1e3), pairs_per_s=B * L * L / (best / 1e3),
                                binding_roof="FP32/SFU issue (not HBM)"))
    si, sj = _cabi.int_array([1, 4]), _cabi.int_array([1, 4])

    def run_c3_generic():
        _cabi.check(lib.ps_pair_angles(xyz.data_ptr(), B, L, A, si, 2, sj, 2, 0, om.data_ptr(), s), "c3g")
    best, med = time_call(run_c3_generic)
    out["results"].append(entry("config3: K2 generic dihedral (omega) B256 L512 A5", best, med, B * (L * L * 4 + L * A * 12), peak))
    del om, th, ph

    # ---- K3 backbone at B256 x L512 x A15
    B, L, A = 256, 512, 15
    xyz, mask = inputs(B, L, A)
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    best, med = time_call(lambda: sb.backbone_features())
    out["results"].append(entry("K3 backbone dihedrals+mask+frames B256 L512 A15 (façade call)", best, med, B * L * (41 + 51), peak))
    best, med = time_call(lambda: sb.center_of_mass())
    out["results"].append(entry("K4 center_of_mass B256 L512 A15 (façade call)", best, med, B * L * 12 + B * 12, peak))

    # ---- config 4: standardize + T=300 diffusion on 1024 x 128 x 15
    B, L, A, T = 1024, 128, 15, 300
    xyz, mask = inputs(B, L, A)
    mu = torch.empty(B, 3, device=DEV)
    sd = torch.empty(B, 3, device=DEV)
    xo = torch.empty_like(xyz)

    def run_std():
        _cabi.check(lib.ps_masked_stats(xyz.data_ptr(), mask.data_ptr(), 0, B, L, A, mu.data_ptr(), sd.data_ptr(), xo.data_ptr(), s), "k4")
    best, med = time_call(run_std)
    out["results"].append(entry("config4: K4 standardize B1024 L128 A15", best, med, B * L * A * (12 + 1 + 12), peak))
    t = torch.arange(T + 1, device=DEV)
    f_t = torch.cos((t / T + 8e-3) / (1 + 8e-3) * torch.pi / 2).square()
    ab = f_t / f_t[0]
    beta = torch.cat([torch.zeros(1, device=DEV), torch.clip(1 - ab[1:] / ab[:-1], min=1e-5, max=0.999)])[:T]
    betas = beta[:, None].repeat(1, B).contiguous()
    per_b = L * A * 3

    def run_single_steps():
        src, dst = xo, xyz
        for k in range(T):
            _cabi.check(lib.ps_diffuse(src.data_ptr(), betas[k].data_ptr(), None, 7, k, 0, dst.data_ptr(), B, per_b, s), "k5")
            src, dst = dst, src
    best, med = time_call(run_single_steps, iters=3, warmup=1)
    out["results"].append(entry("config4: K5 300 single-step launches B1024 L128 A15", best, med, T * 2 * B * per_b * 4, peak,
                                trajectories_per_s=B / (best / 1e3), us_per_step=1e3 * best / T))

    def run_fused_steps():
        _cabi.check(lib.ps_diffuse_steps(xo.data_ptr(), betas.data_ptr(), T, 7, 0, 0, xyz.data_ptr(), B, per_b, s), "k5m")
    best, med = time_call(run_fused_steps, iters=3, warmup=1)
    out["results"].append(entry("config4: K5m 300 steps fused in one launch B1024 L128 A15", best, med, 2 * B * per_b * 4, peak,
                                trajectories_per_s=B / (best / 1e3), normals_per_s=T * B * per_b / (best / 1e3),
                                binding_roof="ALU (Philox + Box-Muller)"))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
