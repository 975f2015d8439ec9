#!/usr/bin/env python
"""Measures the deviation of the CUDA path from the reference golden vectors (run on the GPU box).

    python tests/parity_report.py > gpurun_out/parity_report.json

For every golden case: max |dist - ref| (abs and rel), max circular deviation of omega/theta/phi and of
the backbone dihedrals over ALL finite entries and over the well-conditioned ones (min sin >= 0.1),
frames, statistics, and the bit-exactness flags of masks / NaN placement / diffusion.
"""
import json
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import protstruc_b200 as ps  # noqa: E402
from tests import helpers as H  # noqa: E402
from tests.test_gpu_parity import angle_conditioning, make_batch  # noqa: E402

CASES = ["synthetic_small", "synthetic_floatmask_oddL", "synthetic_A5", "synthetic_A25_like_reference_tests",
         "synthetic_ragged_33", "real_1a6v_HL"]


def dev_stats(actual, expected, cond=None, circular=True):
    a, e = actual.cpu(), expected.cpu()
    nan_equal = bool(torch.equal(torch.isnan(a), torch.isnan(e)))
    fin = ~torch.isnan(e) & ~torch.isnan(a)
    d = H.circular_diff(a, e) if circular else (a.double() - e.double()).abs()
    out = {"nan_placement_equal": nan_equal, "max_all_finite": float(d[fin].max()) if fin.any() else 0.0}
    if cond is not None:
        good = fin & (cond.reshape(e.shape) >= H.SIN_GATE)
        out["max_well_conditioned"] = float(d[good].max()) if good.any() else 0.0
        out["n_well_conditioned"] = int(good.sum())
    out["n_finite"] = int(fin.sum())
    return out


def main():
    report = {"device": torch.cuda.get_device_name(0), "cases": {}}
    for name in CASES:
        g = H.load_golden(name)
        sb = make_batch(g)
        xyz = H.t(g["xyz"])
        A = xyz.shape[2]
        r = {}
        dist, dist_mask = sb.pairwise_distance_matrix()
        if name == "real_1a6v_HL":
            pairs = [(dist[:, :40, :40], H.t(g["ref_dist_crop"])), (dist[:, :, :, 1, 1], H.t(g["ref_d_ca"])),
                     (dist[:, :, :, 4, 4], H.t(g["ref_d_cb"])), (dist[:, :, :, 0, 3], H.t(g["ref_d_no"]))]
            r["dist_mask_bit_exact"] = bool(torch.equal(dist_mask[:, :40, :40].cpu(), H.t(g["ref_dist_mask_crop"])))
        else:
            pairs = [(dist, H.t(g["ref_dist"]))]
            r["dist_mask_bit_exact"] = bool(torch.equal(dist_mask.cpu(), H.t(g["ref_dist_mask"])))
        abs_err = rel_err = 0.0
        nan_ok = True
        for a, e in pairs:
            a = a.cpu()
            nan_ok &= bool(torch.equal(torch.isnan(a), torch.isnan(e)))
            fin = ~torch.isnan(e)
            d = (a[fin].double() - e[fin].double()).abs()
            if d.numel():
                abs_err = max(abs_err, float(d.max()))
                rel_err = max(rel_err, float((d / e[fin].double().clamp_min(1e-30))[e[fin] > 0].max()))
        r["dist"] = {"max_abs": abs_err, "max_rel": rel_err, "nan_placement_equal": nan_ok}
        if A >= 5:
            out = sb.inter_residue_geometry()
            for which in ("omega", "theta", "phi"):
                r[which] = dev_stats(out[which], H.t(g[f"ref_{which}"]), angle_conditioning(xyz, which),
                                     circular=which != "phi")
            # the standalone packed angle kernel (K2f, `trrosetta_angles`): spends the 1e-5 rad contract
            om, th, ph = sb.trrosetta_angles()
            for which, got in (("omega", om), ("theta", th), ("phi", ph)):
                r[f"{which}_packed_kernel"] = dev_stats(got, H.t(g[f"ref_{which}"]), angle_conditioning(xyz, which),
                                                        circular=which != "phi")
        dih, dmask = sb.backbone_dihedrals()
        r["bb_dihedrals"] = dev_stats(dih, H.t(g["ref_bb_dihedrals"]))
        r["bb_dihedral_mask_bit_exact"] = bool(torch.equal(dmask.cpu(), H.t(g["ref_bb_dihedral_mask"])))
        r["frames"] = dev_stats(sb.backbone_orientations(), H.t(g["ref_frames"]), circular=False)
        r["com"] = dev_stats(sb.center_of_mass(), H.t(g["ref_com"]), circular=False)
        sb2 = make_batch(g)
        sb2.standardize()
        r["mu"] = dev_stats(sb2.mu, H.t(g["ref_mu"]), circular=False)
        r["sd"] = dev_stats(sb2.std, H.t(g["ref_sd"]), circular=False)
        r["std_xyz"] = dev_stats(sb2.get_xyz(), H.t(g["ref_std_xyz"]), circular=False)
        sb3 = make_batch(g)
        sb3.diffuse_xyz(H.t(g["ref_beta"]), noise=H.t(g["ref_noise"]))
        ref = H.t(g["ref_diffused"])
        r["diffuse_bit_exact"] = bool(torch.equal(torch.nan_to_num(sb3.get_xyz().cpu()), torch.nan_to_num(ref))
                                      and torch.equal(torch.isnan(sb3.get_xyz().cpu()), torch.isnan(ref)))
        report["cases"][name] = r
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
