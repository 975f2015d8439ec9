"""Generates the golden vectors under tests/golden/ FROM THE REAL REFERENCE.

Run in the build container only (the reference tree is not available on the GPU box):

    python tests/golden/make_golden.py            # needs /root/reference

What it does
  1. imports the unmodified reference package from /root/reference under a stub for `biotite`
     (not installed; only the PDB constructors need it);
  2. runs every hot-path method of the reference on seeded synthetic inputs and on a real structure
     (tests/1a6v_HL.pdb of the reference, parsed by oracle/pdb_fixture_reader.py);
  3. checks the oracle (oracle/feature_oracle.py) against those reference outputs — bit-exact is
     expected because the op sequence is the same — and records the maximum deviation per output in
     tests/golden/MANIFEST.json;
  4. stores inputs + reference outputs as compressed .npz fixtures, which the CPU suite
     (oracle vs golden) and the GPU suite (CUDA vs golden) both consume.
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path
from unittest import mock

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REFERENCE = Path("/root/reference")
sys.path.insert(0, str(REPO))

from oracle import feature_oracle as orc  # noqa: E402
from oracle import pdb_fixture_reader  # noqa: E402


def import_reference():
    """Registers stub modules for biotite and imports the reference package."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    stub("biotite")
    stub("biotite.database")
    stub("biotite.database.rcsb", fetch=lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no network")))
    stub("biotite.structure", AtomArray=object)
    stub("biotite.structure.io")
    stub("biotite.structure.io.pdb", PDBFile=object)
    sys.path.insert(0, str(REFERENCE))
    import protstruc  # noqa: F401
    import protstruc.geometry as geom
    from protstruc import StructureBatch
    return StructureBatch, geom


def synthetic_inputs(seed: int, B: int, L: int, A: int, mask_kind: str = "bool", nan_masked: bool = True):
    """Protein-like random batch: CA random walk + atom offsets, Bernoulli(0.7) slot occupancy with the
    backbone always present, masked slots NaN (as the reference's PDB ingest yields), a zero-padded
    tail with chain_idx = NaN, two chains."""
    g = torch.Generator().manual_seed(seed)
    steps = torch.randn(B, L, 3, generator=g)
    steps = 3.8 * steps / steps.norm(dim=-1, keepdim=True)
    ca = steps.cumsum(dim=1)
    xyz = ca[:, :, None, :] + 1.5 * torch.randn(B, L, A, 3, generator=g)
    mask = torch.rand(B, L, A, generator=g) < 0.7
    mask[:, :, : min(A, 4)] = True
    # a few residues without CB (glycine-like) and a fully missing gap residue
    if A > 4:
        mask[:, 1::5, 4] = False
    if L > 4:
        mask[0, 3, :] = False
    lengths = [L - (b % 3) for b in range(B)]
    chain_idx = torch.zeros(B, L)
    for b in range(B):
        chain_idx[b, lengths[b] // 2: lengths[b]] = 1.0
        chain_idx[b, lengths[b]:] = float("nan")
        mask[b, lengths[b]:] = False
    if nan_masked:
        xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
    for b in range(B):
        xyz[b, lengths[b]:] = 0.0
    if mask_kind == "float":
        mask = mask.float()
    return xyz.contiguous(), mask, chain_idx


def run_reference(StructureBatch, xyz, atom_mask, chain_idx, noise_seed=7):
    """All hot-path outputs of the reference for one batch."""
    B = xyz.shape[0]
    ids = [["A", "B"] for _ in range(B)]
    out = {}
    sb = StructureBatch.from_xyz(xyz.clone(), atom_mask.clone(), chain_idx.clone(), ids)
    out["dist"], out["dist_mask"] = sb.pairwise_distance_matrix()
    if xyz.shape[2] >= 5:
        g = sb.inter_residue_geometry()
        for k in ("omega", "theta", "phi"):
            out[k] = g[k]
        out["psi_like_dihedral_N_CA_C_N"] = sb.pairwise_dihedrals(["N", "CA", "C"], ["N"])
        out["planar_CA_CA_C"] = sb.pairwise_planar_angles(["CA"], ["CA", "C"])
    out["bb_dihedrals"], out["bb_dihedral_mask"] = sb.backbone_dihedrals()
    out["nterm"], out["cterm"] = sb.get_n_terminal_mask(), sb.get_c_terminal_mask()
    out["frames"] = sb.backbone_orientations()
    out["com"] = sb.center_of_mass()
    # standardize: the reference is only valid for B == 1 (Q1) -> run it structure by structure
    std_xyz, mus, sds = [], [], []
    for b in range(B):
        one = StructureBatch.from_xyz(xyz[b:b + 1].clone(), atom_mask[b:b + 1].clone())
        one.standardize()
        std_xyz.append(one.get_xyz())
        mus.append(one.mu)
        sds.append(one.std)
    out["std_xyz"], out["mu"], out["sd"] = torch.cat(std_xyz), torch.cat(mus), torch.cat(sds)
    # diffuse_xyz with injected noise: patch randn_like inside the reference call
    g = torch.Generator().manual_seed(noise_seed)
    noise = torch.randn(xyz.shape, generator=g)
    beta = torch.linspace(0.0, 0.75, B) if B > 1 else torch.tensor([0.3])
    sb2 = StructureBatch.from_xyz(xyz.clone(), atom_mask.clone())
    with mock.patch("torch.randn_like", lambda t: noise):
        sb2.diffuse_xyz(beta)
    out["noise"], out["beta"], out["diffused"] = noise, beta, sb2.get_xyz()
    return out


def run_oracle(xyz, atom_mask, chain_idx, noise, beta):
    residue_mask = atom_mask.any(dim=-1)
    out = {}
    out["dist"], out["dist_mask"] = orc.pair_distances(xyz, atom_mask)
    if xyz.shape[2] >= 5:
        out["omega"], out["theta"], out["phi"] = orc.trrosetta_angles(xyz)
        out["psi_like_dihedral_N_CA_C_N"] = orc.pair_dihedrals(xyz, [0, 1, 2], [0])
        out["planar_CA_CA_C"] = orc.pair_planar_angles(xyz, [1], [1, 2])
    out["bb_dihedrals"], out["bb_dihedral_mask"] = orc.backbone_dihedrals(xyz, chain_idx, residue_mask)
    out["nterm"], out["cterm"] = orc.terminal_masks(chain_idx, residue_mask)
    out["frames"] = orc.frames(xyz)
    out["com"] = orc.center_of_mass(xyz)
    out["std_xyz"], out["mu"], out["sd"] = orc.standardize_per_structure(xyz, atom_mask)
    out["diffused"] = orc.diffuse(xyz, beta, noise)
    return out


def random_rotations(n: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    q, r = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))
    q = q * torch.sign(torch.diagonal(r, dim1=-2, dim2=-1))[:, None, :]
    q[:, :, 2] *= torch.linalg.det(q)[:, None]  # proper rotations
    return q.contiguous()


def make_frames_align_topk(StructureBatch):
    """get_local_xyz, rotate, translate, from_backbone_orientations_translations, align and
    get_topk_nearest_residue_mask of the reference on a seeded synthetic batch (B = 4, L = 33, no NaN: the
    reference's align / frames need finite backbone coordinates) and on the real structure for top-k."""
    xyz, mask, chain_idx = synthetic_inputs(21, 4, 33, 15, "bool", nan_masked=False)
    B = xyz.shape[0]
    ids = [["A", "B"] for _ in range(B)]
    new = lambda: StructureBatch.from_xyz(xyz.clone(), mask.clone(), chain_idx.clone(), ids)  # noqa: E731
    out = {"xyz": xyz, "atom_mask": mask, "chain_idx": chain_idx}
    out["ref_local_xyz"] = new().get_local_xyz()
    rot = random_rotations(B, 5)
    out["rotation"] = rot
    sb = new()
    sb.rotate(rot)
    out["ref_rotated"] = sb.get_xyz().clone()
    sb = new()
    sb.rotate(rot[0])
    out["ref_rotated_single"] = sb.get_xyz().clone()
    g = torch.Generator().manual_seed(6)
    tr_res = torch.randn(B, 33, 3, generator=g)
    tr_one = torch.randn(B, 1, 3, generator=g)
    tr_atom = torch.randn(B, 33, 15, 3, generator=g)
    out.update({"tr_res": tr_res, "tr_one": tr_one, "tr_atom": tr_atom})
    for name, t, atomwise in (("res", tr_res, False), ("one", tr_one, False), ("atom", tr_atom, True)):
        sb = new()
        sb.translate(t.clone(), atomwise=atomwise)
        out[f"ref_translated_{name}"] = sb.get_xyz().clone()
    sb = new()
    frames, trans = sb.backbone_orientations(), sb.backbone_translations().clone()
    out["frames"], out["frame_translations"] = frames, trans
    for cb in (False, True):
        sb2 = StructureBatch.from_backbone_orientations_translations(frames, trans, chain_idx.clone(), ids, None,
                                                                     include_cb=cb)
        out[f"ref_from_frames_xyz_cb{int(cb)}"] = sb2.get_xyz()
        out[f"ref_from_frames_mask_cb{int(cb)}"] = sb2.get_atom_mask()
    # align: target = rigidly moved + slightly noised copy (per-structure targets and a single target)
    moved = torch.einsum("bij,bnaj->bnai", random_rotations(B, 9), xyz) + torch.randn(B, 1, 1, 3, generator=g) * 5
    moved = moved + 0.05 * torch.randn(moved.shape, generator=g)
    out["align_target"] = moved
    src, tgt = new(), StructureBatch.from_xyz(moved.clone(), mask.clone(), chain_idx.clone(), ids)
    src.align(tgt)
    out["ref_aligned"] = src.get_xyz().clone()
    # the reference zips over the target batch, so a single target only aligns... run it per structure instead
    single = []
    for b in range(B):
        one = StructureBatch.from_xyz(xyz[b:b + 1].clone(), mask[b:b + 1].clone())
        one.align(StructureBatch.from_xyz(moved[:1].clone(), mask[:1].clone()),
                  atom_mask=mask[b:b + 1] * mask[:1])
        single.append(one.get_xyz())
    out["ref_aligned_to_first"] = torch.cat(single)
    # top-k on the real structure (B = 1)
    arrays = pdb_fixture_reader.read_batch([REFERENCE / "tests" / "1a6v_HL.pdb"])
    rx, rm = torch.from_numpy(arrays["xyz"]), torch.from_numpy(arrays["atom_mask"])
    real = StructureBatch.from_xyz(rx, rm)
    query = rx[0, 100:103, 1].clone() + 0.3
    out["topk_query"] = query
    out["ref_topk_k32"] = real.get_topk_nearest_residue_mask(query, k=32)
    out["ref_topk_k500"] = real.get_topk_nearest_residue_mask(query, k=500)
    extra = torch.zeros(229, dtype=torch.bool)
    extra[50:150] = True
    out["topk_extra_mask"] = extra
    out["ref_topk_k16_masked"] = real.get_topk_nearest_residue_mask(query, k=16, mask=extra)

    # oracle vs reference
    dev = {}
    dev["local_xyz"] = max_deviation(out["ref_local_xyz"], orc.local_xyz(xyz))
    dev["rotated"] = max_deviation(out["ref_rotated"], orc.rotate(xyz, rot))
    dev["rotated_single"] = max_deviation(out["ref_rotated_single"], orc.rotate(xyz, rot[0]))
    for cb in (False, True):
        ox, om = orc.frames_to_backbone(frames, trans, cb)
        dev[f"from_frames_xyz_cb{int(cb)}"] = max_deviation(out[f"ref_from_frames_xyz_cb{int(cb)}"], ox)
        dev[f"from_frames_mask_cb{int(cb)}"] = max_deviation(out[f"ref_from_frames_mask_cb{int(cb)}"], om)
    dev["aligned"] = max_deviation(out["ref_aligned"], orc.align(xyz, moved, mask * mask)[0])
    dev["aligned_to_first"] = max_deviation(out["ref_aligned_to_first"], orc.align(xyz, moved[:1], mask * mask[:1])[0])
    rmask = rm.any(-1)
    dev["topk_k32"] = max_deviation(out["ref_topk_k32"], orc.topk_nearest_residue_mask(rx, rmask, query, 32))
    dev["topk_k500"] = max_deviation(out["ref_topk_k500"], orc.topk_nearest_residue_mask(rx, rmask, query, 500))
    dev["topk_k16_masked"] = max_deviation(out["ref_topk_k16_masked"],
                                           orc.topk_nearest_residue_mask(rx, rmask, query, 16, extra))
    np.savez_compressed(HERE / "frames_align_topk.npz", **{k: to_np(v) for k, v in out.items()})
    print("frames_align_topk", dev)
    return {"shape": [B, 33, 15], "mask": "bool", "oracle_vs_reference_max_abs": dev}


def max_deviation(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    if a.dtype == torch.bool or b.dtype == torch.bool:
        return float((a.bool() != b.bool()).sum())
    a, b = a.double(), b.double()
    nan_mismatch = (torch.isnan(a) != torch.isnan(b)).sum().item()
    if nan_mismatch:
        return float("inf")
    d = (a - b).abs()
    d = d[~torch.isnan(d)]
    return float(d.max()) if d.numel() else 0.0


def to_np(v):
    return v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)


def main():
    StructureBatch, geom = import_reference()
    manifest = {"torch": torch.__version__, "numpy": np.__version__, "cases": {}}
    torch.set_num_threads(1)  # deterministic reduction order

    cases = {
        # name: (seed, B, L, A, mask_kind)
        "synthetic_small": (11, 2, 12, 15, "bool"),
        "synthetic_floatmask_oddL": (12, 1, 7, 15, "float"),
        "synthetic_A5": (13, 2, 9, 5, "bool"),
        "synthetic_A25_like_reference_tests": (14, 1, 5, 25, "bool"),
        "synthetic_ragged_33": (15, 4, 33, 15, "bool"),
    }
    for name, (seed, B, L, A, kind) in cases.items():
        xyz, mask, chain_idx = synthetic_inputs(seed, B, L, A, kind)
        ref = run_reference(StructureBatch, xyz, mask, chain_idx)
        mine = run_oracle(xyz, mask, chain_idx, ref["noise"], ref["beta"])
        dev = {k: max_deviation(ref[k], mine[k]) for k in mine}
        manifest["cases"][name] = {"shape": [B, L, A], "mask": kind, "oracle_vs_reference_max_abs": dev}
        payload = {"xyz": to_np(xyz), "atom_mask": to_np(mask), "chain_idx": to_np(chain_idx)}
        payload.update({f"ref_{k}": to_np(v) for k, v in ref.items()})
        np.savez_compressed(HERE / f"{name}.npz", **payload)
        print(name, dev)

    # real structure: config 1 of BASELINE.json (tests/1a6v_HL.pdb of the reference)
    arrays = pdb_fixture_reader.read_batch([REFERENCE / "tests" / "1a6v_HL.pdb"])
    xyz = torch.from_numpy(arrays["xyz"])
    mask = torch.from_numpy(arrays["atom_mask"])
    chain_idx = torch.from_numpy(arrays["chain_idx"])
    ref = run_reference(StructureBatch, xyz, mask, chain_idx)
    mine = run_oracle(xyz, mask, chain_idx, ref["noise"], ref["beta"])
    dev = {k: max_deviation(ref[k], mine[k]) for k in mine}
    manifest["cases"]["real_1a6v_HL"] = {"shape": list(xyz.shape[:3]), "mask": "bool",
                                         "oracle_vs_reference_max_abs": dev,
                                         "n_atoms": int(mask.sum()), "chain_ids": arrays["chain_ids"]}
    CA, CB, N, O = 1, 4, 0, 3
    crop = 40
    payload = {
        "xyz": to_np(xyz), "atom_mask": to_np(mask), "chain_idx": to_np(chain_idx),
        # compact features of the full structure
        "ref_d_ca": to_np(ref["dist"][:, :, :, CA, CA]), "ref_d_cb": to_np(ref["dist"][:, :, :, CB, CB]),
        "ref_d_no": to_np(ref["dist"][:, :, :, N, O]),
        "ref_d_ca_mask": to_np(ref["dist_mask"][:, :, :, CA, CA]),
        # full all-atom block of the first `crop` residues
        "ref_dist_crop": to_np(ref["dist"][:, :crop, :crop]), "ref_dist_mask_crop": to_np(ref["dist_mask"][:, :crop, :crop]),
        # order-independent checksums of the full tensors (float64 sums over finite entries)
        "ref_dist_nansum": np.float64(torch.nansum(ref["dist"].double()).item()),
        "ref_dist_nan_count": np.int64(torch.isnan(ref["dist"]).sum().item()),
        "ref_dist_mask_sum": np.int64(ref["dist_mask"].sum().item()),
    }
    for k in ("omega", "theta", "phi", "bb_dihedrals", "bb_dihedral_mask", "nterm", "cterm", "frames", "com",
              "mu", "sd", "std_xyz", "noise", "beta", "diffused"):
        payload[f"ref_{k}"] = to_np(ref[k])
    np.savez_compressed(HERE / "real_1a6v_HL.npz", **payload)
    print("real_1a6v_HL", dev)

    # rows f1 / f2 / f4: rigid-frame family, alignment, top-k neighbours
    manifest["cases"]["frames_align_topk"] = make_frames_align_topk(StructureBatch)

    # the reference's own known-answer tests for the primitives (tests/test_geometry.py)
    ka = {}
    ka["dot"] = float(geom.dot(torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6])))
    a = torch.tensor([[1.0, 0, 0], [1.0, 0, 0]])
    b = torch.zeros(2, 3)
    c = torch.tensor([[0, 1.0, 0], [0.5, np.sqrt(3) / 2, 0]]).float()
    ka["angle_deg"] = geom.angle(a, b, c, to_degree=True).tolist()
    ka["dihedral_deg"] = geom.dihedral(torch.tensor([[1.0, 0, 0]]), torch.zeros(1, 3), torch.tensor([[0, 1.0, 0]]),
                                       torch.tensor([[0, 1.0, 1.0]]), to_degree=True).tolist()
    ideal = geom.ideal_backbone_coordinates(size=(2, 4), include_cb=True)
    ka["ideal_backbone_n_ca_c_cb"] = ideal[0, 0].tolist()
    ka["ideal_frame_is_identity"] = bool((geom.gram_schmidt(ideal[:, :, 0], ideal[:, :, 1], ideal[:, :, 2])
                                          == torch.eye(3).expand(2, 4, -1, -1)).all())
    sched = orc.cosine_variance_schedule(300)
    ka["cosine_beta_1_299"] = [float(sched[0]), float(sched[1]), float(sched[299])]
    manifest["known_answers"] = ka

    (HERE / "MANIFEST.json").write_text(json.dumps(manifest, indent=1, sort_keys=True))
    worst = max(max(c["oracle_vs_reference_max_abs"].values()) for c in manifest["cases"].values())
    print("worst oracle-vs-reference deviation:", worst)


if __name__ == "__main__":
    main()
