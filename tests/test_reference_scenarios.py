"""The scenarios of the reference's own test-suite, on this package.

Each test restates what one test of dohlee/protstruc's `tests/test_StructureBatch.py` / `tests/test_geometry.py`
exercises (cited per test) — same calls, same shapes, same invariants — so a user switching the import finds the
behaviour they rely on.  The reference's tests download entries from the RCSB (1REX: one chain of 130 residues,
4EOT: 184); there is no network here, so structures of exactly those lengths are written as PDB text by a small
generator below and read back through the native ingest, which also covers `from_pdb` end to end.
"""
import math

import numpy as np
import pytest
import torch

from oracle.pdb_fixture_reader import SLOT_NAMES
from protstruc_b200 import ATOM, StructureBatch, constants
from protstruc_b200 import geometry as geom

pytestmark = pytest.mark.gpu

RESIDUE_NAMES = sorted(SLOT_NAMES)


def write_pdb(path, chain_lengths, seed):
    """A protein-like random structure as PDB text: CA random walk with 3.8 A steps, every heavy atom of every
    residue type present (so all 15 slots get exercised), chains named A, B, ..."""
    rng = np.random.default_rng(seed)
    lines, serial = [], 1
    position = np.zeros(3)
    for chain_number, length in enumerate(chain_lengths):
        chain = chr(ord("A") + chain_number)
        for number in range(1, length + 1):
            step = rng.normal(size=3)
            position = position + 3.8 * step / np.linalg.norm(step)
            resname = RESIDUE_NAMES[int(rng.integers(len(RESIDUE_NAMES)))]
            for name in SLOT_NAMES[resname]:
                if not name or name == "OXT":
                    continue
                x, y, z = position + rng.normal(scale=1.2, size=3)
                padded = f" {name:<3s}" if len(name) < 4 else name
                lines.append(f"ATOM  {serial:5d} {padded} {resname} {chain}{number:4d}    {x:8.3f}{y:8.3f}{z:8.3f}  1.00  0.00")
                serial += 1
        lines.append("TER")
    lines.append("END")
    path.write_text("\n".join(lines) + "\n")
    return str(path)


@pytest.fixture(scope="module")
def pdb_files(tmp_path_factory):
    root = tmp_path_factory.mktemp("pdb")
    return {
        "one_chain_130": write_pdb(root / "a130.pdb", [130], 1),    # stands in for 1REX
        "one_chain_184": write_pdb(root / "a184.pdb", [184], 2),    # stands in for 4EOT
        "two_chains_a": write_pdb(root / "hl_a.pdb", [118, 107], 3),  # stand in for the antibody Fv files
        "two_chains_b": write_pdb(root / "hl_b.pdb", [121, 109], 4),
        "two_chains_c": write_pdb(root / "hl_c.pdb", [115, 112], 5),
    }


def random_batch(n_proteins=16, max_n_residues=100, max_n_atoms=25):
    return np.random.default_rng(0).random((n_proteins, max_n_residues, max_n_atoms, 3))


def three_chain_batch():
    xyz = random_batch()
    chain_idx = np.zeros(xyz.shape[:2])
    chain_idx[:, 20:60] = 1.0
    chain_idx[:, 60:] = 2.0
    return StructureBatch.from_xyz(xyz, chain_idx=chain_idx, chain_ids=[["A", "B", "C"]] * len(xyz)), xyz.shape


# ------------------------------------------------------------------ tests/test_StructureBatch.py of the reference
def test_from_xyz_accepts_float64_numpy_and_reports_its_atom_count(native_lib):  # reference :10-21
    sb = StructureBatch.from_xyz(random_batch())
    assert sb.get_max_n_atoms_per_residue() == 25
    assert sb.get_batch_size() == 16 and sb.get_max_n_residues() == 100


def test_from_xyz_with_chain_ids_has_one_terminus_pair_per_chain(native_lib):  # reference :24-40
    sb, (n, length, _, _) = three_chain_batch()
    assert sb.get_n_terminal_mask().shape == (n, length) and sb.get_c_terminal_mask().shape == (n, length)
    assert bool((sb.get_n_terminal_mask().sum(axis=1) == 3).all())
    assert bool((sb.get_c_terminal_mask().sum(axis=1) == 3).all())


def test_from_pdb_single_file(native_lib, pdb_files):  # reference :43-53
    sb = StructureBatch.from_pdb(pdb_files["two_chains_a"])
    assert len(sb.get_xyz()) == 1
    assert bool((sb.get_n_terminal_mask().sum(axis=1) == 2).all())
    assert bool((sb.get_c_terminal_mask().sum(axis=1) == 2).all())


def test_from_pdb_several_files(native_lib, pdb_files):  # reference :56-65
    sb = StructureBatch.from_pdb([pdb_files["two_chains_a"], pdb_files["two_chains_b"], pdb_files["two_chains_c"]])
    assert len(sb.get_xyz()) == 3
    assert bool((sb.get_n_terminal_mask().sum(axis=1) == 2).all())
    assert bool((sb.get_c_terminal_mask().sum(axis=1) == 2).all())
    assert sb.get_total_lengths().tolist() == [225, 230, 227]


def test_backbone_dihedrals_ranges_and_zero_filled_termini(native_lib):  # reference :68-95
    sb, (n, length, _, _) = three_chain_batch()
    dihedrals, dihedral_mask = sb.backbone_dihedrals()
    assert dihedrals.shape == (n, length, 3) and dihedral_mask.shape == (n, length, 3)
    assert bool((dihedrals >= -math.pi).all()) and bool((dihedrals <= math.pi).all())
    assert bool(((dihedrals >= -math.pi) & (dihedrals < 0)).any())
    assert bool(((dihedrals >= 0) & (dihedrals <= math.pi)).any())
    nterm, cterm = sb.get_n_terminal_mask(), sb.get_c_terminal_mask()
    assert bool((dihedrals[nterm][:, 0] == 0.0).all())        # phi is undefined at an N-terminus
    assert bool((dihedrals[cterm][:, [1, 2]] == 0.0).all())   # psi and omega at a C-terminus


def test_pairwise_distance_matrix_shapes_signs_and_atom_enum(native_lib, pdb_files):  # reference :122-137
    sb = StructureBatch.from_pdb(pdb_files["one_chain_130"])
    dist, dist_mask = sb.pairwise_distance_matrix()
    assert dist.shape == (1, 130, 130, 15, 15) and dist_mask.shape == (1, 130, 130, 15, 15)
    ca_dist = dist[:, :, :, ATOM.CA, ATOM.CA]
    cb_dist = dist[:, :, :, ATOM.CB, ATOM.CB]
    assert bool((ca_dist >= 0).all())
    assert bool((cb_dist[~torch.isnan(cb_dist)] >= 0).all())
    assert bool((ca_dist == dist[:, :, :, 1, 1]).all())
    assert bool(torch.isnan(cb_dist).any())  # glycines have no CB: NaN flows through, as in the reference


def test_backbone_orientations_and_translations_shapes(native_lib, pdb_files):  # reference :140-154
    sb = StructureBatch.from_pdb(pdb_files["one_chain_130"])
    assert sb.backbone_orientations("N", "CA", "C").shape == (1, 130, 3, 3)
    for atom in ("N", "CA", "C"):
        assert sb.backbone_translations(atom).shape == (1, 130, 3)


def test_total_lengths_of_a_ragged_batch(native_lib, pdb_files):  # reference :157-163
    sb = StructureBatch.from_pdb([pdb_files["one_chain_130"], pdb_files["one_chain_184"]])
    assert sb.get_total_lengths().tolist() == [130, 184]


def test_pairwise_dihedrals_with_split_atom_lists(native_lib, pdb_files):  # reference :166-176
    sb = StructureBatch.from_pdb([pdb_files["one_chain_130"]])
    phi = sb.pairwise_dihedrals(atoms_i=["C"], atoms_j=["N", "CA", "C"])
    psi = sb.pairwise_dihedrals(atoms_i=["N", "CA", "C"], atoms_j=["N"])
    assert phi.shape == (1, 130, 130) and psi.shape == (1, 130, 130)
    # the (i, i + 1) entries are the backbone angles themselves
    backbone, valid = sb.backbone_dihedrals()
    i = torch.arange(129, device=phi.device)
    same_phi = (phi[0, i, i + 1] - backbone[0, 1:, 0]).abs()
    same_psi = (psi[0, i, i + 1] - backbone[0, :-1, 1]).abs()
    assert float(same_phi[valid[0, 1:, 0]].max()) < 1e-5 and float(same_psi[valid[0, :-1, 1]].max()) < 1e-5


def test_get_local_xyz_shape(native_lib, pdb_files):  # reference :179-186
    sb = StructureBatch.from_pdb([pdb_files["one_chain_130"], pdb_files["one_chain_184"]])
    assert sb.get_local_xyz().shape == (2, 184, sb.get_max_n_atoms_per_residue(), 3)


def test_from_backbone_orientations_translations_round_trip(native_lib, pdb_files):  # reference :189-207
    sb = StructureBatch.from_pdb([pdb_files["one_chain_130"]])
    orientations, translations = sb.backbone_orientations(), sb.backbone_translations()
    args = (orientations, translations, sb.get_chain_idx(), sb.get_chain_ids(), sb.get_seq())
    sb2 = StructureBatch.from_backbone_orientations_translations(*args)
    sb3 = StructureBatch.from_backbone_orientations_translations(*args, include_cb=True)
    assert sb2.get_max_n_atoms_per_residue() == 15 and sb3.get_max_n_atoms_per_residue() == 15
    # the rebuilt backbone carries the frames it was built from
    assert torch.allclose(sb2.backbone_orientations(), orientations, atol=2e-5)
    assert torch.allclose(sb2.backbone_translations(), translations, atol=1e-4)


def test_standardize_family(native_lib, pdb_files):  # reference :210-255
    sb = StructureBatch.from_pdb([pdb_files["one_chain_130"]])
    original = sb.get_xyz().clone()
    atom_mask = sb.get_atom_mask()
    with pytest.raises(ValueError):
        sb.unstandardize()                      # cannot unstandardize first
    sb.standardize()
    assert not bool(torch.isnan(sb.get_xyz()[atom_mask.bool()]).any())
    with pytest.raises(ValueError):
        sb.standardize()                        # cannot standardize twice
    sb.unstandardize()
    assert torch.allclose(original, sb.get_xyz(), equal_nan=True, rtol=1e-4, atol=1e-5)


def test_center_at_origin_and_at_given_points(native_lib, pdb_files):  # reference :258-275
    sb = StructureBatch.from_pdb([pdb_files["one_chain_130"]])
    sb.center_at()
    com = sb.center_of_mass()
    assert torch.allclose(com, torch.zeros_like(com), rtol=1e-4, atol=1e-5)
    sb = StructureBatch.from_pdb([pdb_files["one_chain_130"], pdb_files["one_chain_184"]])
    centers = torch.randn(2, 3, generator=torch.Generator().manual_seed(0))
    sb.center_at(centers)
    assert torch.allclose(sb.center_of_mass().cpu(), centers, rtol=1e-4, atol=1e-4)


def test_residue_mask_seq_idx_and_masked_select(native_lib, pdb_files):  # reference :278-305
    sb = StructureBatch.from_pdb([pdb_files["one_chain_130"], pdb_files["one_chain_184"]])
    residue_mask = sb.get_residue_mask()
    seq_idx = sb.get_seq_idx()
    assert residue_mask.shape == (2, 184) and seq_idx.shape == (2, 184)
    assert bool((seq_idx[~residue_mask.bool()] == 20).all())  # AA.UNK beyond the structure
    one = StructureBatch.from_pdb([pdb_files["one_chain_130"]])
    pick = torch.randint(0, 2, size=one.get_residue_mask().shape, generator=torch.Generator().manual_seed(1)).bool()
    assert one.residue_masked_select(pick).get_xyz().shape == (1, int(pick.sum()), 15, 3)


# ------------------------------------------------------------------ tests/test_geometry.py of the reference
def test_geometry_primitives_numpy_in_numpy_out_tensor_in_tensor_out(native_lib):  # reference :10-89, decorator tests
    a = np.array([[1.0, 2.0, 3.0]], dtype=np.float32)
    b = np.array([[4.0, 5.0, 6.0]], dtype=np.float32)
    assert isinstance(geom.dot(a, b), np.ndarray) and float(geom.dot(a, b).reshape(-1)[0]) == 32.0
    assert isinstance(geom.dot(torch.from_numpy(a), torch.from_numpy(b)), torch.Tensor)
    assert abs(float(geom.norm(a).reshape(-1)[0]) - math.sqrt(14.0)) < 1e-6
    right = geom.angle(np.array([[1.0, 0, 0]], np.float32), np.zeros((1, 3), np.float32), np.array([[0, 1.0, 0]], np.float32),
                       to_degree=True)
    assert abs(float(np.asarray(right).reshape(-1)[0]) - 90.0) < 1e-4


def test_dihedral_sign_convention_and_leading_dimensions(native_lib):  # reference :92-190
    p = [torch.tensor(v, dtype=torch.float32) for v in ([[1.0, 0, 0]], [[0.0, 0, 0]], [[0.0, 1, 0]], [[0.0, 1, 1]])]
    assert abs(float(geom.dihedral(*p, to_degree=True).reshape(-1)[0]) + 90.0) < 1e-4
    stacked = [q.expand(2, 5, 7, 3).contiguous() for q in p]
    out = geom.dihedral(*stacked)
    assert out.shape[:3] == (2, 5, 7) and bool(((out + math.pi / 2).abs() < 1e-6).all())


def test_gram_schmidt_frames_are_orthonormal_and_ideal_backbone_is_identity(native_lib):  # reference :235-262
    g = torch.Generator().manual_seed(5)
    a, b, c = (torch.randn(4, 9, 3, generator=g) for _ in range(3))
    frames = geom.gram_schmidt(a, b, c).cpu()
    eye = torch.eye(3).expand(4, 9, 3, 3)
    assert torch.allclose(frames.transpose(-1, -2) @ frames, eye, atol=1e-5)
    assert torch.allclose(torch.linalg.det(frames), torch.ones(4, 9), atol=1e-5)
    ideal = constants.ideal_backbone()  # (3, 3): N, CA, C of the ideal residue
    frame = geom.gram_schmidt(ideal[None, 0], ideal[None, 1], ideal[None, 2]).cpu()
    assert torch.equal(frame, torch.eye(3)[None])  # the reference asserts exact identity here


def test_align_superimposes_a_moved_copy(native_lib, pdb_files):  # reference :265-283 (kabsch) and protstruc.py:880-918
    target = StructureBatch.from_pdb([pdb_files["one_chain_130"]])
    moved = StructureBatch.from_pdb([pdb_files["one_chain_130"]])
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=torch.Generator().manual_seed(3)))
    if torch.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    moved.rotate(q.unsqueeze(0))
    moved.translate(torch.tensor([[[3.0, -2.0, 7.5]]]))
    moved.align(target)
    valid = target.get_atom_mask().bool()
    assert float((moved.get_xyz()[valid] - target.get_xyz()[valid]).abs().max()) < 1e-3
