"""CPU suite, part 3: host-side logic of the façade — argument validation with the reference's
exception types, index bookkeeping, batch sharding (incl. a world_size-2 gloo run) — and the
guarantee that nothing computes on the CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import protstruc_b200 as ps
from protstruc_b200 import _cabi, sharding
from oracle import feature_oracle as orc
from tests import helpers as H


def cpu_batch(B=2, L=10, A=15, **kw):
    xyz, mask, chain_idx = H.synthetic_batch(1, B, L, A)
    ids = [["A", "B"]] * B
    return ps.StructureBatch.from_xyz(xyz, mask, chain_idx, ids, device="cpu", **kw), xyz, mask, chain_idx


def test_constructor_contract():
    """reference protstruc/protstruc.py:55-91."""
    xyz = np.random.rand(4, 20, 25, 3)
    sb = ps.StructureBatch.from_xyz(xyz, device="cpu")
    assert sb.get_batch_size() == 4 and sb.get_max_n_residues() == 20 and sb.get_max_n_atoms_per_residue() == 25
    assert sb.get_xyz().dtype == torch.float32
    assert sb.residue_mask.dtype == torch.bool and bool(sb.residue_mask.all())
    assert tuple(sb.chain_idx.shape) == (4, 20) and bool((sb.chain_idx == 0).all())
    with pytest.raises(ValueError, match="Both `chain_idx` and `chain_ids`"):
        ps.StructureBatch.from_xyz(xyz, chain_idx=np.zeros((4, 20)), device="cpu")
    with pytest.raises(ValueError, match="Both `chain_idx` and `chain_ids`"):
        ps.StructureBatch.from_xyz(xyz, chain_ids=[["A"]] * 4, device="cpu")
    with pytest.raises(AssertionError, match="Chain index should start from zero"):
        ps.StructureBatch.from_xyz(xyz, chain_idx=np.ones((4, 20)), chain_ids=[["A"]] * 4, device="cpu")
    with pytest.raises(ValueError):
        ps.StructureBatch.from_xyz(np.zeros((4, 20, 3)), device="cpu")


def test_terminal_masks_and_getters_match_the_oracle():
    sb, xyz, mask, chain_idx = cpu_batch(B=4, L=33)
    nterm, cterm = orc.terminal_masks(chain_idx, mask.any(-1))
    assert torch.equal(sb.get_n_terminal_mask(), nterm) and torch.equal(sb.get_c_terminal_mask(), cterm)
    assert torch.equal(sb.get_residue_mask(), mask[:, :, 1].bool())
    assert sb.get_chain_idx().dtype == torch.int64
    assert torch.equal(sb.get_total_lengths(), mask.any(-1).cumsum(1).argmax(1) + 1)
    # three chains -> three termini each (reference tests/test_StructureBatch.py:24-40)
    ci = np.zeros((16, 100))
    ci[:, 20:60] = 1.0
    ci[:, 60:] = 2.0
    sb3 = ps.StructureBatch.from_xyz(np.random.rand(16, 100, 25, 3), chain_idx=ci, chain_ids=[["A", "B", "C"]] * 16,
                                     device="cpu")
    assert bool((sb3.get_n_terminal_mask().sum(axis=1) == 3).all())
    assert bool((sb3.get_c_terminal_mask().sum(axis=1) == 3).all())


def test_validation_errors_are_raised_before_any_launch():
    sb, xyz, mask, _ = cpu_batch()
    with pytest.raises(ValueError, match="Atom QQ is not valid."):
        sb.pairwise_dihedrals(["QQ", "CB"], ["CA", "CB"])
    with pytest.raises(ValueError, match="Atom zz is not valid."):
        sb.pairwise_planar_angles(["CA", "CB"], ["zz"])
    with pytest.raises(KeyError):
        sb.backbone_orientations(a2="CX")
    with pytest.raises(KeyError):
        sb.backbone_translations("CX")
    with pytest.raises(ValueError, match="Only one of atom_mask and residue_mask"):
        sb.standardize(atom_mask=mask, residue_mask=mask.any(-1))
    with pytest.raises(ValueError, match="Cannot unstandardize"):
        sb.unstandardize()
    sb._standardized = True
    with pytest.raises(ValueError, match="already standardized"):
        sb.standardize()
    sb._standardized = False
    with pytest.raises(ValueError, match="`center` must have a shape"):
        sb.center_at(torch.zeros(2, 2))
    with pytest.raises(ValueError, match="`center` must have a shape"):
        sb.center_at(torch.zeros(5, 3))
    assert sb.backbone_translations().shape == (2, 10, 3)


def test_there_is_no_cpu_compute_path():
    sb, *_ = cpu_batch()
    for call in (sb.pairwise_distance_matrix, sb.inter_residue_geometry, sb.backbone_dihedrals,
                 sb.backbone_orientations, sb.center_of_mass, sb.standardize,
                 lambda: sb.pairwise_dihedrals(["CA", "CB"], ["CA", "CB"]),
                 lambda: sb.diffuse_xyz(torch.zeros(2))):
        with pytest.raises(_cabi.NativeLibraryError, match="no CPU fallback"):
            call()
    if not torch.cuda.is_available():
        with pytest.raises(_cabi.NativeLibraryError):
            ps.geometry.angle(np.zeros((1, 3), np.float32), np.zeros((1, 3), np.float32), np.zeros((1, 3), np.float32))


def test_product_package_never_imports_the_oracle():
    from pathlib import Path
    pkg = Path(ps.__file__).resolve().parent
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        text = f.read_text()
        assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"
    # the measurement scripts under tools/ do not import it either (those that need the checker live in tests/)
    for f in (pkg.parent / "tools").glob("*.py"):
        text = f.read_text()
        assert "from oracle" not in text and "import oracle" not in text, f"{f} imports the oracle"


def test_atom_vocabulary():
    """reference protstruc/general.py:4-23 and tests/test_constants.py."""
    assert [int(ps.ATOM[n]) for n in ("N", "CA", "C", "O", "CB")] == [0, 1, 2, 3, 4]
    assert ps.ATOM["ca"] == ps.ATOM.CA and ps.ATOM["Cb"] == 4 and ps.ATOM.is_valid("cb") and not ps.ATOM.is_valid("CG")
    assert ps.MAX_N_ATOMS_PER_RESIDUE == 15


def test_shard_bounds_partition_the_batch():
    for B in (0, 1, 7, 8, 64, 4096):
        for world in (1, 2, 3, 4, 8):
            spans = [sharding.shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(B, world)
    with pytest.raises(ValueError):
        sharding.shard_bounds(8, 2, 2)
    xyz, mask, chain_idx = H.synthetic_batch(3, 8, 16, 15)
    part = sharding.shard_structure_batch(xyz, mask, chain_idx, [["A", "B"]] * 8, rank=1, world_size=4, device="cpu")
    assert part.get_batch_size() == 2 and torch.equal(torch.nan_to_num(part.get_xyz()), torch.nan_to_num(xyz[2:4]))
    assert part._noise_elem_offset == 2 * 16 * 15 * 3
    # odd structure sizes (L = 229: 10,305 floats) keep the exact global offset: no per-rank fallback stream
    x229 = torch.zeros(5, 229, 15, 3)
    offsets = [sharding.shard_structure_batch(x229, rank=r, world_size=3, device="cpu")._noise_elem_offset for r in range(3)]
    assert offsets == [0, 2 * 10_305, 4 * 10_305]
    assert sharding.shard_structure_batch(xyz[:1], mask[:1], rank=1, world_size=2, device="cpu") is None


def test_noise_stream_sessions_follow_the_generator_state():
    """The (key, step) bookkeeping behind diffuse_xyz (no GPU needed): one session per generator object, continued
    while nobody else touches the generator, restarted reproducibly by any re-seed."""
    from protstruc_b200 import structure_batch as sbm

    stream = sbm._PhiloxStream()
    g1, g2 = torch.Generator().manual_seed(1), torch.Generator().manual_seed(2)
    a = [stream.reserve(1, g1), stream.reserve(1, g2), stream.reserve(3, g1), stream.reserve(1, g2), stream.reserve(1, g1)]
    assert a[0][0] == a[2][0] == a[4][0] != a[1][0] == a[3][0]          # one key per generator
    assert [x[1] for x in a] == [0, 0, 1, 1, 4]                          # steps advance, never rewind
    assert stream.reserve(2, torch.Generator().manual_seed(1)) == (a[0][0], 0)   # re-created generator, same seed
    g1.manual_seed(1)
    assert stream.reserve(1, g1) == (a[0][0], 0)                         # re-seeded in place
    torch.manual_seed(9)
    k0 = stream.reserve(300, None)
    assert stream.reserve(1, None) == (k0[0], 300)
    torch.rand(1)                                                        # the global generator was used by someone else
    k1 = stream.reserve(1, None)
    assert k1[0] != k0[0] and k1[1] == 0
    torch.manual_seed(9)
    assert stream.reserve(1, None) == (k0[0], 0)                         # torch.manual_seed restarts the stream
    sbm.manual_seed(9)
    assert sbm._philox.reserve(2, None) == (k0[0], 0)
    assert all(0 <= key < 2 ** 63 for key in (k0[0], k1[0], a[0][0], a[1][0]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, B, L, result_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        xyz, mask, chain_idx = H.synthetic_batch(77, B, L, 15)
        start, stop = sharding.shard_bounds(B, world, rank)
        # the CPU oracle stands in for the kernels here: this test covers the sharding / gather plumbing
        omega, theta, phi = orc.trrosetta_angles(xyz[start:stop])
        d, _ = orc.pair_distances(xyz[start:stop], mask[start:stop])
        local = {"omega": omega, "theta": theta, "phi": phi, "d_ca": d[:, :, :, 1, 1].contiguous()}
        full = sharding.gather_compact_features(local, B)
        if rank == 0:
            torch.save(full, os.path.join(result_dir, "gathered.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 4, 1])  # uneven shards, even shards, an EMPTY shard on rank 1
def test_world_size_two_gloo_shards_reassemble_to_the_unsharded_result(tmp_path, B):
    L, world = 12, 2
    mp.spawn(_gloo_worker, args=(world, _free_port(), B, L, str(tmp_path)), nprocs=world, join=True)
    gathered = torch.load(tmp_path / "gathered.pt")
    torch.set_num_threads(1)
    xyz, mask, _ = H.synthetic_batch(77, B, L, 15)
    omega, theta, phi = orc.trrosetta_angles(xyz)
    d, _ = orc.pair_distances(xyz, mask)
    for name, ref in (("omega", omega), ("theta", theta), ("phi", phi), ("d_ca", d[:, :, :, 1, 1])):
        got = gathered[name]
        assert got.shape == ref.shape
        assert torch.equal(torch.nan_to_num(got, nan=-9.0), torch.nan_to_num(ref, nan=-9.0)), name


# ------------------------------------------------------------------------------ row f3: native PDB ingest (host code)
def test_native_pdb_ingest_on_the_synthetic_fixture(native_lib):
    """tests/golden/mini_two_chain.pdb: altlocs, MSE HETATM, hydrogens, water, numbering gap, insertion code,
    OXT, a second MODEL.  Rules: reference protstruc/pdb.py:24-40, 55-151."""
    from protstruc_b200 import pdb_ingest
    from oracle import pdb_fixture_reader as reader

    path = H.GOLDEN / "mini_two_chain.pdb"
    a = pdb_ingest.read_pdb_arrays(path)
    assert a["one_letter"] == "AGSMXXKWDAGDS"            # MSE -> M, two UNK placeholders for the 4 -> 7 gap
    assert a["chain_ids"] == ["H", "L"] and a["seq"] == {"H": "AGSMXXKWDA", "L": "GDS"}
    assert list(a["residue_number"]) == [1, 2, 3, 4, 5, 6, 7, 8, 8, 9, 1, 2, 3]
    assert a["insertion_code"][8] == "A" and a["insertion_code"][7] == ""
    assert list(a["atom_mask"].sum(1)) == [5, 4, 6, 8, 0, 0, 8, 14, 8, 6, 4, 8, 7]   # no H, no SE, no HOH, LYS lacks NZ
    assert np.isnan(a["xyz"][4]).all() and not a["atom_mask"][1, 4]                    # UNK row / glycine CB
    assert a["atom_mask"][9, 14] and a["atom_mask"][12, 14]                             # OXT in slot 14
    x, m, c, ids = reader.read_structure(path)                                          # independent Python restatement
    assert np.array_equal(a["atom_mask"], m) and np.array_equal(a["chain_idx"], c) and ids == a["chain_ids"]
    assert np.array_equal(np.nan_to_num(a["xyz"], nan=-9.0), np.nan_to_num(x, nan=-9.0))
    # first alternate location wins: SER 3 CB is the 'A' conformer
    text = path.read_text().splitlines()
    cb_a = next(l for l in text if l[12:16].strip() == "CB" and l[17:20] == "SER" and l[16] == "A")
    assert np.allclose(a["xyz"][2, 4], [float(cb_a[30:38]), float(cb_a[38:46]), float(cb_a[46:54])])


def test_from_pdb_pads_like_the_reference(native_lib):
    path = str(H.GOLDEN / "mini_two_chain.pdb")
    sb = ps.StructureBatch.from_pdb([path, path], device="cpu")
    assert tuple(sb.get_xyz().shape) == (2, 13, 15, 3) and sb.get_atom_mask().dtype == torch.bool
    assert sb.get_chain_ids() == [["H", "L"], ["H", "L"]]
    assert bool((sb.get_n_terminal_mask().sum(1) == 2).all()) and bool((sb.get_c_terminal_mask().sum(1) == 2).all())
    seq_idx = sb.get_seq_idx()
    assert tuple(seq_idx.shape) == (2, 13) and seq_idx[0, 4] == 20 and seq_idx[0, 0] == 0
    single = ps.StructureBatch.from_pdb(path, device="cpu")
    assert single.get_batch_size() == 1 and int(single.get_total_lengths()[0]) == 13


@pytest.mark.skipif(not os.path.isdir("/root/reference/tests"), reason="reference tree not present on this box")
def test_native_pdb_ingest_matches_the_python_restatement_on_the_reference_files(native_lib):
    """All PDB files shipped with the reference; lengths pinned by its tests (437 / 130 / 184 / 229)."""
    import glob
    from protstruc_b200 import pdb_ingest
    from oracle import pdb_fixture_reader as reader

    files = sorted(glob.glob("/root/reference/tests/*.pdb")) + sorted(glob.glob("/root/reference/docs/tutorials/*.pdb"))
    assert len(files) >= 10
    pins = {"6dc4.pdb": 437, "1REX.pdb": 130, "4EOT.pdb": 184, "15c8_HL.pdb": 229, "1a6v_HL.pdb": 229}
    import json
    text_pins = json.loads((H.GOLDEN / "pdb_pins.json").read_text())
    assert len(text_pins) == len(files)
    for f in files:
        a = pdb_ingest.read_pdb_arrays(f)
        x, m, c, ids = reader.read_structure(f)
        assert np.array_equal(a["atom_mask"], m) and np.array_equal(a["chain_idx"], c) and a["chain_ids"] == ids, f
        assert np.array_equal(np.nan_to_num(a["xyz"], nan=-9.0), np.nan_to_num(x, nan=-9.0)), f
        name = os.path.basename(f)
        if name in pins:
            assert a["xyz"].shape[0] == pins[name], f
        # text-level pins from a third, independent code path (tests/golden/make_pdb_pins.py): residues after gap
        # filling, heavy atoms that survive the filters, the sum of their coordinates, chain order
        pin = text_pins[f"{os.path.basename(os.path.dirname(f))}/{name}"]
        assert a["xyz"].shape[0] == pin["residues_with_gap_fill"], f
        assert int(a["atom_mask"].sum()) == pin["heavy_atoms"], f
        assert int(a["atom_mask"].any(axis=1).sum()) == pin["residues_with_atoms"], f
        assert abs(float(np.nansum(a["xyz"].astype(np.float64))) - pin["coordinate_sum"]) < 0.05, f
        assert "".join(a["chain_ids"]) == pin["chains"], f
    batch = ps.StructureBatch.from_pdb(["/root/reference/tests/15c8_HL.pdb", "/root/reference/tests/1ad0_DC.pdb",
                                        "/root/reference/tests/5cjx_HL.pdb"], device="cpu")
    assert len(batch.get_xyz()) == 3  # reference tests/test_StructureBatch.py:56-65
    assert bool((batch.get_n_terminal_mask().sum(axis=1) == 2).all())
    assert bool((batch.get_c_terminal_mask().sum(axis=1) == 2).all())
    g = H.load_golden("real_1a6v_HL")
    one = ps.StructureBatch.from_pdb("/root/reference/tests/1a6v_HL.pdb", device="cpu")
    assert np.array_equal(one.get_atom_mask().numpy(), g["atom_mask"])


def test_empty_batches_give_empty_features_without_a_launch():
    """No structures / no residues: shapes and dtypes of the reference's (empty) results, nothing is launched,
    so this works without a GPU."""
    for B, L in ((0, 10), (3, 0)):
        sb = ps.StructureBatch.from_xyz(torch.zeros(B, L, 15, 3), torch.zeros(B, L, 15, dtype=torch.bool), device="cpu")
        dist, dist_mask = sb.pairwise_distance_matrix()
        assert tuple(dist.shape) == (B, L, L, 15, 15) and dist.dtype == torch.float32 and dist_mask.dtype == torch.bool
        feats = sb.inter_residue_geometry()
        assert tuple(feats["omega"].shape) == (B, L, L) and feats["d_ca_mask"].dtype == torch.bool
        dih, dmask = sb.backbone_dihedrals()
        assert tuple(dih.shape) == (B, L, 3) and dmask.dtype == torch.bool
        assert tuple(sb.backbone_orientations().shape) == (B, L, 3, 3)
        assert tuple(sb.pairwise_dihedrals(["CA", "CB"], ["CA", "CB"]).shape) == (B, L, L)
        assert tuple(sb.get_local_xyz().shape) == (B, L, 15, 3)
        com = sb.center_of_mass()
        assert tuple(com.shape) == (B, 3) and bool(torch.isnan(com).all())


def test_numa_binding_helper_is_a_no_op_without_a_gpu_and_never_raises():
    """`bind_host_thread_near_gpu` is opt-in plumbing for multi-rank host streaming: without CUDA / NVML (this
    container) it must leave the affinity alone and return None."""
    import os

    from protstruc_b200.host_pipeline import bind_host_thread_near_gpu

    before = os.sched_getaffinity(0)
    assert bind_host_thread_near_gpu(0) is None or torch.cuda.is_available()
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
