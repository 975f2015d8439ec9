"""GPU suite: the CUDA path (through the façade -> C-ABI -> sm_100a kernels) against
(a) the golden vectors produced by the real reference and (b) the CPU oracle on seeded inputs.

Tolerances are BASELINE.json's north_star: distances <= 1e-5 relative / 1e-4 A, angles <= 1e-5 rad
away from collinear degeneracies (circular comparison), masks / NaN placement / indexing bit-exact,
diffuse_xyz bit-exact given the injected noise.
"""
import math

import numpy as np
import pytest
import torch

import protstruc_b200 as ps
from protstruc_b200 import _cabi
from oracle import feature_oracle as orc
from tests import helpers as H

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device (B200); there is no CPU path")]

SYNTHETIC = ["synthetic_small", "synthetic_floatmask_oddL", "synthetic_A5",
             "synthetic_A25_like_reference_tests", "synthetic_ragged_33"]
DEV = "cuda"


def make_batch(g):
    ids = [["A", "B"] for _ in range(g["xyz"].shape[0])]
    return ps.StructureBatch.from_xyz(g["xyz"], g["atom_mask"], g["chain_idx"], ids)


angle_conditioning = H.trrosetta_conditioning


# ------------------------------------------------------------------------------ K1 distances + mask
@pytest.mark.parametrize("name", SYNTHETIC)
def test_pairwise_distance_matrix_matches_reference_golden(native_lib, name):
    g = H.load_golden(name)
    sb = make_batch(g)
    dist, dist_mask = sb.pairwise_distance_matrix()
    ref_dist, ref_mask = H.t(g["ref_dist"]), H.t(g["ref_dist_mask"])
    assert dist.is_cuda and dist.dtype == torch.float32 and dist.is_contiguous()
    assert tuple(dist.shape) == tuple(ref_dist.shape)
    assert dist_mask.dtype == ref_mask.dtype and tuple(dist_mask.shape) == tuple(ref_mask.shape)
    H.assert_distances_close(dist, ref_dist)
    assert torch.equal(dist_mask.cpu(), ref_mask), "pair mask must be bit-exact"


def test_pairwise_distance_matrix_real_structure(native_lib):
    g = H.load_golden("real_1a6v_HL")
    sb = make_batch(g)
    dist, dist_mask = sb.pairwise_distance_matrix()
    assert tuple(dist.shape) == (1, 229, 229, 15, 15)
    H.assert_distances_close(dist[:, :40, :40], H.t(g["ref_dist_crop"]), "dist crop")
    assert torch.equal(dist_mask[:, :40, :40].cpu(), H.t(g["ref_dist_mask_crop"]))
    H.assert_distances_close(dist[:, :, :, 1, 1], H.t(g["ref_d_ca"]), "d_ca")
    H.assert_distances_close(dist[:, :, :, 4, 4], H.t(g["ref_d_cb"]), "d_cb")
    H.assert_distances_close(dist[:, :, :, 0, 3], H.t(g["ref_d_no"]), "d_no")
    assert int(torch.isnan(dist).sum()) == int(g["ref_dist_nan_count"])
    assert int(dist_mask.sum()) == int(g["ref_dist_mask_sum"])
    got = float(torch.nansum(dist.double()))
    assert math.isclose(got, float(g["ref_dist_nansum"]), rel_tol=1e-6)
    # reference tests/test_StructureBatch.py:122-137: CA-CA >= 0 and ATOM.CA indexes slot 1
    ca = dist[:, :, :, ps.ATOM.CA, ps.ATOM.CA]
    assert bool((ca[~torch.isnan(ca)] >= 0).all())


@pytest.mark.parametrize("B,L,A,kind", [(3, 37, 15, "bool"), (1, 1, 15, "bool"), (2, 2, 15, "float"),
                                        (1, 130, 15, "bool"), (2, 19, 10, "bool"), (1, 6, 37, "float"),
                                        (5, 16, 15, "bool"), (2, 40, 5, "bool"), (2, 33, 10, "float"),
                                        (1, 64, 14, "bool"), (3, 35, 5, "float"), (2, 32, 16, "bool"),
                                        (2, 128, 5, "bool"), (1, 131, 5, "float"), (1, 64, 10, "bool"),
                                        (2, 67, 10, "float"), (3, 150, 5, "bool")])
def test_pairwise_distance_matrix_vs_oracle(native_lib, B, L, A, kind):
    xyz, mask, chain_idx = H.synthetic_batch(100 + L, B, L, A, kind)
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    dist, dist_mask = sb.pairwise_distance_matrix()
    ref_dist, ref_mask = orc.pair_distances(xyz, mask)
    H.assert_distances_close(dist, ref_dist)
    assert dist_mask.dtype == ref_mask.dtype
    assert torch.equal(dist_mask.cpu(), ref_mask)


def test_distance_kernel_variants_agree(native_lib):
    """Staged (TMA-store) kernel vs generic per-element kernel: same arithmetic, bit-equal; the IEEE
    sqrt variant stays within 1 ulp of the MUFU variant."""
    xyz, mask, _ = H.synthetic_batch(7, 2, 45, 15, "bool", nan_masked=False)
    x = xyz.to(DEV)
    m = mask.to(DEV)
    B, L, A = 2, 45, 15
    outs = {}
    for variant in (0, 1, 2, 1 << 8, 2 << 4, 1 << 9):
        d = torch.empty(B, L, L, A, A, device=DEV)
        dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)
        rc = native_lib.ps_pair_dist_mask_ex(x.data_ptr(), m.data_ptr(), 0, d.data_ptr(), dm.data_ptr(), B, L, A,
                                             variant, torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "ps_pair_dist_mask_ex")
        outs[variant] = (d, dm)
    torch.cuda.synchronize()
    base, base_mask = outs[0]
    assert torch.equal(outs[1 << 8][0], base), "generic kernel differs from the staged kernel"
    assert torch.equal(outs[2 << 4][0], base), "tile-buffers-per-CTA override changed the result"
    assert torch.equal(outs[1 << 9][0], base), "warps-per-tile variant changed the result"
    for v in outs:
        assert torch.equal(outs[v][1], base_mask)
    rel = ((outs[2][0] - base).abs() / outs[2][0].clamp_min(1e-30)).max().item()
    assert rel <= 2.5e-7, f"MUFU sqrt deviates from IEEE sqrt by {rel}"
    assert torch.equal(outs[1][0], base), "non-ftz variant differs on normal-range inputs"


@pytest.mark.parametrize("B,L,A,shift,codes", [
    (3, 96, 15, 16, (0, 1, 2, 3, 4, 7)),   # linear-sweep kernel: default (evict_first), explicit policies, 7 = none
    (3, 96, 10, 16, (0, 1, 2, 3)),         # column-strip kernel of the 10-atom layout: default on, 3 = none
    (2, 128, 5, 16, (0, 1, 2, 3)),         # 5-atom layout (128 pairs per tile): default off
    (2, 64, 25, 28, (0, 1, 2, 3)),         # any-A tile kernel (unrolled 25-atom instantiation)
    (2, 40, 20, 28, (0, 1, 2, 3)),         # any-A tile kernel, run-time atom count
])
def test_l2_eviction_policy_of_the_tile_stores_never_changes_a_byte(native_lib, B, L, A, shift, codes):
    """The L2 cache hint on the bulk tile stores (DESIGN K1s, `tools/l2_hint_probe.py`) is a performance knob only:
    every policy — and the launcher's per-kind default — writes the same distances and masks, with and without the
    fused angles."""
    xyz, mask, _ = H.synthetic_batch(90 + A, B, L, A, "bool")
    x, m = xyz.to(DEV), mask.to(DEV)
    s = torch.cuda.current_stream().cuda_stream
    base = None
    for code in codes:
        d = torch.full((B, L, L, A, A), -7.0, device=DEV)
        dm = torch.zeros(B, L, L, A, A, dtype=torch.bool, device=DEV)
        _cabi.check(native_lib.ps_pair_dist_mask_ex(x.data_ptr(), m.data_ptr(), 0, d.data_ptr(), dm.data_ptr(), B, L, A,
                                                    code << shift, s), "ps_pair_dist_mask_ex")
        got = [d.view(torch.int32), dm]
        if A in (15, 10, 5) and shift == 16:  # fused launch on the staged layouts
            om, th, ph = (torch.full((B, L, L), -7.0, device=DEV) for _ in range(3))
            d2 = torch.full((B, L, L, A, A), -7.0, device=DEV)
            dm2 = torch.zeros(B, L, L, A, A, dtype=torch.bool, device=DEV)
            fused_code = code if A != 15 else {0: 0, 1: 1, 2: 6, 3: 5, 4: 4, 7: 7}[code]  # sweep: also the angle-plane hooks
            _cabi.check(native_lib.ps_inter_residue_geometry_ex(x.data_ptr(), m.data_ptr(), 0, d2.data_ptr(), dm2.data_ptr(),
                                                                om.data_ptr(), th.data_ptr(), ph.data_ptr(), B, L, A,
                                                                fused_code << shift, s), "ps_inter_residue_geometry_ex")
            got += [d2.view(torch.int32), dm2, om.view(torch.int32), th.view(torch.int32), ph.view(torch.int32)]
        torch.cuda.synchronize()
        if base is None:
            base = got
        else:
            for k, (a, b) in enumerate(zip(got, base)):
                assert torch.equal(a, b), f"policy code {code} changed output {k}"


@pytest.mark.parametrize("B,L,A,nan_masked", [(2, 140, 5, False), (2, 140, 5, True), (2, 70, 10, True), (1, 192, 10, True),
                                              (1, 37, 14, True), (2, 64, 15, True)])
def test_fused_and_split_dispatch_of_inter_residue_geometry_give_the_same_bits(native_lib, B, L, A, nan_masked):
    """The launcher splits the fused call into distance tiles + the exact-sequence angle kernel for the 5- and 10-atom
    layouts from 32 k pairs (angle-bound when fused).  Fused (variant bit 19), split (bit 20) and the default write
    identical bytes."""
    xyz, mask, _ = H.synthetic_batch(700 + A, B, L, A, "bool", nan_masked=nan_masked)
    x, m = xyz.to(DEV), mask.to(DEV)
    s = torch.cuda.current_stream().cuda_stream
    outs, launches = {}, {}
    for name, variant in (("default", 0), ("fused", 1 << 19), ("split", 1 << 20)):
        if A == 15:
            variant |= 1 << 15  # the column-strip kernel (the sweep kernel has its own fused path)
        d = torch.full((B, L, L, A, A), -7.0, device=DEV)
        dm = torch.zeros(B, L, L, A, A, dtype=torch.bool, device=DEV)
        om, th, ph = (torch.full((B, L, L), -7.0, device=DEV) for _ in range(3))
        _cabi.check(native_lib.ps_inter_residue_geometry_ex(x.data_ptr(), m.data_ptr(), 0, d.data_ptr(), dm.data_ptr(),
                                                            om.data_ptr(), th.data_ptr(), ph.data_ptr(), B, L, A, variant, s),
                    "ps_inter_residue_geometry_ex")
        torch.cuda.synchronize()
        launches[name] = _cabi.last_pair_dist_plan()["launches"]
        outs[name] = [t.view(torch.int32) if t.dtype == torch.float32 else t for t in (d, dm, om, th, ph)]
    assert launches["fused"] == 1 and launches["split"] == 2
    assert launches["default"] == (2 if A in (5, 10) and B * L * L >= 32768 else 1)  # small calls stay one launch
    for name in ("fused", "split"):
        for k, (a, b) in enumerate(zip(outs[name], outs["default"])):
            assert torch.equal(a, b), f"{name} dispatch differs from the default in output {k}"


@pytest.mark.parametrize("B,L,A", [(3, 140, 5), (2, 129, 5), (3, 128, 4), (40, 256, 5)])
def test_small_layouts_with_eight_tile_buffers_write_the_same_bytes(native_lib, B, L, A):
    """The 4- / 5-atom strip kernels run eight tile buffers per CTA (16 warps, 512 threads) on large calls and four on
    small ones: forced 8, forced 4, 6, the default and the any-A kernel write identical distances and masks — also for
    the fused launch of the 5-atom layout (variant bit 19)."""
    xyz, mask, _ = H.synthetic_batch(800 + A + L, B, L, A, "bool")
    x, m = xyz.to(DEV), mask.to(DEV)
    s = torch.cuda.current_stream().cuda_stream
    outs = {}
    for name, variant in (("default", 0), ("s8", 8 << 4), ("s4", 4 << 4), ("s6", 6 << 4), ("anyA", 1 << 8)):
        d = torch.full((B, L, L, A, A), -7.0, device=DEV)
        dm = torch.zeros(B, L, L, A, A, dtype=torch.bool, device=DEV)
        _cabi.check(native_lib.ps_pair_dist_mask_ex(x.data_ptr(), m.data_ptr(), 0, d.data_ptr(), dm.data_ptr(), B, L, A,
                                                    variant, s), "ps_pair_dist_mask_ex")
        torch.cuda.synchronize()
        if name != "anyA":
            plan = _cabi.last_pair_dist_plan()
            assert plan["path"] == 0 and plan["sweep"] == 0
            if name in ("s8", "s4", "s6"):
                assert plan["tile_buffers"] == plan["ctas"] * int(name[1:])
        outs[name] = (d.view(torch.int32), dm)
    if B * L * L // 128 >= 8 * 148:  # a large call: the default IS eight buffers
        _cabi.check(native_lib.ps_pair_dist_mask_ex(x.data_ptr(), m.data_ptr(), 0, outs["default"][0].view(torch.float32).data_ptr(),
                                                    outs["default"][1].data_ptr(), B, L, A, 0, s), "ps_pair_dist_mask_ex")
        assert _cabi.last_pair_dist_plan()["tile_buffers"] == 8 * _cabi.last_pair_dist_plan()["ctas"]
    for name in ("s8", "s4", "s6", "anyA"):
        assert torch.equal(outs[name][0], outs["default"][0]), f"{name}: distances differ"
        assert torch.equal(outs[name][1], outs["default"][1]), f"{name}: mask differs"
    rd, rm = orc.pair_distances(xyz[:2], mask[:2])
    H.assert_distances_close(outs["s8"][0].view(torch.float32)[:2], rd)
    assert torch.equal(outs["s8"][1][:2].cpu(), rm)
    if A == 5:
        fused = {}
        for name, variant in (("s4", (1 << 19) | (4 << 4)), ("s8", (1 << 19) | (8 << 4))):
            d = torch.full((B, L, L, A, A), -7.0, device=DEV)
            dm = torch.zeros(B, L, L, A, A, dtype=torch.bool, device=DEV)
            om, th, ph = (torch.full((B, L, L), -7.0, device=DEV) for _ in range(3))
            _cabi.check(native_lib.ps_inter_residue_geometry_ex(x.data_ptr(), m.data_ptr(), 0, d.data_ptr(), dm.data_ptr(),
                                                                om.data_ptr(), th.data_ptr(), ph.data_ptr(), B, L, A,
                                                                variant, s), "ps_inter_residue_geometry_ex")
            torch.cuda.synchronize()
            fused[name] = [t.view(torch.int32) for t in (d, om, th, ph)] + [dm]
        for k, (a, b) in enumerate(zip(fused["s8"], fused["s4"])):
            assert torch.equal(a, b), f"fused launch with eight buffers differs in output {k}"
        assert torch.equal(fused["s8"][0], outs["default"][0])


@pytest.mark.parametrize("B,L", [(8, 256), (6, 250), (5, 190), (3, 384)])
def test_tile_schedules_write_the_same_bytes(native_lib, B, L):
    """The linear-sweep kernel (default at A = 15), and the column-strip kernel with its cell schedule, its lock-step
    schedule and the relaxed flavour of that, only differ in WHICH tile buffer writes a tile, when, and how residue j
    reaches the lane: every output byte must be identical, for the distance + mask kernel and for the fused kernel,
    including the angle tensors."""
    A = 15
    xyz, mask, _ = H.synthetic_batch(600 + L, B, L, A, "bool")
    x, m = xyz.to(DEV), mask.to(DEV)
    s = torch.cuda.current_stream().cuda_stream
    outs = []
    strip = 1 << 15
    for variant in (1 << 27, strip, strip | (1 << 13), strip | (1 << 11), strip | (1 << 14)):
        d = torch.full((B, L, L, A, A), -1.0, device=DEV)
        dm = torch.zeros(B, L, L, A, A, dtype=torch.bool, device=DEV)
        _cabi.check(native_lib.ps_pair_dist_mask_ex(x.data_ptr(), m.data_ptr(), 0, d.data_ptr(), dm.data_ptr(), B, L, A,
                                                    variant, s), "ps_pair_dist_mask_ex")
        fd = torch.full((B, L, L, A, A), -1.0, device=DEV)
        fm = torch.zeros(B, L, L, A, A, dtype=torch.bool, device=DEV)
        angles = [torch.full((B, L, L), -9.0, device=DEV) for _ in range(3)]
        _cabi.check(native_lib.ps_inter_residue_geometry_ex(x.data_ptr(), m.data_ptr(), 0, fd.data_ptr(), fm.data_ptr(),
                                                            angles[0].data_ptr(), angles[1].data_ptr(),
                                                            angles[2].data_ptr(), B, L, A, variant, s), "fused")
        outs.append([d, dm, fd, fm] + angles)
    torch.cuda.synchronize()
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert torch.equal(torch.nan_to_num(a.float(), nan=-5.0), torch.nan_to_num(b.float(), nan=-5.0))
    assert torch.equal(torch.nan_to_num(outs[0][0], nan=-5.0), torch.nan_to_num(outs[0][2], nan=-5.0))
    rd, rm = orc.pair_distances(xyz, mask)
    H.assert_distances_close(outs[0][0], rd)
    assert torch.equal(outs[0][1].cpu(), rm)


def test_distance_properties_at_baseline_config2(native_lib):
    """BASELINE config 2 (64 x 256 x 15, 3.8 GB of distances): size-independent properties."""
    B, L, A = 64, 256, 15
    g = torch.Generator(device=DEV).manual_seed(2)
    xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
    mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    dist, dist_mask = sb.pairwise_distance_matrix()
    # symmetry: dist[b,i,j,a,c] == dist[b,j,i,c,a], bit for bit
    for b in (0, 31, 63):
        d = dist[b]
        assert torch.equal(d, d.permute(1, 0, 3, 2)), "distance tensor is not symmetric"
        assert bool((torch.diagonal(torch.diagonal(d, dim1=0, dim2=1), dim1=0, dim2=1) == 0).all())
    # mask identity: sum_{i,j,a,c} m_i[a] m_j[c] = (sum m)^2 per structure
    per_struct = dist_mask.view(B, -1).sum(dim=1, dtype=torch.int64)
    assert torch.equal(per_struct, mask.view(B, -1).sum(dim=1, dtype=torch.int64) ** 2)
    # spot check against the oracle on random structures / residue blocks
    for b, i0, j0 in ((0, 0, 0), (17, 100, 200), (63, 250, 3)):
        sub_i = xyz[b, i0:i0 + 6].cpu()
        sub_j = xyz[b, j0:j0 + 6].cpu()
        ref = torch.norm(sub_i[:, None, :, None] - sub_j[None, :, None, :], dim=-1)
        H.assert_distances_close(dist[b, i0:i0 + 6, j0:j0 + 6], ref, "spot block")


@pytest.mark.parametrize("L,expect_lockstep", [(512, True), (384, None), (256, None)])
def test_full_feature_set_at_the_baseline_shapes_vs_oracle(native_lib, L, expect_lockstep):
    """The shapes the numbers of record are quoted on — L = 512 (the metric shape, bench.py), L = 384 (BASELINE
    config 5), L = 256 (config 2); A = 15, bool mask, NaN-masked coordinates, ragged lengths, DEFAULT dispatch through
    the public call — against the CPU oracle (reference protstruc.py:790-817) on the FULL tensors: every distance,
    every mask byte, omega / theta / phi.  The launch plan is read back so that the test fails if the shape stops
    taking the fused staged kernel (and, at L = 512, the lock-step schedule the bench runs)."""
    B, A = 2, 15
    xyz, mask, _ = H.synthetic_batch(5120 + L, B, L, A, "bool")
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    out = sb.inter_residue_geometry()
    plan = _cabi.last_pair_dist_plan()
    assert plan["path"] == 0 and plan["launches"] == 1, plan  # ONE fused launch of the staged tile kernel
    if expect_lockstep is not None:
        assert bool(plan["lockstep"]) == expect_lockstep, plan
    dist, dist_mask = out["d_ca"]._base, out["d_ca_mask"]._base  # the full tensors behind the strided views
    assert tuple(dist.shape) == (B, L, L, A, A) and dist_mask.dtype == torch.bool
    for b in range(B):  # one structure at a time: the oracle needs ~1.5 GB per 512-residue structure
        ref = orc.inter_residue_geometry(xyz[b:b + 1], mask[b:b + 1])
        rd, rm = orc.pair_distances(xyz[b:b + 1], mask[b:b + 1])
        H.assert_distances_close(dist[b:b + 1], rd, f"dist[{b}]")
        assert torch.equal(dist_mask[b:b + 1].cpu(), rm), f"dist_mask[{b}]"
        del rd, rm
        for which in ("omega", "theta"):
            H.assert_angles_close(out[which][b:b + 1], ref[which], angle_conditioning(xyz[b:b + 1], which),
                                  f"{which}[{b}]", all_finite_tol=2e-6)
        H.assert_angles_close(out["phi"][b:b + 1], ref["phi"], angle_conditioning(xyz[b:b + 1], "phi"), f"phi[{b}]",
                              circular=False)
        for key in ("d_ca", "d_cb", "d_no"):
            H.assert_distances_close(out[key][b:b + 1], ref[key], key)
            assert torch.equal(out[key + "_mask"][b:b + 1].cpu(), ref[key + "_mask"])
    # the separate distance call at the same shape (its own default schedule) writes the same bytes
    d2, m2 = sb.pairwise_distance_matrix()
    assert torch.equal(torch.nan_to_num(d2, nan=-5.0), torch.nan_to_num(dist, nan=-5.0)) and torch.equal(m2, dist_mask)


def collinear_batch(L: int = 192):
    """Adversarial input for the unclamped arccos of phi = angle(CA_i, CB_i, CB_j) (reference geometry.py:64-71):
    every CA and CB of a structure lies on ONE line, so every triple is exactly or nearly collinear and the rounded
    cosine lands on either side of +-1 — whether an entry is NaN depends on the last ulp of the reference's op
    sequence.  Structures: axis-aligned line (exactly collinear in fp32), general direction, the same scaled by 100
    and by 0.01, one-ulp perturbations of the axis-aligned case, and a general line far from the origin."""
    g = torch.Generator().manual_seed(777)
    structs = []
    for kind in ("axis", "general", "big", "small", "ulp", "offset"):
        d = torch.randn(3, generator=g)
        if kind in ("axis", "ulp"):
            d = torch.tensor([0.0, 1.0, 0.0])
        p0 = 10.0 * torch.randn(3, generator=g) if kind != "offset" else 500.0 + 10.0 * torch.randn(3, generator=g)
        s_ca = 40.0 * torch.rand(L, generator=g) - 20.0
        s_cb = s_ca + (0.5 + 2.5 * torch.rand(L, generator=g)) * torch.where(torch.rand(L, generator=g) < 0.5, -1.0, 1.0)
        x = 1.5 * torch.randn(L, 15, 3, generator=g) + p0
        x[:, 1] = p0 + s_ca[:, None] * d
        x[:, 4] = p0 + s_cb[:, None] * d
        if kind == "big":
            x = x * 100.0
        if kind == "small":
            x = x * 0.01
        if kind == "ulp":
            bump = torch.rand(L, 3, generator=g) < 0.3
            x[:, 4] = torch.where(bump, torch.nextafter(x[:, 4], torch.full_like(x[:, 4], float("inf"))), x[:, 4])
        structs.append(x)
    xyz = torch.stack(structs).contiguous()
    return xyz, torch.ones(xyz.shape[:3], dtype=torch.bool)


def test_phi_nan_placement_on_collinear_triples_is_the_references(native_lib):
    """NaN placement of phi is decided by the last ulp of cos = (ba.bc) / (|ba| |bc|): the fused kernels must issue
    the reference's exact sequence (ATen norm = FMA chain, rounded product, IEEE division) wherever |cos| is near 1.
    Checked for the fused K1 (inter_residue_geometry), K2f (trrosetta_angles) and the generic planar kernel against
    the CPU oracle on ~220 k exactly / nearly collinear triples — the NaN maps must be IDENTICAL and the finite
    values within 1e-3 rad (arccos near +-1 amplifies one ulp of the cosine to 3.5e-4 rad)."""
    xyz, mask = collinear_batch()
    ref = torch.cat([orc.pair_planar_angles(xyz[b:b + 1], [1, 4], [4]) for b in range(xyz.shape[0])])
    n_nan = int(torch.isnan(ref).sum())
    off_diag = xyz.shape[0] * xyz.shape[1] * (xyz.shape[1] - 1)
    assert 0.05 * off_diag < n_nan - xyz.shape[0] * xyz.shape[1] < 0.95 * off_diag, "the input is not adversarial"
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    fused = sb.inter_residue_geometry()
    assert _cabi.last_pair_dist_plan()["path"] == 0
    k2f = sb.trrosetta_angles()
    generic = sb.pairwise_planar_angles(["CA", "CB"], ["CB"])
    for got, name in ((fused["phi"], "fused K1"), (k2f[2], "K2f"), (generic, "generic planar kernel")):
        H.assert_same_nan(got, ref, f"phi ({name})")
        ok = ~torch.isnan(ref)
        assert (got.cpu()[ok] - ref[ok]).abs().max().item() < 1e-3, name
    # omega / theta of the same batch are degenerate everywhere (all four points on a line / n2 = 0): only the NaN
    # maps are comparable
    ro, rt, _ = [torch.cat(t) for t in zip(*[orc.trrosetta_angles(xyz[b:b + 1]) for b in range(xyz.shape[0])])]
    H.assert_same_nan(fused["omega"], ro, "omega")
    H.assert_same_nan(fused["theta"], rt, "theta")
    H.assert_same_nan(k2f[0], ro, "omega K2f")
    H.assert_same_nan(k2f[1], rt, "theta K2f")


# ------------------------------------------------------------------------------ K2 pairwise angles
@pytest.mark.parametrize("name", ["synthetic_small", "synthetic_floatmask_oddL", "synthetic_A5", "synthetic_ragged_33",
                                  "real_1a6v_HL"])
def test_inter_residue_geometry_matches_reference_golden(native_lib, name):
    g = H.load_golden(name)
    sb = make_batch(g)
    out = sb.inter_residue_geometry()
    xyz = H.t(g["xyz"])
    assert set(out) == {"d_ca", "d_ca_mask", "d_cb", "d_cb_mask", "d_no", "d_no_mask", "omega", "theta", "phi"}
    for which in ("omega", "theta"):
        H.assert_angles_close(out[which], H.t(g[f"ref_{which}"]), angle_conditioning(xyz, which), which,
                              all_finite_tol=2e-6)
    H.assert_angles_close(out["phi"], H.t(g["ref_phi"]), angle_conditioning(xyz, "phi"), "phi", circular=False)
    # diagonal conventions: omega = theta = 0, phi = NaN where the residue is complete
    L = xyz.shape[1]
    eye = torch.eye(L, dtype=torch.bool)
    ref_omega = H.t(g["ref_omega"])
    same_zero = (out["omega"].cpu()[:, eye] == 0) == (ref_omega[:, eye] == 0)
    assert bool(same_zero.all())
    if name == "real_1a6v_HL":
        H.assert_distances_close(out["d_ca"], H.t(g["ref_d_ca"]), "d_ca")
        H.assert_distances_close(out["d_no"], H.t(g["ref_d_no"]), "d_no")
        assert torch.equal(out["d_ca_mask"].cpu(), H.t(g["ref_d_ca_mask"]))
    else:
        ref_dist, ref_mask = H.t(g["ref_dist"]), H.t(g["ref_dist_mask"])
        H.assert_distances_close(out["d_cb"], ref_dist[:, :, :, 4, 4], "d_cb")
        assert torch.equal(out["d_no_mask"].cpu(), ref_mask[:, :, :, 0, 3])


def pair_angles_ex(lib, sb, slots_i, slots_j, kind, variant):
    """ps_pair_angles_ex straight through the C-ABI (kind 0 = dihedral, 1 = planar; variant 1 = exact-sequence kernel)."""
    x = sb.get_xyz()
    B, L, A = x.shape[:3]
    out = torch.empty(B, L, L, device=DEV)
    _cabi.check(lib.ps_pair_angles_ex(x.data_ptr(), B, L, A, _cabi.int_array(slots_i), len(slots_i), _cabi.int_array(slots_j),
                                      len(slots_j), kind, out.data_ptr(), variant, torch.cuda.current_stream().cuda_stream),
                "ps_pair_angles_ex")
    return out


@pytest.mark.parametrize("name", ["synthetic_small", "synthetic_A5", "real_1a6v_HL"])
def test_pairwise_angle_methods_match_reference_golden(native_lib, name):
    g = H.load_golden(name)
    sb = make_batch(g)
    xyz = H.t(g["xyz"])
    # the public methods (default: the packed generic kernel): the contract — NaN placement of the reference,
    # <= 1e-5 rad where min sin(bond angle) >= 0.1, <= 1e-6 / sin below
    pomega = sb.pairwise_dihedrals(["CA", "CB"], ["CA", "CB"])
    ptheta = sb.pairwise_dihedrals(["N", "CA", "CB"], ["CB"])
    pphi = sb.pairwise_planar_angles(["CA", "CB"], ["CB"])
    assert tuple(pomega.shape) == tuple(g["ref_omega"].shape) and pomega.dtype == torch.float32
    H.assert_angles_close(pomega, H.t(g["ref_omega"]), angle_conditioning(xyz, "omega"), "omega (packed generic)")
    H.assert_angles_close(ptheta, H.t(g["ref_theta"]), angle_conditioning(xyz, "theta"), "theta (packed generic)")
    H.assert_angles_close(pphi, H.t(g["ref_phi"]), angle_conditioning(xyz, "phi"), "phi (packed generic)", circular=False)
    eye_l = torch.eye(xyz.shape[1], dtype=torch.bool)
    for got, ref in ((pomega, g["ref_omega"]), (ptheta, g["ref_theta"]), (pphi, g["ref_phi"])):  # exact on the diagonal
        assert torch.equal(torch.nan_to_num(got.cpu()[:, eye_l], nan=-9.0).abs(), torch.nan_to_num(H.t(ref)[:, eye_l], nan=-9.0).abs())
    # the exact-sequence generic kernel (ps_pair_angles_ex variant 1): a few ulp on EVERY finite entry
    omega = pair_angles_ex(native_lib, sb, [1, 4], [1, 4], 0, 1)
    theta = pair_angles_ex(native_lib, sb, [0, 1, 4], [4], 0, 1)
    phi = pair_angles_ex(native_lib, sb, [1, 4], [4], 1, 1)
    H.assert_angles_close(omega, H.t(g["ref_omega"]), angle_conditioning(xyz, "omega"), "omega", all_finite_tol=2e-6)
    H.assert_angles_close(theta, H.t(g["ref_theta"]), angle_conditioning(xyz, "theta"), "theta", all_finite_tol=2e-6)
    H.assert_angles_close(phi, H.t(g["ref_phi"]), angle_conditioning(xyz, "phi"), "phi", circular=False)
    # K2f, exact-sequence variant: shares the geometry with the generic kernel and finishes with tuned scalar steps
    # (rsqrt / polynomial atan2): same NaN map, values within a few ulp of the straightforward IEEE kernel
    B, L, A = xyz.shape[:3]
    s = torch.cuda.current_stream().cuda_stream
    eo, et, ep = (torch.empty(B, L, L, device=DEV) for _ in range(3))
    _cabi.check(native_lib.ps_trrosetta_angles_ex(sb.get_xyz().data_ptr(), B, L, A, 0, eo.data_ptr(), et.data_ptr(),
                                                  ep.data_ptr(), 1, s), "ps_trrosetta_angles_ex")
    for a, b, tol in ((eo, omega, 1e-6), (et, theta, 1e-6)):
        assert torch.equal(torch.isnan(a), torch.isnan(b))
        assert H.circular_diff(torch.nan_to_num(a).cpu(), torch.nan_to_num(b).cpu()).max().item() <= tol
    assert torch.equal(torch.isnan(ep), torch.isnan(phi))
    H.assert_angles_close(ep, phi.cpu(), angle_conditioning(xyz, "phi"), "phi exact K2f vs generic", circular=False)
    # K2f, default (packed FP32, fused multiply-adds): the contract — NaN placement bit-exact against the REFERENCE,
    # <= 1e-5 rad where min sin(bond angle) >= 0.1, <= 1e-6 / sin below; diagonal and zero-padded entries exact
    fo, ft, fp = sb.trrosetta_angles()
    H.assert_angles_close(fo, H.t(g["ref_omega"]), angle_conditioning(xyz, "omega"), "omega (packed K2f)")
    H.assert_angles_close(ft, H.t(g["ref_theta"]), angle_conditioning(xyz, "theta"), "theta (packed K2f)")
    H.assert_angles_close(fp, H.t(g["ref_phi"]), angle_conditioning(xyz, "phi"), "phi (packed K2f)", circular=False)
    eye = torch.eye(L, dtype=torch.bool)
    for got, ref in ((fo, g["ref_omega"]), (ft, g["ref_theta"])):
        rd = H.t(ref)[:, eye]
        assert torch.equal(torch.nan_to_num(got.cpu()[:, eye], nan=-9.0), torch.nan_to_num(rd, nan=-9.0)), "diagonal"
    if name != "real_1a6v_HL":
        gen = sb.pairwise_dihedrals(["n", "ca", "c"], ["N"])  # case-insensitive names
        p = H.pair_points(xyz, [0, 1, 2], [0])
        cond = H.dihedral_conditioning(p[..., 0, :], p[..., 1, :], p[..., 2, :], p[..., 3, :])
        H.assert_angles_close(gen, H.t(g["ref_psi_like_dihedral_N_CA_C_N"]), cond, "generic dihedral")
        pl = sb.pairwise_planar_angles(["CA"], ["CA", "C"])
        p = H.pair_points(xyz, [1], [1, 2])
        H.assert_angles_close(pl, H.t(g["ref_planar_CA_CA_C"]), H.planar_conditioning(p[..., 0, :], p[..., 1, :], p[..., 2, :]),
                              "generic planar", circular=False)


def test_pairwise_angles_at_backbone_scale_vs_oracle(native_lib):
    """A slice of BASELINE config 3 (backbone N,CA,C,O,CB; L = 512) small enough for the CPU oracle."""
    xyz, mask, _ = H.synthetic_batch(3, 2, 512, 5, "bool", nan_masked=False, full_length=True)
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    omega, theta, phi = sb.trrosetta_angles()
    ro, rt, rp = orc.trrosetta_angles(xyz)
    H.assert_angles_close(omega, ro, angle_conditioning(xyz, "omega"), "omega")
    H.assert_angles_close(theta, rt, angle_conditioning(xyz, "theta"), "theta")
    H.assert_angles_close(phi, rp, angle_conditioning(xyz, "phi"), "phi", circular=False)
    assert bool((omega.abs() <= math.pi + 1e-6).all()) and bool((phi[~torch.isnan(phi)] >= 0).all())
    # the generic packed kernel (pairwise_dihedrals / pairwise_planar_angles, any slot lists) at the same size: every
    # split of the point list between the two residues, NaN-masked and ragged inputs included
    for nan_masked in (False, True):
        x2, m2, _ = H.synthetic_batch(5, 2, 512, 5, "bool", nan_masked=nan_masked, full_length=not nan_masked)
        sb2 = ps.StructureBatch.from_xyz(x2, m2)
        for si, sj in (([1, 4], [1, 4]), ([0, 1, 4], [4]), ([2], [0, 1, 2]), ([1], [4, 0, 3])):
            got = pair_angles_ex(native_lib, sb2, si, sj, 0, 0)
            p = H.pair_points(x2, si, sj)
            cond = H.dihedral_conditioning(p[..., 0, :], p[..., 1, :], p[..., 2, :], p[..., 3, :])
            H.assert_angles_close(got, orc.pair_dihedrals(x2, si, sj), cond, f"generic dihedral {si} {sj}")
        for si, sj in (([1, 4], [4]), ([1], [1, 2])):
            got = pair_angles_ex(native_lib, sb2, si, sj, 1, 0)
            p = H.pair_points(x2, si, sj)
            cond = H.planar_conditioning(p[..., 0, :], p[..., 1, :], p[..., 2, :])
            H.assert_angles_close(got, orc.pair_planar_angles(x2, si, sj), cond, f"generic planar {si} {sj}", circular=False)


def test_virtual_cb_option_vs_oracle(native_lib):
    xyz, mask, _ = H.synthetic_batch(21, 2, 40, 15, "bool", nan_masked=False, full_length=True)
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    omega, theta, phi = sb.trrosetta_angles(virtual_cb=True)
    ro, rt, rp = orc.trrosetta_angles_virtual_cb(xyz)
    x5 = xyz[:, :, :5].clone()
    x5[:, :, 4] = orc.virtual_cb(xyz[:, :, 0], xyz[:, :, 1], xyz[:, :, 2])
    H.assert_angles_close(omega, ro, angle_conditioning(x5, "omega"), "omega(virtual CB)")
    H.assert_angles_close(theta, rt, angle_conditioning(x5, "theta"), "theta(virtual CB)")
    H.assert_angles_close(phi, rp, angle_conditioning(x5, "phi"), "phi(virtual CB)", circular=False)


@pytest.mark.parametrize("B,L,A,kind", [(2, 96, 15, "bool"), (3, 45, 15, "bool"), (2, 40, 15, "float"), (2, 20, 25, "bool"),
                                        (2, 140, 5, "bool"), (1, 31, 15, "bool")])
def test_compact_feature_planes_equal_the_strided_views(native_lib, B, L, A, kind):
    """inter_residue_geometry_compact: the six dense (B, L, L) planes written by the fused launch (read back from the
    finished tile) are bit-identical to the reference-style outputs — the angle tensors and the strided d_ca / d_cb /
    d_no views of the full distance tensor — for the staged kernel, the fp32-mask path and the any-A fallback."""
    xyz, mask, _ = H.synthetic_batch(900 + L, B, L, A, kind)
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    ref = sb.inter_residue_geometry()
    guard = torch.full((6 * B * L * L + 64,), -77.0, device=DEV)
    buf = guard[32:32 + 6 * B * L * L].view(6, B, L, L)
    got = sb.inter_residue_geometry_compact(out=buf)
    assert got["compact"].data_ptr() == buf.data_ptr()
    for name in ("omega", "theta", "phi", "d_ca", "d_cb", "d_no"):
        assert got[name].is_contiguous() and tuple(got[name].shape) == (B, L, L)
        assert torch.equal(torch.nan_to_num(got[name], nan=-5.0), torch.nan_to_num(ref[name], nan=-5.0)), name
    assert torch.equal(torch.nan_to_num(got["dist"], nan=-5.0), torch.nan_to_num(ref["d_ca"]._base, nan=-5.0))
    assert torch.equal(got["dist_mask"], ref["d_ca_mask"]._base)
    assert bool((guard[:32] == -77.0).all()) and bool((guard[-32:] == -77.0).all())


@pytest.mark.parametrize("B,L,kind", [(3, 64, "bool"), (2, 45, "bool"), (2, 40, "float")])
def test_fused_gather_push_writes_every_peer_buffer(native_lib, B, L, kind):
    """ps_inter_residue_geometry_push with the "peers" emulated by three buffers on ONE GPU (the kernel only sees
    addresses): rank 1 of a world of 3 stores its six compact planes into slab 1 of every buffer, bit-identical to the
    reference-style outputs, and touches nothing else; the distance tensor / mask of the launch are the usual ones.
    (The multi-GPU path over symmetric memory is exercised by tools/multi_gpu_check.py and bench.py --with-gather.)"""
    import ctypes

    A, world, rank, shard = 15, 3, 1, B + 1
    xyz, mask, _ = H.synthetic_batch(1200 + L, B, L, A, kind)
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    ref = sb.inter_residue_geometry()
    m, code = sb._mask_for_kernel(sb.atom_mask)
    bufs = [torch.full((6, world, shard, L, L), -55.0, device=DEV) for _ in range(world)]
    dist = torch.empty(B, L, L, A, A, device=DEV)
    dmask = torch.empty(B, L, L, A, A, dtype=m.dtype, device=DEV)
    peers = (ctypes.c_void_p * world)(*[b.data_ptr() for b in bufs])
    rc = native_lib.ps_inter_residue_geometry_push(sb.get_xyz().data_ptr(), m.data_ptr(), code, dist.data_ptr(), dmask.data_ptr(),
                                                   peers, world, rank, None, shard, B, L, A,
                                                   torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "ps_inter_residue_geometry_push")
    torch.cuda.synchronize()
    assert _cabi.last_pair_dist_plan()["sweep"] == 1
    names = ("omega", "theta", "phi", "d_ca", "d_cb", "d_no")
    for buf in bufs:
        for k, name in enumerate(names):
            got = buf[k, rank, :B]
            assert torch.equal(torch.nan_to_num(got, nan=-5.0), torch.nan_to_num(ref[name], nan=-5.0)), name
        untouched = buf.clone()
        untouched[:, rank, :B] = -55.0
        assert bool((untouched == -55.0).all()), "the push wrote outside its slab"
    assert torch.equal(torch.nan_to_num(dist, nan=-5.0), torch.nan_to_num(ref["d_ca"]._base, nan=-5.0))
    assert torch.equal(dmask, ref["d_ca_mask"]._base.to(dmask.dtype))
    # shapes the linear-sweep kernel does not take are refused, not silently gathered wrong
    small = ps.StructureBatch.from_xyz(xyz[:, :20], mask[:, :20])
    rc = native_lib.ps_inter_residue_geometry_push(small.get_xyz().data_ptr(), m.data_ptr(), code, dist.data_ptr(), dmask.data_ptr(),
                                                   peers, world, rank, None, shard, B, 20, A,
                                                   torch.cuda.current_stream().cuda_stream)
    assert rc == -1 and "linear-sweep" in _cabi.last_error()


# ------------------------------------------------------------------------------ K3 backbone
@pytest.mark.parametrize("name", SYNTHETIC + ["real_1a6v_HL"])
def test_backbone_features_match_reference_golden(native_lib, name):
    g = H.load_golden(name)
    sb = make_batch(g)
    xyz = H.t(g["xyz"])
    dihedrals, dihedral_mask = sb.backbone_dihedrals()
    ref = H.t(g["ref_bb_dihedrals"])
    assert dihedral_mask.dtype == torch.bool and tuple(dihedrals.shape) == tuple(ref.shape)
    assert torch.equal(dihedral_mask.cpu(), H.t(g["ref_bb_dihedral_mask"])), "dihedral mask must be bit-exact"
    assert torch.equal(sb.get_n_terminal_mask().cpu(), H.t(g["ref_nterm"]))
    assert torch.equal(sb.get_c_terminal_mask().cpu(), H.t(g["ref_cterm"]))
    n, ca, c = xyz[:, :, 0].double(), xyz[:, :, 1].double(), xyz[:, :, 2].double()
    big = torch.full(xyz.shape[:2], 1.0, dtype=torch.float64)
    cond = torch.stack([big.clone(), big.clone(), big.clone()], dim=-1)
    cond[:, 1:, 0] = H.dihedral_conditioning(c[:, :-1], n[:, 1:], ca[:, 1:], c[:, 1:])
    cond[:, :-1, 1] = H.dihedral_conditioning(n[:, :-1], ca[:, :-1], c[:, :-1], n[:, 1:])
    cond[:, :-1, 2] = H.dihedral_conditioning(ca[:, :-1], c[:, :-1], n[:, 1:], ca[:, 1:])
    cond = torch.nan_to_num(cond, nan=0.0)
    H.assert_angles_close(dihedrals, ref, cond, "backbone dihedrals", all_finite_tol=2e-6)
    # exact zero fill at termini (reference tests/test_StructureBatch.py:91-95)
    assert torch.equal(dihedrals.cpu() == 0, ref == 0)
    frames = sb.backbone_orientations()
    ref_frames = H.t(g["ref_frames"])
    H.assert_same_nan(frames, ref_frames, "frames")
    ok = ~torch.isnan(ref_frames)
    err = (frames.cpu()[ok] - ref_frames[ok]).abs().max().item() if ok.any() else 0.0
    assert err <= 2e-6, f"frames deviate by {err}"
    d2, m2, f2 = sb.backbone_features()
    assert torch.equal(torch.nan_to_num(d2), torch.nan_to_num(dihedrals)) and torch.equal(m2, dihedral_mask)
    assert torch.equal(torch.nan_to_num(f2), torch.nan_to_num(frames))


def test_reference_backbone_dihedral_test_case(native_lib):
    """Port of reference tests/test_StructureBatch.py:68-95 (float64 uniform xyz, 3 chains, A = 25)."""
    rng = np.random.default_rng(0)
    n_proteins, L, A = 16, 100, 25
    xyz = rng.random((n_proteins, L, A, 3))
    chain_idx = np.zeros((n_proteins, L))
    chain_idx[:, 20:60] = 1.0
    chain_idx[:, 60:] = 2.0
    sb = ps.StructureBatch.from_xyz(xyz, chain_idx=chain_idx, chain_ids=[["A", "B", "C"]] * n_proteins)
    assert sb.get_max_n_atoms_per_residue() == 25
    dihedrals, dihedral_mask = sb.backbone_dihedrals()
    assert dihedrals.shape == (n_proteins, L, 3) and dihedral_mask.shape == (n_proteins, L, 3)
    assert bool(((dihedrals >= -np.pi) & (dihedrals <= np.pi)).all())
    assert bool(((dihedrals >= -np.pi) & (dihedrals < 0)).any()) and bool(((dihedrals >= 0) & (dihedrals <= np.pi)).any())
    nterm, cterm = sb.get_n_terminal_mask(), sb.get_c_terminal_mask()
    assert bool((nterm.sum(dim=1) == 3).all()) and bool((cterm.sum(dim=1) == 3).all())
    assert bool((dihedrals[nterm][:, 0] == 0.0).all())
    assert bool((dihedrals[cterm][:, [1, 2]] == 0.0).all())
    ref, ref_mask = orc.backbone_dihedrals(torch.from_numpy(xyz).float(), torch.from_numpy(chain_idx).float(),
                                           torch.ones(n_proteins, L, dtype=torch.bool))
    assert torch.equal(dihedral_mask.cpu(), ref_mask)
    assert H.circular_diff(dihedrals.cpu(), ref).max().item() < 1e-4  # unit-cube coordinates: ill conditioned


def test_ideal_backbone_gives_identity_frames(native_lib):
    """reference tests/test_geometry.py:246-262: `frame == eye(3)` exactly."""
    import json
    ka = json.loads((H.GOLDEN / "MANIFEST.json").read_text())["known_answers"]
    ideal = torch.tensor(ka["ideal_backbone_n_ca_c_cb"])  # (4, 3): N, CA, C, CB
    xyz = torch.zeros(16, 30, 15, 3)
    xyz[:, :, :3] = ideal[:3]
    xyz[:, :, 4] = ideal[3]
    sb = ps.StructureBatch.from_xyz(xyz)
    frames = sb.backbone_orientations()
    assert frames.shape == (16, 30, 3, 3)
    assert bool((frames.cpu() == torch.eye(3).expand(16, 30, -1, -1)).all())
    fr = ps.geometry.gram_schmidt(xyz[:, :, 0], xyz[:, :, 1], xyz[:, :, 2])
    assert bool((fr.cpu() == torch.eye(3).expand(16, 30, -1, -1)).all())
    tr = sb.backbone_translations()
    assert tr.shape == (16, 30, 3) and tr.data_ptr() == sb.xyz[:, :, 1].data_ptr()  # a view, like the reference


# ------------------------------------------------------------------------------ K4 statistics
@pytest.mark.parametrize("name", SYNTHETIC + ["real_1a6v_HL"])
def test_standardize_and_center_of_mass_match_reference_golden(native_lib, name):
    g = H.load_golden(name)
    sb = make_batch(g)
    com = sb.center_of_mass()
    ref_com = H.t(g["ref_com"])
    H.assert_same_nan(com, ref_com, "com")
    assert torch.allclose(com.cpu(), ref_com, rtol=1e-5, atol=1e-4, equal_nan=True)
    original = sb.get_xyz().clone()
    sb.standardize()
    assert tuple(sb.mu.shape) == (sb.batch_size, 3) and tuple(sb.std.shape) == (sb.batch_size, 3)
    assert torch.allclose(sb.mu.cpu(), H.t(g["ref_mu"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(sb.std.cpu(), H.t(g["ref_sd"]), rtol=1e-5, atol=1e-6)
    ref_xyz = H.t(g["ref_std_xyz"])
    H.assert_same_nan(sb.get_xyz(), ref_xyz, "standardized xyz")
    assert torch.allclose(sb.get_xyz().cpu(), ref_xyz, rtol=1e-4, atol=1e-5, equal_nan=True)
    with pytest.raises(ValueError):
        sb.standardize()
    sb.unstandardize()
    # reference tests/test_StructureBatch.py:246-255
    assert torch.allclose(sb.get_xyz(), original, rtol=1e-4, atol=1e-5, equal_nan=True)
    with pytest.raises(ValueError):
        sb.unstandardize()


def test_center_at_moves_the_ca_centre(native_lib):
    """reference tests/test_StructureBatch.py:258-275."""
    xyz, mask, chain_idx = H.synthetic_batch(5, 4, 50, 15, "bool")
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    sb.center_at()
    assert torch.allclose(sb.center_of_mass(), torch.zeros(4, 3, device=DEV), atol=1e-4)
    target = torch.tensor([[1.0, 2.0, 3.0]]).repeat(4, 1)
    sb.center_at(target)
    assert torch.allclose(sb.center_of_mass().cpu(), target, atol=1e-4)
    sb.center_at(torch.tensor([5.0, 5.0, 5.0]))
    assert torch.allclose(sb.center_of_mass().cpu(), torch.full((4, 3), 5.0), atol=1e-4)
    with pytest.raises(ValueError):
        sb.center_at(torch.zeros(3, 3))
    with pytest.raises(ValueError):
        sb.center_at(torch.zeros(4, 2))


def test_standardize_statistics_vs_oracle_config4_slice(native_lib):
    """BASELINE config 4 shape (L = 128, A = 15), 64 of its 1024 structures."""
    xyz, mask, _ = H.synthetic_batch(4, 64, 128, 15, "bool")
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    sb.standardize()
    ref_xyz, mu, sd = orc.standardize_per_structure(xyz, mask)
    assert torch.allclose(sb.mu.cpu(), mu, rtol=1e-5, atol=1e-5)
    assert torch.allclose(sb.std.cpu(), sd, rtol=1e-5, atol=1e-6)
    assert torch.allclose(sb.get_xyz().cpu(), ref_xyz, rtol=1e-4, atol=1e-5, equal_nan=True)
    valid = sb.get_xyz()[mask.to(DEV)]
    assert not bool(torch.isnan(valid).any())  # reference tests/test_StructureBatch.py:218-226


@pytest.mark.parametrize("B,L,A,kind", [(1024, 128, 15, "bool"), (7, 229, 15, "bool"), (3, 701, 15, "float"), (300, 33, 15, "bool"),
                                        (2, 2101, 15, "bool"), (1, 3000, 15, "bool"), (5, 10, 7, "float"), (1, 1, 15, "bool"),
                                        (16, 512, 15, "bool"), (64, 128, 16, "float"), (2, 4096, 15, "bool"), (9, 4, 15, "bool"),
                                        (3, 8, 4, "float")])
def test_standardize_register_resident_kernel_vs_oracle_and_three_pass_kernel(native_lib, B, L, A, kind):
    """K4: the register-resident single-read kernels (default: the quad kernel — 128-bit loads, four atoms per group —
    where a structure holds a multiple of 4 atoms, else the scalar-mapped one; one CTA or a cluster of 2-8 CTAs per
    structure) against the CPU oracle (reference protstruc.py:696-734 per structure) and against the three-pass
    kernel of round 1 — same element arithmetic, so the statistics agree to the last bits of the partial sums."""
    xyz, mask, _ = H.synthetic_batch(4000 + L, B, L, A, kind)
    x = xyz.to(DEV)
    m = mask.to(DEV).contiguous()
    code = _cabi.PS_MASK_BOOL if kind == "bool" else _cabi.PS_MASK_F32
    s = torch.cuda.current_stream().cuda_stream
    res = []
    for variant in (0, 2, 1):  # default, scalar-mapped register kernel, three-pass kernel
        mu, sd = torch.full((B, 3), -1.0, device=DEV), torch.full((B, 3), -1.0, device=DEV)
        out = torch.full_like(x, -99.0)
        _cabi.check(native_lib.ps_masked_stats_ex(x.data_ptr(), m.data_ptr(), code, B, L, A, mu.data_ptr(), sd.data_ptr(),
                                                  out.data_ptr(), variant, s), "ps_masked_stats_ex")
        res.append((mu.cpu(), sd.cpu(), out.cpu()))
    (mu0, sd0, x0), (mu2, sd2, x2), (mu1, sd1, x1) = res
    for mu_v, sd_v, x_v in ((mu2, sd2, x2), (mu1, sd1, x1)):
        assert torch.allclose(mu0, mu_v, rtol=1e-6, atol=1e-6, equal_nan=True)
        assert torch.allclose(sd0, sd_v, rtol=1e-6, atol=1e-7, equal_nan=True)
        assert torch.equal(torch.isnan(x0), torch.isnan(x_v))
        assert torch.allclose(x0, x_v, rtol=1e-5, atol=1e-5, equal_nan=True)
    if B * L <= 8192:
        rx, rmu, rsd = orc.standardize_per_structure(xyz, mask)
        assert torch.allclose(mu0, rmu, rtol=1e-5, atol=1e-5, equal_nan=True)
        assert torch.allclose(sd0, rsd, rtol=1e-5, atol=1e-6, equal_nan=True)
        H.assert_same_nan(x0, rx, "standardized xyz")
        assert torch.allclose(x0, rx, rtol=1e-4, atol=1e-5, equal_nan=True)


@pytest.mark.parametrize("B,L,kind", [(2, 2101, "bool"), (3, 701, "float"), (1, 4097, "bool"), (5, 300, "bool")])
def test_standardize_few_large_structures_take_the_cluster_path(native_lib, B, L, kind):
    """Few, large structures: every structure is reduced by a thread-block cluster (2 - 8 CTAs exchanging partial
    sums through distributed shared memory); shares are uneven on purpose (atom counts not divisible by the cluster
    size).  Same statistics and coordinates as the oracle, and the same as the C-ABI call writing in place."""
    xyz, mask, _ = H.synthetic_batch(90 + L, B, L, 15, kind)
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    sb.standardize()
    ref_xyz, mu, sd = orc.standardize_per_structure(xyz, mask)
    assert torch.allclose(sb.mu.cpu(), mu, rtol=1e-5, atol=1e-5)
    assert torch.allclose(sb.std.cpu(), sd, rtol=1e-5, atol=1e-6)
    assert torch.allclose(sb.get_xyz().cpu(), ref_xyz, rtol=1e-4, atol=1e-5, equal_nan=True)
    # in place (xyz_out aliases xyz): a CTA only ever reads the share it later overwrites
    x = xyz.to(DEV).contiguous()
    m = (mask.float() if kind == "float" else mask).to(DEV).contiguous()
    mu_d, sd_d = torch.empty(B, 3, device=DEV), torch.empty(B, 3, device=DEV)
    rc = native_lib.ps_masked_stats(x.data_ptr(), m.data_ptr(), 1 if kind == "float" else 0, B, L, 15, mu_d.data_ptr(),
                                    sd_d.data_ptr(), x.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "ps_masked_stats")
    assert torch.equal(mu_d, sb.mu.reshape(B, 3)) and torch.equal(sd_d, sb.std.reshape(B, 3))
    assert torch.equal(torch.nan_to_num(x, nan=-3.0), torch.nan_to_num(sb.get_xyz(), nan=-3.0))


# ------------------------------------------------------------------------------ K5 diffusion
@pytest.mark.parametrize("name", SYNTHETIC + ["real_1a6v_HL"])
def test_diffuse_xyz_is_bit_exact_given_the_noise(native_lib, name):
    g = H.load_golden(name)
    sb = make_batch(g)
    before = sb.get_xyz()
    sb.diffuse_xyz(H.t(g["ref_beta"]), noise=H.t(g["ref_noise"]))
    after = sb.get_xyz()
    assert after.data_ptr() != before.data_ptr()  # rebinds, does not alias
    ref = H.t(g["ref_diffused"])
    H.assert_same_nan(after, ref, "diffused")
    assert torch.equal(torch.nan_to_num(after.cpu()), torch.nan_to_num(ref)), "diffuse_xyz must be bit-exact"


def test_philox_stream_matches_the_oracle_definition_and_is_normal(native_lib):
    n = 1 << 20
    out = torch.empty(n, device=DEV)
    rc = native_lib.ps_philox_normal(out.data_ptr(), n, 1234, 5, 0, torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "ps_philox_normal")
    z = out.cpu().double().numpy()
    ref = orc.philox_normal(n, seed=1234, step=5)
    # same counters / same Box-Muller; the kernel evaluates log / sin / cos with MUFU intrinsics
    err = np.abs(z - ref)
    assert err.max() < 2e-3 and np.quantile(err, 0.999) < 1e-5
    assert abs(z.mean()) < 4e-3 and abs(z.std() - 1.0) < 4e-3
    assert abs(((z - z.mean()) ** 3).mean()) < 1e-2 and abs(((z - z.mean()) ** 4).mean() - 3.0) < 3e-2
    # Kolmogorov-Smirnov distance to N(0,1)
    zs = np.sort(z)
    cdf = 0.5 * (1.0 + np.vectorize(math.erf)(zs[::64] / math.sqrt(2.0)))
    emp = (np.arange(0, n, 64) + 0.5) / n
    assert np.max(np.abs(cdf - emp)) < 3e-3
    # tail sanity: finite, |z| can exceed 4
    assert np.isfinite(z).all() and np.abs(z).max() > 4.0
    shard = torch.empty(4096, device=DEV)
    rc = native_lib.ps_philox_normal(shard.data_ptr(), 4096, 1234, 5, 40000, torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "ps_philox_normal")
    assert torch.equal(shard, out[40000:44096]), "stream must be addressed by the global element index"


def test_diffusion_distribution_fused_steps_and_shard_invariance(native_lib):
    B, L, A, T = 8, 128, 15, 300
    betas_all = orc.cosine_variance_schedule(T)[:T]  # t = 0 .. T-1, like the tutorial loop
    betas = betas_all[:, None].repeat(1, B)
    xyz, mask, _ = H.synthetic_batch(44, B, L, A, "bool", nan_masked=False, full_length=True)
    ps.manual_seed(99)
    seq = ps.StructureBatch.from_xyz(xyz, mask)
    seq.standardize()
    start = seq.get_xyz().clone()
    for t in range(T):
        seq.diffuse_xyz(betas[t])
    ps.manual_seed(99)
    fused = ps.StructureBatch.from_xyz(xyz, mask)
    fused.standardize()
    fused.diffuse_xyz_steps(betas)
    assert torch.equal(seq.get_xyz(), fused.get_xyz()), "fused T-step kernel must equal T single steps bit for bit"
    final = fused.get_xyz()
    assert abs(final.mean().item()) < 0.02 and abs(final.std().item() - 1.0) < 0.02  # final std 1.001
    assert not torch.equal(final, start)
    # sharding the batch over 2 "ranks" reproduces the unsharded result exactly
    from protstruc_b200.sharding import shard_structure_batch
    parts = []
    for rank in range(2):
        ps.manual_seed(99)
        sb = shard_structure_batch(start.cpu(), mask, rank=rank, world_size=2)
        sb.diffuse_xyz_steps(betas[:, rank * 4:(rank + 1) * 4])
        parts.append(sb.get_xyz())
    assert torch.equal(torch.cat(parts), final)
    # one step: the implied noise (x' - sqrt(1-b) x) / sqrt(b) is N(0,1)
    ps.manual_seed(5)
    one = ps.StructureBatch.from_xyz(xyz, mask)
    beta = torch.full((B,), 0.25)
    x0 = one.get_xyz().clone()
    one.diffuse_xyz(beta)
    z = (one.get_xyz() - math.sqrt(0.75) * x0) / 0.5
    assert abs(z.mean().item()) < 0.02 and abs(z.std().item() - 1.0) < 0.02


def test_diffusion_loop_equals_the_reference_style_python_loop(native_lib):
    """`diffusion_loop` (fused, and with the trajectory kept) against T separate `diffuse_xyz` calls on the same noise
    stream: bit-identical states at every step (reference loop: README.md:131-146)."""
    B, L, A, T = 4, 61, 15, 25
    xyz, mask, _ = H.synthetic_batch(61, B, L, A, "bool", nan_masked=False, full_length=True)
    betas = orc.cosine_variance_schedule(300)[40:40 + T]
    ps.manual_seed(31)
    ref = ps.StructureBatch.from_xyz(xyz, mask)
    states = []
    for t in range(T):
        ref.diffuse_xyz(betas[t].repeat(B))
        states.append(ref.get_xyz().clone())
    ps.manual_seed(31)
    a = ps.StructureBatch.from_xyz(xyz, mask)
    assert a.diffusion_loop(betas) is None
    assert torch.equal(a.get_xyz(), states[-1])
    ps.manual_seed(31)
    b = ps.StructureBatch.from_xyz(xyz, mask)
    traj = b.diffusion_loop(betas[:, None].repeat(1, B), return_trajectory=True)
    assert tuple(traj.shape) == (T, B, L, A, 3) and torch.equal(traj, torch.stack(states))
    # the final state is a copy: an in-place mutator must not reach into the returned trajectory
    assert torch.equal(b.get_xyz(), traj[T - 1]) and b.get_xyz().data_ptr() != traj[T - 1].data_ptr()
    b.translate(torch.ones(B, 1, 3, device=DEV))
    assert torch.equal(traj[T - 1], states[-1])
    with pytest.raises(ValueError):
        b.diffusion_loop(torch.zeros(3, B + 1))


def test_philox_stream_is_addressed_per_element_for_any_shard_offset(native_lib):
    """north_star / SURVEY 8(e): results independent of the number of GPUs.  L = 229 (the real 1a6v_HL length) gives
    10,305 floats per structure, so shard offsets are not multiples of 4; the stream is addressed per element
    (counter = global index >> 2, lane = global index & 3) and every shard reproduces its slice of the global
    stream bit for bit — for the raw normals, one diffusion step and the fused T-step kernel, on 2, 3 and 5 ranks."""
    from protstruc_b200.sharding import shard_bounds, shard_structure_batch

    s = torch.cuda.current_stream().cuda_stream
    n = 50_001
    full = torch.empty(n, device=DEV)
    _cabi.check(native_lib.ps_philox_normal(full.data_ptr(), n, 77, 3, 0, s), "ps_philox_normal")
    for off, cnt in ((1, 17), (2, 4096), (3, 5), (10_305, 10_305), (20_610 + 3, 1), (49_999, 2)):
        part = torch.full((cnt,), -7.0, device=DEV)
        _cabi.check(native_lib.ps_philox_normal(part.data_ptr(), cnt, 77, 3, off, s), "ps_philox_normal")
        assert torch.equal(part, full[off:off + cnt]), (off, cnt)
    ref = orc.philox_normal(64, seed=77, step=3, elem_offset=10_305)
    assert np.abs(full[10_305:10_305 + 64].cpu().double().numpy() - ref).max() < 2e-3

    B, L, A, T = 5, 229, 15, 7
    xyz, mask, _ = H.synthetic_batch(229, B, L, A, "bool", nan_masked=False, full_length=True)
    betas = orc.cosine_variance_schedule(300)[100:100 + T, None].repeat(1, B).contiguous()
    ps.manual_seed(4321)
    whole = ps.StructureBatch.from_xyz(xyz, mask)
    whole.diffuse_xyz(betas[0])
    one_step = whole.get_xyz().clone()
    whole.diffuse_xyz_steps(betas[1:])
    all_steps = whole.get_xyz()
    for world in (2, 3, 5):
        got_one, got_all = [], []
        for rank in range(world):
            lo, hi = shard_bounds(B, world, rank)
            ps.manual_seed(4321)  # every rank seeds alike, as ranks of one job do
            sb = shard_structure_batch(xyz, mask, rank=rank, world_size=world)
            assert sb._noise_elem_offset == lo * L * A * 3
            sb.diffuse_xyz(betas[0, lo:hi])
            got_one.append(sb.get_xyz().clone())
            sb.diffuse_xyz_steps(betas[1:, lo:hi])
            got_all.append(sb.get_xyz())
        assert torch.equal(torch.cat(got_one), one_step), f"one step differs on {world} ranks"
        assert torch.equal(torch.cat(got_all), all_steps), f"fused steps differ on {world} ranks"


def test_noise_stream_follows_the_generator_state(native_lib):
    """alternating generators must not rewind each other, and re-seeding (torch.manual_seed, a re-created
    generator with the same seed) must restart the stream reproducibly — like torch.randn_like in the reference
    (protstruc.py:876), whose noise is a function of the generator state."""
    xyz, mask, _ = H.synthetic_batch(8, 2, 40, 15, "bool", nan_masked=False, full_length=True)
    beta = torch.tensor([0.3, 0.6])

    def step(sb, **kw):
        sb.diffuse_xyz(beta, **kw)
        return sb.get_xyz().clone()

    def fresh():
        return ps.StructureBatch.from_xyz(xyz, mask)

    g1, g2 = torch.Generator().manual_seed(11), torch.Generator().manual_seed(22)
    a = [step(fresh(), generator=g) for g in (g1, g2, g1, g2)]
    assert not torch.equal(a[0], a[2]) and not torch.equal(a[1], a[3]), "switching generators replayed noise"
    assert not torch.equal(a[0], a[1])
    g1b = torch.Generator().manual_seed(11)  # re-created, same seed: same stream from the start
    assert torch.equal(step(fresh(), generator=g1b), a[0]) and torch.equal(step(fresh(), generator=g1b), a[2])
    g1.manual_seed(11)  # re-seeded in place
    assert torch.equal(step(fresh(), generator=g1), a[0])
    torch.manual_seed(5)
    b0, b1 = step(fresh()), step(fresh())
    assert not torch.equal(b0, b1), "consecutive calls must not reuse noise"
    torch.manual_seed(5)  # plain torch re-seed (not ps.manual_seed) restarts too
    assert torch.equal(step(fresh()), b0) and torch.equal(step(fresh()), b1)
    torch.manual_seed(5)
    torch.rand(3)  # someone else consumed the generator: a different, but reproducible, stream
    c0 = step(fresh())
    torch.manual_seed(5)
    torch.rand(3)
    assert torch.equal(step(fresh()), c0) and not torch.equal(c0, b0)


# ------------------------------------------------------------------------------ geometry free functions
def test_geometry_known_answers(native_lib):
    """reference tests/test_geometry.py:10-190 and tests/test_decorator.py type propagation."""
    assert ps.geometry.dot(torch.tensor([1, 2, 3]), torch.tensor([4, 5, 6])) == 32
    assert ps.geometry.dot(np.array([1, 2, 3]), np.array([4, 5, 6])) == 32
    a32 = np.array([[1, 2, 3], [4, 5, 6]]).astype(np.float32)
    nrm = ps.geometry.norm(a32)
    assert isinstance(nrm, np.ndarray) and nrm.shape == (2, 1) and np.allclose(nrm, [[14**0.5], [77**0.5]])
    nt = ps.geometry.norm(torch.from_numpy(a32))
    assert isinstance(nt, torch.Tensor) and nt.shape == (2, 1)
    a = np.array([[1, 0, 0], [1, 0, 0]], dtype=np.float32)
    b = np.zeros((2, 3), dtype=np.float32)
    c = np.array([[0, 1, 0], [0.5, np.sqrt(3) / 2, 0]], dtype=np.float32)
    ang = ps.geometry.angle(a, b, c, to_degree=True)
    assert isinstance(ang, np.ndarray) and ang.shape == (2,) and np.allclose(ang, [90.0, 60.0])
    ang_t = ps.geometry.angle(torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(c), to_degree=True)
    assert isinstance(ang_t, torch.Tensor) and torch.allclose(ang_t.cpu(), torch.tensor([90.0, 60.0]))
    pts = [np.array([[1, 0, 0]], dtype=np.float32), np.zeros((1, 3), dtype=np.float32),
           np.array([[0, 1, 0]], dtype=np.float32), np.array([[0, 1, 1]], dtype=np.float32)]
    dih = ps.geometry.dihedral(*pts, to_degree=True)
    assert isinstance(dih, np.ndarray) and dih.shape == (1,) and np.allclose(dih, [-90.0])
    dih_t = ps.geometry.dihedral(*[torch.from_numpy(p) for p in pts], to_degree=True)
    assert isinstance(dih_t, torch.Tensor) and dih_t.shape == (1,) and torch.allclose(dih_t.cpu(), torch.tensor([-90.0]))
    dih_113 = ps.geometry.dihedral(*[torch.from_numpy(p).reshape(1, 1, 3) for p in pts], to_degree=True)
    assert dih_113.shape == (1, 1)
    mixed = ps.geometry.dihedral(pts[0], torch.from_numpy(pts[1]), pts[2], pts[3])
    assert isinstance(mixed, torch.Tensor) and abs(mixed.item() + math.pi / 2) < 1e-6
    fr = ps.geometry.gram_schmidt(torch.randn(16, 30, 3), torch.randn(16, 30, 3), torch.randn(16, 30, 3))
    assert fr.shape == (16, 30, 3, 3)
    ident = torch.einsum("bnij,bnik->bnjk", fr, fr)
    assert torch.allclose(ident.cpu(), torch.eye(3).expand(16, 30, 3, 3), atol=1e-5)


def test_geometry_functions_vs_oracle_random(native_lib):
    g = torch.Generator().manual_seed(8)
    pts = [3.0 * torch.randn(5000, 3, generator=g) for _ in range(4)]
    dih = ps.geometry.dihedral(*pts)
    ref = orc.dihedral(*pts)
    cond = H.dihedral_conditioning(*[p.double() for p in pts])
    H.assert_angles_close(dih, ref, cond, "geometry.dihedral")
    ang = ps.geometry.angle(*pts[:3])
    H.assert_angles_close(ang, orc.planar_angle(*pts[:3]), H.planar_conditioning(*[p.double() for p in pts[:3]]),
                          "geometry.angle", circular=False)
    # dot / norm / unit are kernels too (row a9): points (D = 3) bit for bit, other widths to rounding
    for D in (3, 7, 1):
        u, v = 5.0 * torch.randn(4, 250, D, generator=g), 5.0 * torch.randn(4, 250, D, generator=g)
        got_dot, got_norm, got_unit = ps.geometry.dot(u, v), ps.geometry.norm(u), ps.geometry.unit(u)
        assert got_dot.shape == (4, 250, 1) and got_norm.shape == (4, 250, 1) and got_unit.shape == (4, 250, D)
        if D == 3:
            assert torch.equal(got_dot.cpu(), orc.dot(u, v)) and torch.equal(got_norm.cpu(), orc.norm(u))
            assert torch.equal(got_unit.cpu(), u / orc.norm(u))
        else:
            assert torch.allclose(got_dot.cpu(), orc.dot(u, v), rtol=1e-5, atol=1e-4)
            assert torch.allclose(got_norm.cpu(), orc.norm(u), rtol=1e-6) and torch.allclose(got_unit.cpu(), u / orc.norm(u), rtol=1e-6, atol=1e-7)
    assert torch.isnan(ps.geometry.unit(torch.zeros(2, 3))).all()  # 0 / 0, like the reference
    assert isinstance(ps.geometry.unit(np.ones((2, 3), dtype=np.float32)), np.ndarray)
    fr = ps.geometry.gram_schmidt(*pts[:3]).cpu()
    ref_fr = orc.frames_from_points(*pts[:3])
    sin = H.planar_conditioning(*[p.double() for p in pts[:3]])  # angle between (a-b) and (c-b)
    err = (fr - ref_fr).abs().amax(dim=(-1, -2)).double()
    assert bool((err[sin >= H.SIN_GATE] <= 5e-6).all()), "frames deviate in the well-conditioned region"
    assert bool((err <= 1e-6 / sin.clamp_min(1e-12)).all()), "frames deviate beyond 1e-6 / sin"


# ------------------------------------------------------------------------------ API behaviour on the device
def test_error_behaviour_matches_the_reference(native_lib):
    xyz, mask, _ = H.synthetic_batch(9, 2, 8, 15, "bool")
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    with pytest.raises(ValueError, match="Atom XX is not valid."):
        sb.pairwise_dihedrals(["CA", "XX"], ["CA", "CB"])
    with pytest.raises(KeyError):
        sb.backbone_orientations(a1="XX")
    with pytest.raises(KeyError):
        sb.backbone_translations("nope")
    with pytest.raises(ValueError):
        sb.standardize(atom_mask=mask, residue_mask=mask.any(-1))
    no_mask = ps.StructureBatch.from_xyz(xyz)
    with pytest.raises(TypeError):
        no_mask.pairwise_distance_matrix()
    d, m = no_mask.backbone_dihedrals()  # works without atom_mask, like the reference
    assert d.shape == (2, 8, 3)
    # C-ABI level: bad arguments give a status code and a message, never a crash
    rc = native_lib.ps_pair_dist_mask(None, None, 0, None, None, 1, 1, 15, None)
    assert rc == -2 and b"xyz" in native_lib.ps_last_error_string()
    rc = native_lib.ps_pair_dist_mask(sb.xyz.data_ptr(), None, 0, sb.xyz.data_ptr(), None, 0, 8, 15, None)
    assert rc == -1


def test_dtype_and_input_handling(native_lib):
    """float64 numpy input (what the reference's tests feed) is accepted and computed in fp32; mask dtypes
    other than bool / float32 come back in the caller's dtype."""
    rng = np.random.default_rng(3)
    xyz64 = rng.standard_normal((2, 10, 15, 3)) * 5
    mask_i64 = (rng.random((2, 10, 15)) < 0.6).astype(np.int64)
    sb = ps.StructureBatch.from_xyz(xyz64, mask_i64)
    assert sb.get_xyz().dtype == torch.float32
    dist, dist_mask = sb.pairwise_distance_matrix()
    ref_dist, ref_mask = orc.pair_distances(torch.from_numpy(xyz64).float(), torch.from_numpy(mask_i64))
    H.assert_distances_close(dist, ref_dist)
    assert dist_mask.dtype == torch.int64 and torch.equal(dist_mask.cpu(), ref_mask)
    assert sb.get_total_lengths().shape == (2,)


# ------------------------------------------------------------------------------ rows f1 / f2 / f4
def test_rigid_frame_family_matches_reference_golden(native_lib):
    """get_local_xyz, rotate, translate, from_backbone_orientations_translations (reference
    protstruc.py:263-362, 662-694; reference tests/test_StructureBatch.py:179-207)."""
    g = H.load_golden("frames_align_topk")
    ids = [["A", "B"]] * 4
    new = lambda: ps.StructureBatch.from_xyz(g["xyz"], g["atom_mask"], g["chain_idx"], ids)  # noqa: E731
    local = new().get_local_xyz()
    assert tuple(local.shape) == (4, 33, 15, 3)
    H.assert_same_nan(local, H.t(g["ref_local_xyz"]), "local xyz")  # zero-padded residues have no frame
    assert torch.allclose(local.cpu(), H.t(g["ref_local_xyz"]), rtol=1e-5, atol=2e-4, equal_nan=True)
    sb = new()
    sb.rotate(H.t(g["rotation"]))
    assert torch.allclose(sb.get_xyz().cpu(), H.t(g["ref_rotated"]), rtol=1e-5, atol=1e-4)
    sb = new()
    sb.rotate(H.t(g["rotation"])[0])
    assert torch.allclose(sb.get_xyz().cpu(), H.t(g["ref_rotated_single"]), rtol=1e-5, atol=1e-4)
    for name, atomwise in (("res", False), ("one", False), ("atom", True)):
        sb = new()
        before = sb.get_xyz().data_ptr()
        sb.translate(H.t(g[f"tr_{name}"]), atomwise=atomwise)
        assert sb.get_xyz().data_ptr() == before  # in place, like the reference's `+=`
        assert torch.equal(sb.get_xyz().cpu(), H.t(g[f"ref_translated_{name}"])), f"translate {name} must be exact"
    for cb in (0, 1):
        sb2 = ps.StructureBatch.from_backbone_orientations_translations(
            H.t(g["frames"]), H.t(g["frame_translations"]), H.t(g["chain_idx"]), ids, None, include_cb=bool(cb))
        assert sb2.get_max_n_atoms_per_residue() == 15
        assert sb2.get_atom_mask().dtype == torch.float32
        assert torch.equal(sb2.get_atom_mask().cpu(), H.t(g[f"ref_from_frames_mask_cb{cb}"]))
        H.assert_same_nan(sb2.get_xyz(), H.t(g[f"ref_from_frames_xyz_cb{cb}"]), "frames -> backbone")
        assert torch.allclose(sb2.get_xyz().cpu(), H.t(g[f"ref_from_frames_xyz_cb{cb}"]), rtol=1e-5, atol=1e-4,
                              equal_nan=True)
    # frames -> coordinates -> frames is the identity on the backbone
    sb = new()
    rebuilt = ps.StructureBatch.from_backbone_orientations_translations(sb.backbone_orientations(), sb.backbone_translations())
    assert torch.allclose(rebuilt.backbone_orientations(), sb.backbone_orientations(), atol=1e-5, equal_nan=True)
    assert torch.allclose(rebuilt.backbone_translations()[~torch.isnan(rebuilt.backbone_translations())],
                          sb.backbone_translations()[~torch.isnan(rebuilt.backbone_translations())], atol=1e-5)


def test_translate_broadcasts_like_the_reference(native_lib):
    """every shape torch's `xyz += translation` accepts (protstruc.py:662-679), including a size-1
    coordinate axis and non-contiguous translations; in place (the tensor object is kept)."""
    xyz, mask, _ = H.synthetic_batch(31, 3, 21, 15, "bool", nan_masked=False, full_length=True)
    g = torch.Generator().manual_seed(1)
    cases = [(torch.randn(3, 21, 3, generator=g), False), (torch.randn(3, 1, 3, generator=g), False),
             (torch.randn(3, 21, 1, generator=g), False), (torch.randn(1, 1, 1, generator=g), False),
             (torch.randn(3, 21, 6, generator=g)[:, :, ::2], False), (torch.randn(3, 21, 15, 3, generator=g), True),
             (torch.randn(3, 21, 15, 1, generator=g), True), (torch.randn(1, 21, 1, 3, generator=g), True)]
    for tr, atomwise in cases:
        sb = ps.StructureBatch.from_xyz(xyz, mask)
        before = sb.get_xyz()
        sb.translate(tr, atomwise=atomwise)
        assert sb.get_xyz().data_ptr() == before.data_ptr()
        ref = xyz + (tr if atomwise else tr.unsqueeze(-2))
        assert torch.equal(sb.get_xyz().cpu(), ref), (tuple(tr.shape), atomwise)
    with pytest.raises(RuntimeError):
        ps.StructureBatch.from_xyz(xyz, mask).translate(torch.zeros(3, 20, 3))


def test_align_with_rank_deficient_selections(native_lib):
    """collinear selections, two-atom and one-atom masks and an empty mask must give a proper rotation
    (the reference's SVD path does, geometry.py:442-480) and leave no NaN behind."""
    xyz, mask, _ = H.synthetic_batch(77, 4, 12, 15, "bool", nan_masked=False, full_length=True)
    g = torch.Generator().manual_seed(3)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    q = q * torch.sign(torch.linalg.det(q))
    target_xyz = xyz @ q.T + torch.tensor([3.0, -2.0, 5.0])
    sel = torch.zeros(4, 12, 15, dtype=torch.bool)
    line = torch.linspace(-5, 5, 12)[:, None] * torch.tensor([1.0, 2.0, -1.0])
    xyz[0, :, 1] = line  # structure 0: collinear CA atoms selected
    target_xyz[0, :, 1] = line @ q.T + torch.tensor([3.0, -2.0, 5.0])
    sel[0, :, 1] = True
    sel[1, 0, 1] = sel[1, 5, 1] = True  # two atoms
    sel[2, 3, 1] = True  # one atom
    # structure 3: nothing selected
    src = ps.StructureBatch.from_xyz(xyz.clone(), mask)
    tgt = ps.StructureBatch.from_xyz(target_xyz, mask)
    rot = src.align(tgt, atom_mask=sel)
    torch.cuda.synchronize()
    rot_c, out = rot.cpu().double(), src.get_xyz().cpu()
    assert bool(torch.isfinite(rot_c).all()) and bool(torch.isfinite(out).all())
    eye = torch.eye(3, dtype=torch.float64)
    assert (rot_c @ rot_c.transpose(1, 2) - eye).abs().max().item() < 1e-5
    assert (torch.linalg.det(rot_c) - 1.0).abs().max().item() < 1e-5
    for b in (0, 1, 2):  # the selected atoms are superimposed (distances between them allow it exactly)
        m = sel[b]
        assert (out[b][m] - target_xyz[b][m]).abs().max().item() < 2e-3, b
    assert torch.equal(out[3], xyz[3]), "an empty selection must not move the structure"


def test_empty_batches_are_no_ops(native_lib):
    """B == 0 or L == 0 (e.g. residue_masked_select with an all-False mask) never reach the C-ABI."""
    sb = ps.StructureBatch.from_xyz(torch.zeros(1, 4, 15, 3), torch.ones(1, 4, 15, dtype=torch.bool))
    empty = sb.residue_masked_select(torch.zeros(1, 4, dtype=torch.bool))
    assert empty.get_max_n_residues() == 0
    for e in (empty, ps.StructureBatch.from_xyz(torch.zeros(0, 7, 15, 3), torch.ones(0, 7, 15, dtype=torch.bool))):
        B, L = e.get_batch_size(), e.get_max_n_residues()
        e.translate(torch.zeros(B, L, 3))
        e.rotate(torch.eye(3))
        e.center_at()
        e.standardize()
        assert tuple(e.mu.shape) == (B, 3)
        e.unstandardize()
        e.diffuse_xyz(torch.zeros(B))
        e.diffuse_xyz_steps(torch.zeros(4, B))
        assert tuple(e.align(e).shape) == (B, 3, 3)
        assert tuple(e.get_xyz().shape) == (B, L, 15, 3)
        assert tuple(e.inter_residue_geometry()["omega"].shape) == (B, L, L)
    assert tuple(empty.get_topk_nearest_residue_mask(torch.zeros(2, 3)).shape) == (1, 0)


def test_align_matches_reference_golden(native_lib):
    """Batched Kabsch vs the reference's per-structure SVD loop (protstruc.py:880-918)."""
    g = H.load_golden("frames_align_topk")
    src = ps.StructureBatch.from_xyz(g["xyz"], g["atom_mask"])
    tgt = ps.StructureBatch.from_xyz(g["align_target"], g["atom_mask"])
    rot = src.align(tgt)
    assert tuple(rot.shape) == (4, 3, 3)
    assert torch.allclose(rot @ rot.transpose(1, 2), torch.eye(3, device=DEV).expand(4, 3, 3), atol=1e-5)
    assert torch.allclose(torch.linalg.det(rot), torch.ones(4, device=DEV), atol=1e-5)
    assert torch.allclose(src.get_xyz().cpu(), H.t(g["ref_aligned"]), rtol=1e-5, atol=5e-4)
    # after alignment the masked RMSD to the target is at the noise level (0.05 A per coordinate)
    m = H.t(g["atom_mask"]).to(DEV)
    rmsd = ((src.get_xyz() - tgt.get_xyz())[m] ** 2).sum(-1).mean().sqrt().item()
    assert rmsd < 0.15
    one = ps.StructureBatch.from_xyz(g["xyz"], g["atom_mask"])
    one.align(ps.StructureBatch.from_xyz(g["align_target"][:1], g["atom_mask"][:1]))
    assert torch.allclose(one.get_xyz().cpu(), H.t(g["ref_aligned_to_first"]), rtol=1e-5, atol=5e-4)
    with pytest.raises(ValueError):
        src.align(ps.StructureBatch.from_xyz(g["align_target"][:2], g["atom_mask"][:2]))


def _kabsch_fp64(a, b, m):
    """Reference Kabsch of geometry.py:442-480 in float64: rotation R and translation t with  b ~ a R^T + t."""
    a, b = a[m].double(), b[m].double()
    ca, cb = a.mean(0), b.mean(0)
    h = (a - ca).T @ (b - cb)
    u, _, vt = torch.linalg.svd(h)
    d = torch.sign(torch.linalg.det(vt.T @ u.T))
    r = vt.T @ torch.diag(torch.tensor([1.0, 1.0, float(d)], dtype=torch.float64)) @ u.T
    return r, cb - r @ ca


@pytest.mark.parametrize("B,n_atoms", [(5, 960), (3, 7680), (40, 256), (2, 1924)])
def test_kabsch_vector_and_scalar_kernels_agree_with_an_fp64_svd(native_lib, B, n_atoms):
    """Round 2: structures of a multiple of 4 atoms take `kabsch_quad_kernel` (128-bit loads, fp32 Jacobi angles); a mask
    pointer that is not 4-byte aligned keeps the scalar kernel.  Both against a float64 SVD Kabsch, through the C-ABI."""
    g = torch.Generator().manual_seed(11 * B + n_atoms)
    a = 30.0 * torch.randn(B, n_atoms, 3, generator=g) + 100.0 * torch.randn(B, 1, 3, generator=g)  # far from the origin
    q, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
    q = q * torch.sign(torch.linalg.det(q))[:, None, None]
    b = a @ q.transpose(1, 2) + 50.0 * torch.randn(B, 1, 3, generator=g) + 0.05 * torch.randn(B, n_atoms, 3, generator=g)
    m = torch.rand(B, n_atoms, generator=g) < 0.6
    a[~m] = float("nan")  # unselected atoms may hold anything
    ad, bd = a.to(DEV).contiguous(), b.to(DEV).contiguous()
    mbuf = torch.zeros(B * n_atoms + 8, dtype=torch.uint8, device=DEV)
    s = torch.cuda.current_stream().cuda_stream
    results = []
    for shift in (0, 1):  # 0: 4-byte aligned mask -> vector kernel; 1: scalar kernel
        mview = mbuf[shift:shift + B * n_atoms]
        mview.copy_(m.reshape(-1).to(torch.uint8))
        rot = torch.full((B, 3, 3), 7.0, device=DEV)
        tr = torch.full((B, 3), 7.0, device=DEV)
        _cabi.check(native_lib.ps_kabsch(ad.data_ptr(), bd.data_ptr(), mview.data_ptr(), B, B, n_atoms, rot.data_ptr(),
                                         tr.data_ptr(), s), "ps_kabsch")
        torch.cuda.synchronize()
        results.append((rot.cpu().double(), tr.cpu().double()))
    for rot, tr in results:
        for k in range(B):
            r_ref, t_ref = _kabsch_fp64(torch.nan_to_num(a[k]), b[k], m[k])
            assert (rot[k] - r_ref).abs().max().item() < 2e-6, (k, (rot[k] - r_ref).abs().max().item())
            assert (tr[k] - t_ref).abs().max().item() < 5e-4, (k, (tr[k] - t_ref).abs().max().item())
    assert (results[0][0] - results[1][0]).abs().max().item() < 1e-6
    assert (results[0][1] - results[1][1]).abs().max().item() < 2e-4


def test_packed_angle_kernel_register_budgets_and_unrolled_any_a_kernels_give_the_same_bits(native_lib):
    """include/protstruc_b200.h: every packed variant of ps_trrosetta_angles_ex (3 / 4 / 5 / 6 CTAs per SM) produces
    the same bits; the unrolled A = 25 / 37 instantiations of the any-A tile kernel equal the run-time-A instantiation."""
    s = torch.cuda.current_stream().cuda_stream
    same = lambda p, q: torch.equal(p.view(torch.int32), q.view(torch.int32))  # noqa: E731
    for (B, L, A) in ((3, 100, 5), (2, 77, 15), (1, 600, 5)):
        xyz, _, _ = H.synthetic_batch(4200 + L, B, L, A, "bool")
        x = xyz.to(DEV).contiguous()
        outs = {}
        for variant in (0, 3, 4, 5, 6):
            o = torch.empty(3, B, L, L, device=DEV)
            _cabi.check(native_lib.ps_trrosetta_angles_ex(x.data_ptr(), B, L, A, 0, o[0].data_ptr(), o[1].data_ptr(),
                                                          o[2].data_ptr(), variant, s), "ps_trrosetta_angles_ex")
            outs[variant] = o
        torch.cuda.synchronize()
        for variant in (3, 4, 5, 6):
            assert same(outs[variant], outs[0]), (B, L, A, variant)
    force = 1 << 8
    for (B, L, A) in ((2, 40, 25), (2, 23, 37)):
        xyz, mask, _ = H.synthetic_batch(4300 + A, B, L, A, "bool")
        x, m = xyz.to(DEV).contiguous(), mask.to(DEV).contiguous()
        outs = []
        for variant in (force, force | (13 << 24)):
            d = torch.empty(B, L, L, A, A, device=DEV)
            dm = torch.empty(B, L, L, A, A, dtype=torch.bool, device=DEV)
            _cabi.check(native_lib.ps_pair_dist_mask_ex(x.data_ptr(), m.data_ptr(), 0, d.data_ptr(), dm.data_ptr(), B, L, A,
                                                        variant, s), "ps_pair_dist_mask_ex")
            outs.append((d, dm))
        torch.cuda.synchronize()
        assert same(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (B, L, A)
        rd, rm = orc.pair_distances(xyz, mask)
        H.assert_distances_close(outs[0][0], rd)
        assert torch.equal(outs[0][1].cpu(), rm)


@pytest.mark.parametrize("B,L,A", [(3, 8, 15), (3, 7, 15), (2, 33, 4), (5, 128, 15), (1, 4, 1)])
def test_elementwise_maps_vector_and_scalar_paths_are_bit_exact(native_lib, B, L, A):
    """Round 2: scale_shift / translate / rotate take 128-bit kernels when a structure is a multiple of 4 floats (atoms)
    on 16-byte aligned arrays and the scalar kernels otherwise (a pointer offset by one float forces them).  Both must
    return the reference's separately rounded  x * scale + shift,  x + t  bit for bit (protstruc.py:736-744, 759-788);
    rotate within a few ulp of R x (its dot products are not contracted)."""
    g = torch.Generator().manual_seed(100 * B + L + A)
    n = B * L * A * 3
    x = 20.0 * torch.randn(B, L, A, 3, generator=g)
    x[0, 0, 0, 1] = float("nan")
    scale, shift = torch.rand(B, 3, generator=g) + 0.5, 10.0 * torch.randn(B, 3, generator=g)
    q, _ = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))
    s = torch.cuda.current_stream().cuda_stream
    sc_d, sh_d, rot_d = scale.to(DEV), shift.to(DEV), q.contiguous().to(DEV)
    for offset in (0, 1):  # floats: 0 = 16-byte aligned (torch allocations are), 1 = only 4-byte aligned
        xin = torch.zeros(n + 8, device=DEV)
        xin[offset:offset + n] = x.reshape(-1).to(DEV)
        out = torch.full((n + 8,), -7.0, device=DEV)
        xp, op = xin[offset:].data_ptr(), out[offset:].data_ptr()
        view = lambda: out[offset:offset + n].view(B, L, A, 3).cpu()  # noqa: E731
        _cabi.check(native_lib.ps_scale_shift(xp, sc_d.data_ptr(), sh_d.data_ptr(), B, L, A, op, s), "ps_scale_shift")
        ref = x * scale[:, None, None, :] + shift[:, None, None, :]
        assert torch.equal(torch.nan_to_num(view(), nan=-3.0), torch.nan_to_num(ref, nan=-3.0)), ("scale_shift", offset)
        assert bool((out[:offset] == -7.0).all()) and bool((out[offset + n:] == -7.0).all())
        for rows in (B, 1):
            _cabi.check(native_lib.ps_translate(xp, sh_d.data_ptr(), rows, B, L, A, op, s), "ps_translate")
            ref = x + (shift[:, None, None, :] if rows == B else shift[:1, None, None, :])
            assert torch.equal(torch.nan_to_num(view(), nan=-3.0), torch.nan_to_num(ref, nan=-3.0)), ("translate", offset, rows)
        _cabi.check(native_lib.ps_translate(xp, sh_d.data_ptr(), B, B, L, A, xp, s), "ps_translate in place")
        ref = x + shift[:, None, None, :]
        got = xin[offset:offset + n].view(B, L, A, 3).cpu()
        assert torch.equal(torch.nan_to_num(got, nan=-3.0), torch.nan_to_num(ref, nan=-3.0)), ("translate in place", offset)
        xin[offset:offset + n] = x.reshape(-1).to(DEV)
        _cabi.check(native_lib.ps_rotate(xp, rot_d.data_ptr(), B, B, L, A, op, s), "ps_rotate")
        ref = torch.einsum("bij,blaj->blai", q.double(), torch.nan_to_num(x).double())
        finite = ~torch.isnan(x).any(-1)
        assert (view().double() - ref)[finite].abs().max().item() < 2e-5, ("rotate", offset)
        assert bool(torch.isnan(view()[~finite]).any())


def test_topk_nearest_residue_mask_and_select_match_reference_golden(native_lib):
    g = H.load_golden("frames_align_topk")
    real = H.load_golden("real_1a6v_HL")
    sb = ps.StructureBatch.from_xyz(real["xyz"], real["atom_mask"])
    q = H.t(g["topk_query"])
    for k, key in ((32, "ref_topk_k32"), (500, "ref_topk_k500")):
        got = sb.get_topk_nearest_residue_mask(q, k=k)
        assert got.dtype == torch.bool and tuple(got.shape) == (1, 229)
        assert torch.equal(got.cpu(), H.t(g[key]))
    got = sb.get_topk_nearest_residue_mask(q, k=16, mask=H.t(g["topk_extra_mask"]))
    assert torch.equal(got.cpu(), H.t(g["ref_topk_k16_masked"]))
    picked = sb.residue_masked_select(got)
    assert tuple(picked.get_xyz().shape) == (1, 16, 15, 3)
    two = ps.StructureBatch.from_xyz(np.zeros((2, 5, 15, 3), dtype=np.float32), np.ones((2, 5, 15), dtype=bool))
    with pytest.raises(ValueError):
        two.get_topk_nearest_residue_mask(q)
    with pytest.raises(ValueError):
        two.residue_masked_select(torch.ones(2, 5, dtype=torch.bool))


def test_from_pdb_to_features_on_the_device(native_lib):
    """Row f3 end to end: PDB text -> native ingest -> device -> kernels, against the oracle on the same arrays."""
    path = str(H.GOLDEN / "mini_two_chain.pdb")
    sb = ps.StructureBatch.from_pdb([path, path])
    assert sb.get_xyz().is_cuda and tuple(sb.get_xyz().shape) == (2, 13, 15, 3)
    xyz, mask, chain_idx = sb.get_xyz().cpu(), sb.get_atom_mask().cpu(), sb.chain_idx.cpu()
    dih, dmask = sb.backbone_dihedrals()
    ref, ref_mask = orc.backbone_dihedrals(xyz, chain_idx, mask.any(-1))
    assert torch.equal(dmask.cpu(), ref_mask)
    H.assert_same_nan(dih, ref, "dihedrals from pdb")
    assert H.circular_diff(torch.nan_to_num(dih.cpu()), torch.nan_to_num(ref)).max().item() < 2e-6
    dist, dist_mask = sb.pairwise_distance_matrix()
    rd, rm = orc.pair_distances(xyz, mask)
    H.assert_distances_close(dist, rd)
    assert torch.equal(dist_mask.cpu(), rm)


def test_host_buffer_pipeline_matches_the_device_api(native_lib):
    """C-ABI ps_host_inter_residue_geometry: host arrays in, host arrays out, chunked (ragged last chunk)."""
    from protstruc_b200.host_pipeline import HostFeaturePipeline

    B, L, A = 5, 40, 15
    xyz, mask, _ = H.synthetic_batch(31, B, L, A, "bool")
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    ref = sb.inter_residue_geometry()
    dist, dist_mask = sb.pairwise_distance_matrix()
    pipe = HostFeaturePipeline(chunk=2, L=L, A=A)
    for pinned in (True, False):
        out = HostFeaturePipeline.allocate_host_outputs(B, L, A, pinned=pinned)
        x_h = xyz.pin_memory() if pinned else xyz.clone()
        m_h = mask.pin_memory() if pinned else mask.clone()
        pipe.run(x_h, m_h, out)
        assert not out["dist"].is_cuda
        assert torch.equal(torch.nan_to_num(out["dist"], nan=-1.0), torch.nan_to_num(dist.cpu(), nan=-1.0))
        assert torch.equal(out["dist_mask"], dist_mask.cpu())
        for k in ("omega", "theta", "phi"):
            assert torch.equal(torch.nan_to_num(out[k], nan=-9.0), torch.nan_to_num(ref[k].cpu(), nan=-9.0)), k
    assert pipe.launches == 2 * 3  # ceil(5 / 2) chunks per run
    with pytest.raises(ValueError):
        pipe.run(xyz.cuda(), mask, out)
    with pytest.raises(ValueError):  # right number of rows, wrong trailing shape: refused before anything is written
        pipe.run(xyz, mask, dict(out, omega=torch.empty(B, L, L - 1)))
    pipe.close()


def test_config5_shape_chunked_streaming_is_bit_identical(native_lib):
    """BASELINE config 5 (L = 384, A = 15) is streamed through the GPU in batch chunks; chunking must not change a bit,
    and shards computed independently (as on 2 or 8 GPUs) must reassemble to the unsharded result."""
    from protstruc_b200.sharding import shard_bounds

    B, L, A = 6, 384, 15
    g = torch.Generator(device=DEV).manual_seed(5)
    xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
    mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
    xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
    whole = ps.StructureBatch.from_xyz(xyz, mask).inter_residue_geometry()
    dist_w, mask_w = ps.StructureBatch.from_xyz(xyz, mask).pairwise_distance_matrix()
    for world in (2, 4):
        parts = []
        for rank in range(world):
            a, b = shard_bounds(B, world, rank)
            sb = ps.StructureBatch.from_xyz(xyz[a:b], mask[a:b])
            parts.append((sb.inter_residue_geometry(), sb.pairwise_distance_matrix()))
        for key in ("omega", "theta", "phi", "d_ca", "d_no_mask"):
            joined = torch.cat([p[0][key] for p in parts])
            assert torch.equal(torch.nan_to_num(joined.float(), nan=-7.0), torch.nan_to_num(whole[key].float(), nan=-7.0)), key
        assert torch.equal(torch.nan_to_num(torch.cat([p[1][0] for p in parts]), nan=-7.0), torch.nan_to_num(dist_w, nan=-7.0))
        assert torch.equal(torch.cat([p[1][1] for p in parts]), mask_w)
    # exact structure of the output: symmetric, zero diagonal, NaN exactly where an atom is missing
    d0 = dist_w[0]
    assert torch.equal(torch.nan_to_num(d0, nan=-7.0), torch.nan_to_num(d0.permute(1, 0, 3, 2), nan=-7.0))
    expect_nan = ~(mask[0][:, None, :, None] & mask[0][None, :, None, :])
    assert torch.equal(torch.isnan(d0), expect_nan)


@pytest.mark.parametrize("B,L", [(1, 32), (1, 33), (3, 37), (2, 45), (1, 229), (7, 40)])
def test_fused_kernel_never_writes_outside_its_outputs(native_lib, B, L):
    """Guard bands around every output of the fused kernel (tail tiles, odd L, ragged strips): the sentinel bytes
    before and after each buffer must survive, and the payload must equal the unguarded call."""
    A, guard = 15, 4096
    xyz, mask, _ = H.synthetic_batch(B * 100 + L, B, L, A, "bool")
    x, m = xyz.to(DEV), mask.to(DEV)
    n_pairs, n_elems = B * L * L, B * L * L * A * A

    def guarded(nbytes):
        buf = torch.full((guard + nbytes + guard,), 0xA5, dtype=torch.uint8, device=DEV)
        return buf, buf[guard:guard + nbytes]

    bufs = {name: guarded(n) for name, n in (("dist", n_elems * 4), ("mask", n_elems), ("omega", n_pairs * 4),
                                             ("theta", n_pairs * 4), ("phi", n_pairs * 4))}
    ptr = {k: v[1].data_ptr() for k, v in bufs.items()}
    assert all(p % 16 == 0 for p in ptr.values())
    rc = native_lib.ps_inter_residue_geometry(x.data_ptr(), m.data_ptr(), 0, ptr["dist"], ptr["mask"], ptr["omega"],
                                              ptr["theta"], ptr["phi"], B, L, A, torch.cuda.current_stream().cuda_stream)
    _cabi.check(rc, "ps_inter_residue_geometry")
    torch.cuda.synchronize()
    for name, (buf, view) in bufs.items():
        assert bool((buf[:guard] == 0xA5).all()) and bool((buf[-guard:] == 0xA5).all()), f"{name}: guard band overwritten"
    ref = ps.StructureBatch.from_xyz(xyz, mask).inter_residue_geometry()
    dist_ref, mask_ref = ps.StructureBatch.from_xyz(xyz, mask).pairwise_distance_matrix()
    got_dist = bufs["dist"][1].view(torch.float32).view(B, L, L, A, A)
    assert torch.equal(torch.nan_to_num(got_dist, nan=-3.0), torch.nan_to_num(dist_ref, nan=-3.0))
    assert torch.equal(bufs["mask"][1].view(torch.bool).view(B, L, L, A, A), mask_ref)
    for k in ("omega", "theta", "phi"):
        got = bufs[k][1].view(torch.float32).view(B, L, L)
        assert torch.equal(torch.nan_to_num(got, nan=-3.0), torch.nan_to_num(ref[k], nan=-3.0)), k


@pytest.mark.parametrize("B,L,A", [(2, 48, 5), (1, 37, 14), (2, 33, 10), (2, 140, 5), (1, 129, 5), (2, 70, 10)])
def test_staged_kernels_for_other_atom_counts(native_lib, B, L, A):
    """A = 5 (backbone + CB), 10 and 14 (atom14) run the staged TMA-store kernel too: fused features vs the oracle,
    bit-identical to the generic kernel, guard bands intact."""
    xyz, mask, _ = H.synthetic_batch(500 + A, B, L, A, "bool")
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    out = sb.inter_residue_geometry()
    ro, rt, rp = orc.trrosetta_angles(xyz)
    H.assert_angles_close(out["omega"], ro, angle_conditioning(xyz, "omega"), "omega", all_finite_tol=2e-6)
    H.assert_angles_close(out["theta"], rt, angle_conditioning(xyz, "theta"), "theta", all_finite_tol=2e-6)
    H.assert_angles_close(out["phi"], rp, angle_conditioning(xyz, "phi"), "phi", circular=False)
    rd, rm = orc.pair_distances(xyz, mask)
    H.assert_distances_close(out["d_ca"], rd[:, :, :, 1, 1], "d_ca")
    assert torch.equal(out["d_no_mask"].cpu(), rm[:, :, :, 0, 3])
    x, m = xyz.to(DEV), mask.to(DEV)
    n = B * L * L * A * A
    results = []
    for variant in (0, 1 << 8):  # staged, generic
        guard = 1024
        dbuf = torch.full((guard + n * 4 + guard,), 0x5A, dtype=torch.uint8, device=DEV)
        mbuf = torch.full((guard + n + guard,), 0x5A, dtype=torch.uint8, device=DEV)
        rc = native_lib.ps_pair_dist_mask_ex(x.data_ptr(), m.data_ptr(), 0, dbuf[guard:].data_ptr(), mbuf[guard:].data_ptr(),
                                             B, L, A, variant, torch.cuda.current_stream().cuda_stream)
        _cabi.check(rc, "ps_pair_dist_mask_ex")
        torch.cuda.synchronize()
        for buf in (dbuf, mbuf):
            assert bool((buf[:guard] == 0x5A).all()) and bool((buf[-guard:] == 0x5A).all())
        results.append((dbuf[guard:guard + n * 4].view(torch.float32).clone(), mbuf[guard:guard + n].clone()))
    assert torch.equal(torch.nan_to_num(results[0][0], nan=-2.0), torch.nan_to_num(results[1][0], nan=-2.0))
    assert torch.equal(results[0][1], results[1][1])
    H.assert_distances_close(results[0][0].view(B, L, L, A, A), rd)
    assert torch.equal(results[0][1].view(torch.bool).view(B, L, L, A, A).cpu(), rm)


@pytest.mark.parametrize("B,L,A", [(2, 20, 25), (1, 9, 37), (3, 50, 4), (2, 64, 3), (5, 3, 6), (1, 1, 7), (2, 33, 16),
                                   (1, 130, 8), (2, 2, 1), (1, 7, 53), (1, 40, 15), (2, 17, 15), (1, 130, 3), (2, 128, 4),
                                   (1, 141, 4), (1, 12, 130), (1, 5, 64), (1, 4, 100), (1, 3, 128), (1, 3, 129)])
def test_any_shape_distance_kernels(native_lib, B, L, A):
    """Atom counts / lengths the staged kernel does not cover: the any-A tile kernel (TMA bulk stores), its
    plain-store flavour for outputs that are not 16-B aligned, and the row kernel (variant bit 12) — every output
    kind, bit-identical to each other, within the distance tolerance of the oracle, guard bands intact."""
    xyz, mask, _ = H.synthetic_batch(900 + 7 * A + L, B, L, A, "bool")
    rd, rm = orc.pair_distances(xyz, mask)
    fmask = mask.float() * torch.rand(B, L, A, generator=torch.Generator().manual_seed(A))  # non-trivial fp32 mask
    _, rfm = orc.pair_distances(xyz, fmask)
    x, m, fm = xyz.to(DEV), mask.to(DEV), fmask.to(DEV).contiguous()
    n = B * L * L * A * A
    s = torch.cuda.current_stream().cuda_stream
    guard = 256

    def run(variant, shift, mask_dtype, want_dist, want_mask):
        """shift = byte offset of the outputs inside 256-B aligned buffers (4 = only fp32-aligned)."""
        item = 4 if mask_dtype == 1 else 1
        dbuf = torch.full((guard + n * 4 + guard,), 0x3C, dtype=torch.uint8, device=DEV)
        mbuf = torch.full((guard + n * item + guard,), 0x3C, dtype=torch.uint8, device=DEV)
        dptr = dbuf[guard + shift - 16:].data_ptr() if want_dist else 0
        mshift = shift if mask_dtype == 1 else (shift + 1 if shift else 0)  # byte masks may sit on any address
        mptr = mbuf[guard + mshift - 16:].data_ptr() if want_mask else 0
        am = (fm if mask_dtype == 1 else m).data_ptr() if want_mask else 0
        rc = native_lib.ps_pair_dist_mask_ex(x.data_ptr(), am, mask_dtype, dptr, mptr, B, L, A, variant, s)
        _cabi.check(rc, "ps_pair_dist_mask_ex")
        torch.cuda.synchronize()
        d0, m0 = guard + shift - 16, guard + mshift - 16
        out = {}
        if want_dist:
            assert bool((dbuf[:d0] == 0x3C).all()) and bool((dbuf[d0 + n * 4:] == 0x3C).all()), "dist guard band"
            out["dist"] = dbuf[d0:d0 + n * 4].clone().view(torch.float32).view(B, L, L, A, A)
        else:
            assert bool((dbuf == 0x3C).all())
        if want_mask:
            assert bool((mbuf[:m0] == 0x3C).all()) and bool((mbuf[m0 + n * item:] == 0x3C).all()), "mask guard band"
            raw = mbuf[m0:m0 + n * item].clone()
            out["mask"] = (raw.view(torch.float32) if mask_dtype == 1 else raw.view(torch.bool)).view(B, L, L, A, A)
        else:
            assert bool((mbuf == 0x3C).all())
        return out

    force = 1 << 8  # keep the staged kernel out even for A = 15
    base = run(force, 16, 0, True, True)
    H.assert_distances_close(base["dist"], rd)
    assert torch.equal(base["mask"].cpu(), rm)
    same = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-2.0), torch.nan_to_num(b, nan=-2.0))  # noqa: E731
    for variant in (force, force | (1 << 12)):
        for shift in (16, 20):
            both = run(variant, shift, 0, True, True)
            assert same(both["dist"], base["dist"]) and torch.equal(both["mask"], base["mask"]), (variant, shift)
            assert same(run(variant, shift, 0, True, False)["dist"], base["dist"]), (variant, shift)
            assert torch.equal(run(variant, shift, 0, False, True)["mask"], base["mask"]), (variant, shift)
            f = run(variant, shift, 1, True, True)
            assert same(f["dist"], base["dist"]) and torch.equal(f["mask"].cpu(), rfm), (variant, shift)
            assert torch.equal(run(variant, shift, 1, False, True)["mask"].cpu(), rfm), (variant, shift)
    # whatever the default dispatch picks for this shape (staged kernel for A = 3 / 4 / 15 when L allows) agrees
    for mask_dtype in (0, 1):
        d = run(0, 16, mask_dtype, True, True)
        assert same(d["dist"], base["dist"])
        assert torch.equal(d["mask"].cpu(), rfm if mask_dtype else rm)
    assert torch.equal(run(0, 16, 0, False, True)["mask"], base["mask"])
    # IEEE square root flavour through the same kernels
    ieee = run(force | 2, 16, 0, True, False)["dist"]
    H.assert_distances_close(ieee, rd)


def test_random_shapes_of_the_distance_tensor_against_a_device_side_check(native_lib):
    """60 seeded random (B, L, A) — tile boundaries fall everywhere relative to residue rows and structures — through
    the default dispatch and the any-A tile kernel with forced small / large tiles.  The check is evaluated with
    torch on the device (same subtraction, fp32 sum of squares in the kernel's order is within 1 ulp of it), masks
    exact; the any-A tile kernel must be bit-identical across tile sizes."""
    rng = np.random.default_rng(77)
    s = torch.cuda.current_stream().cuda_stream
    for case in range(60):
        A = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 9, 11, 12, 13, 15, 16, 17, 21, 25, 30, 37, 40]))
        L = int(rng.integers(1, 70 if A > 20 else 150))
        B = int(rng.integers(1, 5))
        g = torch.Generator(device=DEV).manual_seed(1000 + case)
        xyz = (20.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)).contiguous()
        mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.6
        diff = xyz[:, :, None, :, None, :] - xyz[:, None, :, None, :, :]
        want = torch.sqrt((diff * diff).sum(-1))
        want_mask = mask[:, :, None, :, None] & mask[:, None, :, None, :]
        outs = []
        for variant in (0, 1 << 8, (1 << 8) | (1 << 16) | (1 << 24), (1 << 8) | (40 << 16) | (1 << 25)):
            d = torch.full((B, L, L, A, A), -1.0, device=DEV)
            m = torch.zeros(B, L, L, A, A, dtype=torch.bool, device=DEV)
            rc = native_lib.ps_pair_dist_mask_ex(xyz.data_ptr(), mask.data_ptr(), 0, d.data_ptr(), m.data_ptr(), B, L, A,
                                                 variant, s)
            _cabi.check(rc, "ps_pair_dist_mask_ex")
            tag = f"case {case}: B={B} L={L} A={A} variant={variant:#x}"
            assert torch.equal(m, want_mask), tag
            err = (d - want).abs()
            assert bool((err <= 4e-7 * want + 1e-30).all()), f"{tag}: max err {err.max().item()}"
            outs.append(d)
        assert torch.equal(outs[1], outs[2]) and torch.equal(outs[1], outs[3]), f"case {case}: tile size changed the result"
        assert torch.equal(outs[0], outs[1]), f"case {case}: default dispatch differs from the any-A tile kernel"


@pytest.mark.parametrize("A", [15, 16])
def test_more_than_2_31_output_elements(native_lib, A):
    """One structure of 3100 residues: 2.16 G (A = 15, staged kernel) / 2.46 G (A = 16, any-A tile kernel) output
    elements, i.e. element offsets beyond 32 bits.  Sampled residue rows (first, last, around the 2^31-element
    boundary) are checked against a device-side evaluation; the mask exactly."""
    B, L = 1, 3100
    g = torch.Generator(device=DEV).manual_seed(31)
    xyz = (30.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)).contiguous()
    mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    dist, dist_mask = sb.pairwise_distance_matrix()
    assert dist.numel() > 2 ** 31
    boundary_row = (2 ** 31) // (L * A * A)
    rows = sorted({0, 1, 17, boundary_row - 1, boundary_row, boundary_row + 1, L // 2, L - 2, L - 1})
    for i in rows:
        diff = xyz[0, i, None, :, None, :] - xyz[0, :, None, :, :]          # (L, A, A, 3)
        want = torch.sqrt((diff * diff).sum(-1))
        got = dist[0, i]
        assert bool(((got - want).abs() <= 4e-7 * want + 1e-30).all()), f"row {i}"
        assert torch.equal(dist_mask[0, i], mask[0, i, None, :, None] & mask[0, :, None, :]), f"mask row {i}"
    # every element was written (the buffer is torch.empty): the diagonal blocks hold exact zeros, nothing is NaN
    assert bool((dist[0, torch.arange(L), torch.arange(L)].diagonal(dim1=-2, dim2=-1) == 0).all())
    assert not bool(torch.isnan(dist[0, ::97]).any())
    del dist, dist_mask


def test_feature_calls_can_be_captured_in_a_cuda_graph(native_lib):
    """The launches take the caller's stream, allocate nothing and never synchronise, so a sequence of feature calls
    can be captured once in a CUDA graph and replayed on new coordinates written into the same input buffer (what a
    latency-sensitive caller does for small structures): results equal the eager calls bit for bit."""
    B, L, A = 3, 77, 15
    xyz, mask, chain_idx = H.synthetic_batch(41, B, L, A, "bool")
    sb = ps.StructureBatch.from_xyz(xyz, mask, chain_idx, [["A", "B"]] * B)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):  # warm-up outside the capture
        sb.inter_residue_geometry()
        sb.backbone_dihedrals()
        sb.get_local_xyz()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        feats = sb.inter_residue_geometry()
        dihedrals, dihedral_mask = sb.backbone_dihedrals()
        local = sb.get_local_xyz()
    new_xyz, _, _ = H.synthetic_batch(42, B, L, A, "bool")
    for data in (new_xyz, xyz):
        sb.get_xyz().copy_(data.to(DEV))
        graph.replay()
        torch.cuda.synchronize()
        eager = ps.StructureBatch.from_xyz(data, mask, chain_idx, [["A", "B"]] * B)
        want = eager.inter_residue_geometry()
        same = lambda a, b: torch.equal(torch.nan_to_num(a, nan=-5.0), torch.nan_to_num(b, nan=-5.0))  # noqa: E731
        for k in ("omega", "theta", "phi", "d_ca", "d_cb", "d_no"):
            assert same(feats[k], want[k]), k
        assert torch.equal(feats["d_no_mask"], want["d_no_mask"])
        assert same(dihedrals, eager.backbone_dihedrals()[0]) and same(local, eager.get_local_xyz())


def test_randomised_shapes_against_the_oracle(native_lib):
    """Seeded sweep over odd shapes (tail tiles, L around the 32-pair tile size, every staged atom count and the
    generic path, bool and float masks, ragged lengths): every feature family vs the CPU oracle."""
    rng = np.random.default_rng(2024)
    shapes = [(1, 31, 15), (1, 32, 15), (2, 34, 15), (3, 63, 15), (2, 65, 15), (1, 96, 15), (4, 50, 5), (2, 41, 10),
              (1, 77, 14), (2, 20, 15), (3, 9, 7), (1, 36, 12), (2, 127, 5), (2, 133, 5), (3, 66, 10), (1, 200, 5)]
    for idx, (B, L, A) in enumerate(shapes):
        kind = "float" if idx % 3 == 2 else "bool"
        xyz, mask, chain_idx = H.synthetic_batch(int(rng.integers(1 << 30)), B, L, A, kind)
        ids = [["A", "B"]] * B
        sb = ps.StructureBatch.from_xyz(xyz, mask, chain_idx, ids)
        tag = f"B={B} L={L} A={A} {kind}"
        dist, dist_mask = sb.pairwise_distance_matrix()
        rd, rm = orc.pair_distances(xyz, mask)
        H.assert_distances_close(dist, rd, f"dist {tag}")
        assert dist_mask.dtype == rm.dtype and torch.equal(dist_mask.cpu(), rm), tag
        if A >= 5:
            out = sb.inter_residue_geometry()
            ro, rt, rp = orc.trrosetta_angles(xyz)
            H.assert_angles_close(out["omega"], ro, angle_conditioning(xyz, "omega"), f"omega {tag}", all_finite_tol=2e-6)
            H.assert_angles_close(out["theta"], rt, angle_conditioning(xyz, "theta"), f"theta {tag}", all_finite_tol=2e-6)
            H.assert_angles_close(out["phi"], rp, angle_conditioning(xyz, "phi"), f"phi {tag}", circular=False)
            H.assert_distances_close(out["d_cb"], rd[:, :, :, 4, 4], f"d_cb {tag}")
        dih, dmask = sb.backbone_dihedrals()
        rdi, rdm = orc.backbone_dihedrals(xyz, chain_idx, mask.bool().any(-1))
        assert torch.equal(dmask.cpu(), rdm), tag
        H.assert_same_nan(dih, rdi, f"dihedrals {tag}")
        assert H.circular_diff(torch.nan_to_num(dih.cpu()), torch.nan_to_num(rdi)).max().item() <= 2e-6, tag
        com = sb.center_of_mass()
        assert torch.allclose(com.cpu(), orc.center_of_mass(xyz), rtol=1e-5, atol=1e-4, equal_nan=True), tag
        sb.standardize()
        rx, mu, sd = orc.standardize_per_structure(xyz, mask)
        assert torch.allclose(sb.mu.cpu(), mu, rtol=1e-5, atol=1e-5, equal_nan=True), tag
        assert torch.allclose(sb.std.cpu(), sd, rtol=1e-5, atol=1e-6, equal_nan=True), tag
        assert torch.allclose(sb.get_xyz().cpu(), rx, rtol=1e-4, atol=1e-5, equal_nan=True), tag


def test_exact_symmetries_at_baseline_config3_size(native_lib):
    """BASELINE config 3 at full size (256 x 512 backbone slots, 67 M pairs): properties that hold bit for bit.
    Mirroring x negates every cross product exactly, so dihedrals change sign and planar angles / distances do not;
    reversing the residue order permutes the pair axes."""
    B, L, A = 256, 512, 5
    g = torch.Generator(device=DEV).manual_seed(3)
    xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
    mask = torch.ones(B, L, A, dtype=torch.bool, device=DEV)
    sb = ps.StructureBatch.from_xyz(xyz, mask)
    omega, theta, phi = sb.trrosetta_angles()
    assert bool(torch.isfinite(omega).all()) and bool((omega.abs() <= math.pi).all())
    assert bool((phi[~torch.isnan(phi)] >= 0).all()) and bool((phi[~torch.isnan(phi)] <= math.pi).all())
    eye = torch.eye(L, dtype=torch.bool, device=DEV)
    assert bool((omega[:, eye] == 0).all()) and bool((theta[:, eye] == 0).all()) and bool(torch.isnan(phi[:, eye]).all())
    mirrored = xyz * torch.tensor([-1.0, 1.0, 1.0], device=DEV)
    mo, mt, mp = ps.StructureBatch.from_xyz(mirrored, mask).trrosetta_angles()
    # ... except on the branch cut: a sine that cancels to +0 stays +0 under the mirror (a - a = +0 either way), so
    # atan2(+0, x < 0) = +pi in both; only such entries (|angle| == fp32 pi, a handful in 67 M) may keep their sign
    pi32 = torch.tensor(math.pi, dtype=torch.float32, device=DEV)
    for mirrored_angle, angle in ((mo, omega), (mt, theta)):
        off = mirrored_angle != -angle
        assert int(off.sum()) <= 8 and bool((angle[off].abs() == pi32).all())
        assert torch.equal(mirrored_angle[off], angle[off])
    assert torch.equal(torch.nan_to_num(mp, nan=-1.0), torch.nan_to_num(phi, nan=-1.0))
    flipped = torch.flip(xyz, dims=[1])
    fo, ft, fp = ps.StructureBatch.from_xyz(flipped, mask).trrosetta_angles()
    assert torch.equal(torch.flip(fo, dims=[1, 2]), omega) and torch.equal(torch.flip(ft, dims=[1, 2]), theta)
    assert torch.equal(torch.nan_to_num(torch.flip(fp, dims=[1, 2]), nan=-1.0), torch.nan_to_num(phi, nan=-1.0))
    # the generic single-feature kernels agree with the fused one (a few ulp) at this size as well
    go = sb.pairwise_dihedrals(["CA", "CB"], ["CA", "CB"])[:2].cpu()
    cond = angle_conditioning(xyz[:2].cpu(), "omega").reshape(go.shape)
    assert torch.equal(torch.isnan(go), torch.isnan(omega[:2].cpu()))
    diff = H.circular_diff(torch.nan_to_num(go), torch.nan_to_num(omega[:2].cpu()))
    assert diff[cond >= H.SIN_GATE].max().item() <= 2e-5, "packed K2f vs packed generic kernel (each within 1e-5 of the truth)"
    del mo, mt, mp, fo, ft, fp, go


def test_exact_symmetries_of_the_distance_tensor_under_mirror_and_reversal(native_lib):
    B, L, A = 8, 256, 15
    g = torch.Generator(device=DEV).manual_seed(22)
    xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
    mask = torch.rand(B, L, A, device=DEV, generator=g) < 0.5
    d, m = ps.StructureBatch.from_xyz(xyz, mask).pairwise_distance_matrix()
    dm, mm = ps.StructureBatch.from_xyz(xyz * torch.tensor([1.0, -1.0, 1.0], device=DEV), mask).pairwise_distance_matrix()
    assert torch.equal(dm, d) and torch.equal(mm, m)
    df, mf = ps.StructureBatch.from_xyz(torch.flip(xyz, dims=[1]), torch.flip(mask, dims=[1])).pairwise_distance_matrix()
    assert torch.equal(torch.flip(df, dims=[1, 2]), d) and torch.equal(torch.flip(mf, dims=[1, 2]), m)
    # translation by a power of two that keeps every coordinate exactly representable changes nothing either
    shifted = (xyz.double() + 64.0).float()
    exact = (shifted.double() - 64.0).float() == xyz
    if bool(exact.all()):
        ds, _ = ps.StructureBatch.from_xyz(shifted, mask).pairwise_distance_matrix()
        assert torch.equal(ds, d)
