"""Comparison helpers shared by the parity tests (tolerances are BASELINE.json's north_star)."""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"

# north_star tolerances (fp32)
DIST_REL = 1e-5      # relative
DIST_ABS = 1e-4      # Angstrom
ANGLE_TOL = 1e-5     # rad, away from collinear degeneracies (min sin(bond angle) >= 0.1)
SIN_GATE = 0.1


def load_golden(name: str) -> dict:
    with np.load(GOLDEN / f"{name}.npz", allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def t(x, device="cpu"):
    return torch.as_tensor(np.asarray(x)).to(device)


def assert_same_nan(actual: torch.Tensor, expected: torch.Tensor, what: str):
    a, e = torch.isnan(actual.cpu()), torch.isnan(expected.cpu())
    assert a.shape == e.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(e.shape)}"
    assert torch.equal(a, e), f"{what}: NaN placement differs at {(a != e).sum().item()} positions"


def assert_distances_close(actual: torch.Tensor, expected: torch.Tensor, what: str = "dist"):
    """|d - ref| <= max(1e-4 A, 1e-5 * |ref|), NaN placement and exact zeros bit-exact."""
    actual, expected = actual.cpu(), expected.cpu()
    assert_same_nan(actual, expected, what)
    ok = ~torch.isnan(expected)
    a, e = actual[ok].double(), expected[ok].double()
    inf = torch.isinf(e)
    assert torch.equal(a[inf], e[inf]), f"{what}: infinities differ"
    a, e = a[~inf], e[~inf]
    err = (a - e).abs()
    bound = torch.maximum(torch.full_like(e, DIST_ABS), DIST_REL * e.abs())
    worst = (err / bound).max().item() if err.numel() else 0.0
    assert worst <= 1.0, f"{what}: error/bound = {worst:.3g} (max abs err {err.max().item():.3g})"
    assert torch.equal(a == 0, e == 0), f"{what}: exact zeros (diagonal) differ"
    return err.max().item() if err.numel() else 0.0


def circular_diff(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    d = (a.double() - b.double()).abs()
    return torch.minimum(d, 2 * math.pi - d)


def _sin_between(u: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    cr = torch.linalg.cross(u.double(), v.double(), dim=-1).norm(dim=-1)
    return cr / (u.double().norm(dim=-1) * v.double().norm(dim=-1))


def dihedral_conditioning(a, b, c, d) -> torch.Tensor:
    """min sin of the two bond angles (a-b-c) and (b-c-d); 0/NaN for degenerate quadruples."""
    return torch.minimum(_sin_between(a - b, c - b), _sin_between(b - c, d - c))


def planar_conditioning(a, b, c) -> torch.Tensor:
    return _sin_between(a - b, c - b)


def assert_angles_close(actual: torch.Tensor, expected: torch.Tensor, cond: torch.Tensor, what: str,
                        circular: bool = True, well_tol: float = ANGLE_TOL, all_finite_tol: float = None):
    """<= 1e-5 rad where the geometry is well conditioned (min sin >= 0.1), <= 1e-6 / sin below that
    (SURVEY 8c), NaN placement bit-exact; degenerate entries (sin == 0 or NaN) must agree in NaN-ness
    and are otherwise exempt from the value check unless both are finite and `sin` is NaN-free."""
    actual, expected, cond = actual.cpu(), expected.cpu(), cond.cpu().reshape(expected.shape)
    assert_same_nan(actual, expected, what)
    finite = ~torch.isnan(expected)
    diff = circular_diff(actual, expected) if circular else (actual.double() - expected.double()).abs()
    good = finite & (cond >= SIN_GATE)
    if good.any():
        worst = diff[good].max().item()
        assert worst <= well_tol, f"{what}: max deviation {worst:.3g} rad in the well-conditioned region"
    if all_finite_tol is not None and finite.any():
        # dihedrals: the kernels issue the reference's exact fp32 op sequence up to the final sqrt / div /
        # atan2, so even degenerate and padded entries must agree to a few ulp of pi
        worst_all = diff[finite].max().item()
        assert worst_all <= all_finite_tol, f"{what}: max deviation {worst_all:.3g} rad over all finite entries"
    soft = finite & (cond > 0) & (cond < SIN_GATE)
    if soft.any():
        ratio = (diff[soft] / (1e-6 / cond[soft])).max().item()
        assert ratio <= 1.0, f"{what}: deviation/bound = {ratio:.3g} in the ill-conditioned region"
    return diff[good].max().item() if good.any() else 0.0


def pair_points(xyz: torch.Tensor, slots_i, slots_j):
    """(B, L, L, n, 3) float64 points of every residue pair (test-side helper, independent of the oracle)."""
    B, L = xyz.shape[:2]
    pi = xyz[:, :, list(slots_i)].double()[:, :, None].expand(B, L, L, len(slots_i), 3)
    pj = xyz[:, :, list(slots_j)].double()[:, None, :].expand(B, L, L, len(slots_j), 3)
    return torch.cat([pi, pj], dim=-2)


def trrosetta_conditioning(xyz: torch.Tensor, which: str) -> torch.Tensor:
    """min sin(bond angle) of every residue pair for omega / theta / phi as the reference defines them
    (protstruc.py:810-815): the gate of the 1e-5 rad tolerance ("away from collinear degeneracies")."""
    N, CA, CB = 0, 1, 4
    if which == "omega":
        p = pair_points(xyz, [CA, CB], [CA, CB])
        return dihedral_conditioning(p[..., 0, :], p[..., 1, :], p[..., 2, :], p[..., 3, :])
    if which == "theta":
        p = pair_points(xyz, [N, CA, CB], [CB])
        return dihedral_conditioning(p[..., 0, :], p[..., 1, :], p[..., 2, :], p[..., 3, :])
    p = pair_points(xyz, [CA, CB], [CB])
    return planar_conditioning(p[..., 0, :], p[..., 1, :], p[..., 2, :])


def synthetic_batch(seed: int, B: int, L: int, A: int, mask_kind: str = "bool", nan_masked: bool = True,
                    full_length: bool = False):
    """Same generator family as tests/golden/make_golden.py::synthetic_inputs (protein-like walk)."""
    g = torch.Generator().manual_seed(seed)
    steps = torch.randn(B, L, 3, generator=g)
    steps = 3.8 * steps / steps.norm(dim=-1, keepdim=True)
    ca = steps.cumsum(dim=1)
    xyz = ca[:, :, None, :] + 1.5 * torch.randn(B, L, A, 3, generator=g)
    mask = torch.rand(B, L, A, generator=g) < 0.7
    mask[:, :, : min(A, 4)] = True
    if A > 4:
        mask[:, 1::5, 4] = False
    if L > 4:
        mask[0, 3, :] = False
    lengths = [L if full_length else max(1, L - (b % 3)) for b in range(B)]
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
