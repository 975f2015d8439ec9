"""CPU ORACLE of the geometric-feature hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs may
import this module, and only as the checker (or as the timed CPU baseline).  Nothing under
`protstruc_b200/` imports it; the product path has no CPU fallback.

What it is: a restatement, in plain torch-CPU / numpy array operations, of the arithmetic the
reference (dohlee/protstruc, pure Python) performs on the hot path.  The reference's arithmetic *is*
a sequence of third-party array ops (torch 2.11.0 ATen ops and numpy 2.3.5 `np.cross` /
`np.arctan2`; the reference pins no versions, `setup.py:24-33`), so each function below issues the
same op sequence on the same shapes, citing the reference lines it follows.  Because the ops and
their order are identical, the oracle is bit-identical to the reference on CPU; that is verified —
not assumed — by `tests/golden/make_golden.py`, which imports the real reference from
/root/reference in the build container, compares every function here against it on seeded random
and real-PDB inputs, and writes the golden vectors committed under `tests/golden/`.
PARITY STATUS: pinned (reference-generated golden vectors + the reference's own known-answer tests,
see tests/test_oracle_golden.py).

Documented deviations:
  * `standardize_per_structure` applies the reference formula to each structure separately (Q1: the
    reference broadcast is only valid for B == 1);
  * `frames` takes the cross product along the last axis (Q2: `torch.cross` without `dim` picks the
    first size-3 axis, wrong when B == 3 or L == 3);
  * `virtual_cb` restates the only virtual-CB formula in the reference (geometry.py:217-221).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

SLOT = {"N": 0, "CA": 1, "C": 2, "O": 3, "CB": 4}  # reference protstruc/general.py:4-16


# ------------------------------------------------------------------ geometry primitives (a7-a9)
def dot(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """reference protstruc/geometry.py:24-26"""
    return (x * y).sum(dim=-1, keepdim=True)


def norm(x: torch.Tensor) -> torch.Tensor:
    """reference protstruc/geometry.py:29-31"""
    return x.norm(dim=-1, keepdim=True)


def planar_angle(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, to_degree: bool = False) -> torch.Tensor:
    """reference protstruc/geometry.py:64-71 — arccos of the normalised dot product, no clamp."""
    u = a - b
    v = c - b
    cosine = dot(u, v) / (norm(u) * norm(v))
    out = torch.arccos(cosine)
    if to_degree:
        out = torch.rad2deg(out)
    return out.squeeze(-1)


def dihedral(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, d: torch.Tensor,
             to_degree: bool = False) -> torch.Tensor:
    """reference protstruc/geometry.py:110-124.  The three cross products and the arctan2 go
    through numpy exactly as the reference does (fp32 ndarray ops); the dot / norm / division are
    torch ops."""
    b0 = a - b
    b1 = c - b
    b2 = d - c
    n1 = np.cross(b0.numpy(), b1.numpy())
    n2 = np.cross(b2.numpy(), b1.numpy())
    m = np.cross(n1, n2)
    x = dot(torch.from_numpy(n1), torch.from_numpy(n2)).numpy()
    y = dot(torch.from_numpy(m), b1) / norm(b1)
    out = np.arctan2(y.numpy(), x)
    if to_degree:
        out = np.degrees(out)
    return torch.from_numpy(np.ascontiguousarray(out)).squeeze(-1)


def frames_from_points(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """reference protstruc/geometry.py:430-439 (Gram-Schmidt), cross product on the last axis."""
    v1 = c - b
    e1 = v1 / norm(v1)
    v2 = a - b
    u2 = v2 - dot(e1, v2) * e1
    e2 = u2 / norm(u2)
    e3 = torch.cross(e1, e2, dim=-1)
    return torch.stack([e1, e2, e3], dim=-1)


def virtual_cb(n: torch.Tensor, ca: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """reference protstruc/geometry.py:217-221 generalised from the ideal residue to any N, CA, C."""
    vb = ca - n
    vc = c - ca
    va = torch.cross(vb, vc, dim=-1)
    return -0.58273431 * va + 0.56802827 * vb - 0.54067466 * vc + ca


# ------------------------------------------------------------------ pairwise features (a2-a6)
def pair_distances(xyz: torch.Tensor, atom_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference protstruc/protstruc.py:477-484 — broadcast difference, 2-norm over xyz, mask product."""
    diff = xyz[:, :, None, :, None] - xyz[:, None, :, None, :]
    dist = torch.norm(diff, dim=-1)
    pair_mask = atom_mask[:, :, None, :, None] * atom_mask[:, None, :, None, :]
    return dist, pair_mask


def gather_pair_points(xyz: torch.Tensor, slots_i: Sequence[int], slots_j: Sequence[int]) -> torch.Tensor:
    """reference protstruc/protstruc.py:610-618 — row p = i*L + j holds slots_i of i, then slots_j of j."""
    L = xyz.shape[1]
    pts_i = xyz[:, :, list(slots_i)].repeat_interleave(L, dim=1)
    pts_j = xyz[:, :, list(slots_j)].repeat(1, L, 1, 1)
    return torch.cat([pts_i, pts_j], dim=-2)


def pair_dihedrals(xyz: torch.Tensor, slots_i: Sequence[int], slots_j: Sequence[int]) -> torch.Tensor:
    """reference protstruc/protstruc.py:634-640"""
    L = xyz.shape[1]
    p = gather_pair_points(xyz, slots_i, slots_j)
    return dihedral(p[:, :, 0], p[:, :, 1], p[:, :, 2], p[:, :, 3]).reshape(-1, L, L)


def pair_planar_angles(xyz: torch.Tensor, slots_i: Sequence[int], slots_j: Sequence[int]) -> torch.Tensor:
    """reference protstruc/protstruc.py:656-660"""
    L = xyz.shape[1]
    p = gather_pair_points(xyz, slots_i, slots_j)
    return planar_angle(p[:, :, 0], p[:, :, 1], p[:, :, 2]).reshape(-1, L, L)


def trrosetta_angles(xyz: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """reference protstruc/protstruc.py:810-815 — omega, theta, phi as the reference defines them."""
    N, CA, CB = SLOT["N"], SLOT["CA"], SLOT["CB"]
    omega = pair_dihedrals(xyz, [CA, CB], [CA, CB])
    theta = pair_dihedrals(xyz, [N, CA, CB], [CB])
    phi = pair_planar_angles(xyz, [CA, CB], [CB])
    return omega, theta, phi


def trrosetta_angles_virtual_cb(xyz: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Same three angles with slot 4 replaced by `virtual_cb(N, CA, C)` (additive option, Q7)."""
    x = xyz[:, :, :5].clone()
    x[:, :, SLOT["CB"]] = virtual_cb(xyz[:, :, 0], xyz[:, :, 1], xyz[:, :, 2])
    return trrosetta_angles(x)


def inter_residue_geometry(xyz: torch.Tensor, atom_mask: torch.Tensor) -> Dict[str, torch.Tensor]:
    """reference protstruc/protstruc.py:797-817"""
    dist, pair_mask = pair_distances(xyz, atom_mask)
    N, CA, O, CB = SLOT["N"], SLOT["CA"], SLOT["O"], SLOT["CB"]
    out = {
        "d_ca": dist[:, :, :, CA, CA], "d_ca_mask": pair_mask[:, :, :, CA, CA],
        "d_cb": dist[:, :, :, CB, CB], "d_cb_mask": pair_mask[:, :, :, CB, CB],
        "d_no": dist[:, :, :, N, O], "d_no_mask": pair_mask[:, :, :, N, O],
    }
    out["omega"], out["theta"], out["phi"] = trrosetta_angles(xyz)
    return out


# ------------------------------------------------------------------ per-residue features (a10-a11)
def terminal_masks(chain_idx: torch.Tensor, residue_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference protstruc/protstruc.py:442-443, 452-453 — NaN-padded shifted compare."""
    left = torch.nn.functional.pad(chain_idx, (1, 0), mode="constant", value=float("nan"))
    right = torch.nn.functional.pad(chain_idx, (0, 1), mode="constant", value=float("nan"))
    nterm = (left[:, :-1] != left[:, 1:]).bool() * residue_mask
    cterm = (right[:, :-1] != right[:, 1:]).bool() * residue_mask
    return nterm, cterm


def backbone_dihedrals(xyz: torch.Tensor, chain_idx: torch.Tensor,
                       residue_mask: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference protstruc/protstruc.py:512-541"""
    pad = torch.nn.functional.pad
    n, ca, c = xyz[:, :, SLOT["N"]], xyz[:, :, SLOT["CA"]], xyz[:, :, SLOT["C"]]
    nterm, cterm = terminal_masks(chain_idx, residue_mask)
    phi = pad(dihedral(c[:, :-1], n[:, 1:], ca[:, 1:], c[:, 1:]), (1, 0, 0, 0), value=0.0)
    phi[nterm] = 0.0
    psi = pad(dihedral(n[:, :-1], ca[:, :-1], c[:, :-1], n[:, 1:]), (0, 1, 0, 0), value=0.0)
    psi[cterm] = 0.0
    omega = pad(dihedral(ca[:, :-1], c[:, :-1], n[:, 1:], ca[:, 1:]), (0, 1, 0, 0), value=0.0)
    omega[cterm] = 0.0
    angles = torch.stack([phi, psi, omega], dim=-1)
    valid = ~torch.stack([nterm, cterm, cterm], dim=-1)
    valid = valid * residue_mask[:, :, None]
    return angles, valid


def frames(xyz: torch.Tensor, a1: int = 0, a2: int = 1, a3: int = 2) -> torch.Tensor:
    """reference protstruc/protstruc.py:567-571"""
    return frames_from_points(xyz[:, :, a1], xyz[:, :, a2], xyz[:, :, a3])


# ------------------------------------------------------------------ rigid-frame family (row f1)
def local_xyz(xyz: torch.Tensor) -> torch.Tensor:
    """reference protstruc/protstruc.py:353-362 — R^T x minus the residue's global CA."""
    n_atoms = xyz.shape[2]
    orientation = frames(xyz)[:, :, None].expand(-1, -1, n_atoms, -1, -1)
    local = torch.einsum("bnaji,bnaj->bnai", orientation, xyz)
    return local - xyz[:, :, SLOT["CA"]].unsqueeze(-2)


def rotate(xyz: torch.Tensor, rotation: torch.Tensor) -> torch.Tensor:
    """reference protstruc/protstruc.py:688-694"""
    rotation = rotation[None, None, None] if rotation.ndim == 2 else rotation[:, None, None]
    return torch.einsum("bnaij,bnaj->bnai", rotation, xyz)


def ideal_backbone(include_cb: bool = False) -> torch.Tensor:
    """reference protstruc/geometry.py:206-224 with the constants of protstruc/constants/ideal.py."""
    NA, AC, NAC = 1.458, 1.523, 1.937
    ca = torch.zeros(3)
    c = torch.tensor([AC, 0.0, 0.0])
    n = torch.tensor([NA * math.cos(NAC), NA * math.sin(NAC), 0.0])
    if include_cb:
        _b, _c = (ca - n), (c - ca)
        _a = torch.linalg.cross(_b, _c)
        cb = -0.58273431 * _a + 0.56802827 * _b - 0.54067466 * _c + ca
        return torch.stack([n, ca, c, cb])
    return torch.stack([n, ca, c])


def frames_to_backbone(orientations: torch.Tensor, translations: torch.Tensor, include_cb: bool = False,
                       n_slots: int = 15) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference protstruc/protstruc.py:289-314 — rotate + translate the ideal residue, zero-pad to 15 slots."""
    B, L = orientations.shape[:2]
    ideal = ideal_backbone(include_cb).expand(B, L, -1, -1)
    n_atoms = ideal.shape[2]
    rot = orientations[:, :, None].expand(-1, -1, n_atoms, -1, -1)
    atoms = torch.einsum("bnaij,bnaj->bnai", rot, ideal) + translations[:, :, None, :]
    mask = torch.ones_like(atoms[..., 0])
    atoms = torch.cat([atoms, torch.zeros(B, L, n_slots - n_atoms, 3)], dim=-2)
    mask = torch.cat([mask, torch.zeros(B, L, n_slots - n_atoms)], dim=-1)
    return atoms, mask


def kabsch(a: torch.Tensor, b: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """reference protstruc/geometry.py:454-480 — SVD of the covariance of the centred point sets."""
    ca, cb = a.mean(dim=-2, keepdim=True), b.mean(dim=-2, keepdim=True)
    h = torch.einsum("ki,kj->ij", a - ca, b - cb)
    u, _, vt = torch.linalg.svd(h)
    v, ut = vt.transpose(-2, -1), u.transpose(-2, -1)
    d = torch.sign(torch.linalg.det(torch.einsum("ij,jk->ik", v, ut)))
    diag = torch.eye(3).clone()
    diag[2, 2] = d
    rot = torch.einsum("ij,jk,kl->...il", v, diag, ut)
    return rot, cb.squeeze(-2) - torch.einsum("ij,j->i", rot, ca.squeeze(-2))


def align(source_xyz: torch.Tensor, target_xyz: torch.Tensor, atom_mask: torch.Tensor):
    """reference protstruc/protstruc.py:899-918, with a single target broadcast to every source structure."""
    B = source_xyz.shape[0]
    src = source_xyz.reshape(B, -1, 3)
    tgt = target_xyz.reshape(target_xyz.shape[0], -1, 3).expand(B, -1, -1)
    msk = atom_mask.reshape(atom_mask.shape[0], -1).bool().expand(B, -1)
    rots, trans = [], []
    for s, t, m in zip(src, tgt, msk):
        r, tr = kabsch(s[m], t[m])
        rots.append(r)
        trans.append(tr)
    rots, trans = torch.stack(rots), torch.stack(trans)
    moved = rotate(source_xyz, rots) + trans[:, None, None, :]
    return moved, rots, trans


def topk_nearest_residue_mask(xyz: torch.Tensor, residue_mask: torch.Tensor, query_xyz: torch.Tensor, k: int = 128,
                              mask: torch.Tensor = None) -> torch.Tensor:
    """reference protstruc/protstruc.py:844-862 (batch of one)."""
    ca = xyz[0, :, SLOT["CA"]]
    dist = torch.norm(ca[:, None] - query_xyz, dim=-1)
    dist, _ = dist.min(dim=-1)
    valid = residue_mask[0] if mask is None else residue_mask[0] & mask
    dist[~valid] = 1e9
    k = min(k, int(valid.sum()))
    _, idx = dist.topk(k, largest=False)
    return torch.zeros(xyz.shape[1], dtype=torch.bool).scatter(0, idx, True).unsqueeze(0)


# ------------------------------------------------------------------ statistics / diffusion (a13-a15)
def _standardize_one(xyz: torch.Tensor, atom_mask: torch.Tensor):
    """reference protstruc/protstruc.py:720-733 for a batch of one (the only shape it is valid for)."""
    b, n, a = atom_mask.shape
    count = atom_mask.reshape(b, n * a).sum(axis=1, keepdims=True)
    masked = (xyz * atom_mask.unsqueeze(-1)).reshape(b, n * a, 3)
    mu = masked.nan_to_num(0.0).sum(axis=1) / count
    centred = xyz.nan_to_num(0.0) - mu.reshape(b, 1, 1, 3)
    centred = (centred**2 * atom_mask.unsqueeze(-1)).reshape(b, n * a, 3)
    sd = torch.sqrt(centred.sum(axis=1) / count)
    return (xyz - mu) / sd, mu, sd


def standardize_per_structure(xyz: torch.Tensor, atom_mask: torch.Tensor):
    """Reference formula applied to each structure's `b:b+1` slice and concatenated (Q1)."""
    outs, mus, sds = [], [], []
    for b in range(xyz.shape[0]):
        o, m, s = _standardize_one(xyz[b:b + 1], atom_mask[b:b + 1])
        outs.append(o)
        mus.append(m)
        sds.append(s)
    return torch.cat(outs), torch.cat(mus), torch.cat(sds)


def unstandardize(xyz: torch.Tensor, mu: torch.Tensor, sd: torch.Tensor) -> torch.Tensor:
    """reference protstruc/protstruc.py:743 with the per-structure broadcast."""
    return xyz * sd[:, None, None, :] + mu[:, None, None, :]


def center_of_mass(xyz: torch.Tensor) -> torch.Tensor:
    """reference protstruc/protstruc.py:756-757"""
    return xyz[:, :, SLOT["CA"]].nanmean(axis=1)


def diffuse(xyz: torch.Tensor, beta: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    """reference protstruc/protstruc.py:875-878 with `randn_like` replaced by the given tensor."""
    beta = beta.reshape(-1, 1, 1, 1)
    scaled_noise = noise * beta.sqrt()
    return (1 - beta).sqrt() * xyz + scaled_noise


def cosine_variance_schedule(T: int, s: float = 8e-3, beta_max: float = 0.999) -> torch.Tensor:
    """The schedule the reference's tutorial defines (docs/tutorials/diffusing_xyz_coordinates.ipynb,
    cell 2); returns beta[0..T] with beta[0] = 0."""
    t = torch.arange(T + 1)
    f_t = torch.cos((t / T + s) / (1 + s) * math.pi / 2.0).square()
    alpha_bar = f_t / f_t[0]
    return torch.cat([torch.tensor([0.0]), torch.clip(1 - alpha_bar[1:] / alpha_bar[:-1], min=1e-5, max=beta_max)])


# ------------------------------------------------------------------ Philox stream of K5 (own RNG)
_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: Tuple[int, int]) -> np.ndarray:
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw; SC'11).  ctr: (n, 4) uint32 -> (n, 4) uint32."""
    c = ctr.astype(np.uint32).copy()
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    mask32 = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[:, 0].astype(np.uint64)
            p1 = _M1 * c[:, 2].astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & mask32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & mask32).astype(np.uint32)
            c = np.stack([hi1 ^ c[:, 1] ^ k0, lo1, hi0 ^ c[:, 3] ^ k1, lo0], axis=1)
            k0 = np.uint32(k0 + _W0)
            k1 = np.uint32(k1 + _W1)
    return c


def philox_normal(n: int, seed: int, step: int, elem_offset: int = 0) -> np.ndarray:
    """float64 evaluation of the N(0,1) stream the K5 kernels define: element e (global index
    e + elem_offset, ANY offset) takes lane (global & 3) of the counter (global >> 2 as lo/hi words,
    step_lo, step_hi), key = seed; uniforms from the top 23 bits of each word; Box-Muller on (r0, r1)
    and (r2, r3)."""
    shift = elem_offset & 3
    groups = (n + shift + 3) // 4
    g = np.arange(groups, dtype=np.uint64) + np.uint64(elem_offset >> 2)
    ctr = np.stack([
        (g & np.uint64(0xFFFFFFFF)).astype(np.uint32), (g >> np.uint64(32)).astype(np.uint32),
        np.full(groups, step & 0xFFFFFFFF, dtype=np.uint32), np.full(groups, (step >> 32) & 0xFFFFFFFF, dtype=np.uint32),
    ], axis=1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))
    u = ((r >> np.uint32(9)).astype(np.float64) + 0.5) * 2.0**-23  # top 23 bits, centred: u in (0, 1)
    rad_a = np.sqrt(-2.0 * np.log(u[:, 0]))
    rad_b = np.sqrt(-2.0 * np.log(u[:, 2]))
    ang_a = 2.0 * np.pi * u[:, 1]
    ang_b = 2.0 * np.pi * u[:, 3]
    z = np.stack([rad_a * np.sin(ang_a), rad_a * np.cos(ang_a), rad_b * np.sin(ang_b), rad_b * np.cos(ang_b)], axis=1)
    return z.reshape(-1)[shift:shift + n]
