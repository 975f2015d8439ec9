"""`StructureBatch` — the reference's batch API for the geometric-feature hot path, on B200 kernels.

Drop-in for the hot-path surface of the reference class (reference protstruc/protstruc.py:32-956):
same method names, argument meaning, return shapes / dtypes and exceptions.  Every feature method
launches a hand-written sm_100a kernel through the C-ABI in `include/protstruc_b200.h`; PyTorch
only owns the device buffers and the stream.  There is no CPU path: tensors are moved to the CUDA
device at construction, and a compute call without a CUDA tensor or without the built library
raises.

Parity decisions that differ from a literal reading of the reference are listed in DESIGN.md
("quirk ledger", following SURVEY.md Appendix A): fp32 contract (Q3), per-structure broadcast in
`standardize` (Q1), last-axis cross product in frames (Q2), derived state created on the data's
device (Q10), usable mask arguments in `standardize` (Q8).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple, Union

import numpy as np
import torch

from . import _cabi, constants
from .general import ATOM, MAX_N_ATOMS_PER_RESIDUE

ArrayLike = Union[np.ndarray, torch.Tensor]


# ------------------------------------------------------------------------------------------------
# RNG stream of diffuse_xyz: (key, step) of the Philox generator inside the kernels.
class _PhiloxStream:
    """Derives the (key, step) pair of every diffusion step from the state of a torch generator (the global
    CPU generator unless `generator=` is given), the way `torch.randn_like` in the reference consumes it
    (reference protstruc.py:876):

    * the first call on a generator — and the first call after anybody else touched that generator
      (`torch.manual_seed`, `generator.manual_seed`, any other random draw) — starts a SESSION: 63 random bits
      are drawn from the generator and become the Philox key, the step counter starts at 0;
    * while the generator is left alone, later calls continue the session: the step counter advances by the
      number of diffusion steps taken, so T calls of `diffuse_xyz` and one `diffuse_xyz_steps` of T steps
      use the same noise.

    Consequences: switching between generators never rewinds either of them (one session per generator
    object); re-seeding — also with the same seed, also by re-creating a generator — restarts the stream
    reproducibly; no noise is ever reused without a re-seed."""

    _MAX_SESSIONS = 64

    def __init__(self) -> None:
        self._sessions: Dict[int, list] = {}  # id(generator) -> [generator, fingerprint, key, step]

    @staticmethod
    def _fingerprint(generator: torch.Generator) -> bytes:
        # the state itself (5 KB for the CPU generator), compared byte for byte: hashing it would cost as much
        # again as reading it, and this sits on the per-call path of the reference-style diffusion loop
        return generator.get_state().numpy().tobytes()

    def reset(self) -> None:
        self._sessions.clear()

    def reserve(self, n_steps: int, generator: Optional[torch.Generator]) -> Tuple[int, int]:
        g = torch.default_generator if generator is None else generator
        session = self._sessions.get(id(g))
        if session is None or session[1] != self._fingerprint(g):
            draw = torch.randint(0, 2 ** 63 - 1, (1,), generator=g, dtype=torch.int64, device=g.device)
            if len(self._sessions) >= self._MAX_SESSIONS:
                self._sessions.pop(next(iter(self._sessions)))
            session = [g, self._fingerprint(g), int(draw.item()), 0]
            self._sessions[id(g)] = session
        first = session[3]
        session[3] += n_steps
        return session[2], first


_philox = _PhiloxStream()


def manual_seed(seed: int) -> None:
    """Seeds torch's global generator and restarts the diffusion noise stream."""
    torch.manual_seed(seed)
    _philox.reset()


def _always_tensor(x):
    return torch.from_numpy(x) if isinstance(x, np.ndarray) else x


def _as_bytes(mask: torch.Tensor) -> torch.Tensor:
    """0 / 1 bytes of a mask for the kernels that read `const uint8_t*`: a bool tensor is reinterpreted in place (no
    conversion kernel: these calls are latency-critical for single small structures), anything else is converted."""
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    return (mask != 0).contiguous().view(torch.uint8)


def _target_device(xyz: torch.Tensor, device) -> torch.device:
    if device is not None:
        return torch.device(device)
    if xyz.is_cuda:
        return xyz.device
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return xyz.device  # no GPU: construction and validation still work, compute raises


class StructureBatch:
    """A padded batch of protein structures: `xyz` (B, L, A, 3) fp32 plus masks."""

    def __init__(
        self,
        xyz: torch.Tensor,
        atom_mask: Optional[torch.Tensor] = None,
        chain_idx: Optional[torch.Tensor] = None,
        chain_ids: Optional[List[List[str]]] = None,
        seq: Optional[List[Dict[str, str]]] = None,
        residue_idx: Optional[torch.Tensor] = None,
        device=None,
    ):
        if (chain_idx is not None and chain_ids is None) or (chain_idx is None and chain_ids is not None):
            raise ValueError("Both `chain_idx` and `chain_ids` should be provided or None.")
        if xyz.ndim != 4 or xyz.shape[-1] != 3:
            raise ValueError(f"`xyz` must have shape (batch, residues, atoms, 3), got {tuple(xyz.shape)}")

        dev = _target_device(xyz, device)
        # fp32 is the contract of the kernels (fp64 input is converted here)
        self.xyz = xyz.to(device=dev, dtype=torch.float32).contiguous()
        self.atom_mask = atom_mask.to(dev) if atom_mask is not None else None
        self.batch_size, self.n_residues, self.max_n_atoms_per_residue = self.xyz.shape[:3]

        if atom_mask is not None:
            if tuple(atom_mask.shape) != tuple(self.xyz.shape[:3]):
                raise ValueError(
                    f"`atom_mask` shape {tuple(atom_mask.shape)} does not match xyz {tuple(self.xyz.shape[:3])}"
                )
            self.residue_mask = self.atom_mask.any(dim=-1)
        else:
            self.residue_mask = torch.ones(self.batch_size, self.n_residues, dtype=torch.bool, device=dev)

        if chain_idx is not None:
            for i, chidx in enumerate(chain_idx):
                msk = ~torch.isnan(chidx)
                assert chidx[msk].min() == 0, f"Protein {i}: Chain index should start from zero"
            self.chain_idx = chain_idx.to(dev)
        else:
            self.chain_idx = torch.zeros(self.batch_size, self.n_residues, device=dev)

        self.chain_ids = chain_ids
        self.seq = seq
        self.residue_idx = residue_idx
        self._standardized = False
        # first global element of this shard in the diffusion noise stream (see sharding.py)
        self._noise_elem_offset = 0

    # -------------------------------------------------------------------------------- constructors
    @classmethod
    def from_xyz(
        cls,
        xyz: ArrayLike,
        atom_mask: Optional[ArrayLike] = None,
        chain_idx: Optional[ArrayLike] = None,
        chain_ids: Optional[List[List[str]]] = None,
        seq: Optional[List[Dict[str, str]]] = None,
        **kwargs,
    ) -> "StructureBatch":
        """Builds a batch from coordinates (B, L, A, 3); numpy inputs are accepted
        (reference protstruc.py:93-128)."""
        return cls(_always_tensor(xyz), _always_tensor(atom_mask), _always_tensor(chain_idx), chain_ids, seq,
                   **kwargs)

    @classmethod
    def from_pdb(cls, pdb_path, **kwargs) -> "StructureBatch":
        """Builds a batch from one PDB file or a list of them (reference protstruc.py:130-192), parsed by the
        native ingest (no biotite): xyz (B, L, 15, 3) with NaN for missing atoms and zeros for padding, a bool
        atom mask, NaN-padded chain indices, chain ids and per-chain sequences."""
        from . import pdb_ingest

        paths = pdb_path if isinstance(pdb_path, list) else [pdb_path]
        arrays = pdb_ingest.read_pdb_batch(paths)
        return cls(torch.from_numpy(arrays["xyz"]), torch.from_numpy(arrays["atom_mask"]),
                   torch.from_numpy(arrays["chain_idx"]), arrays["chain_ids"], arrays["seq"],
                   torch.from_numpy(arrays["residue_idx"]), **kwargs)

    @classmethod
    def from_backbone_orientations_translations(
        cls,
        orientations: ArrayLike,
        translations: ArrayLike,
        chain_idx: Optional[ArrayLike] = None,
        chain_ids: Optional[List[List[str]]] = None,
        seq: Optional[List[Dict[str, str]]] = None,
        residue_idx: Optional[ArrayLike] = None,
        include_cb: bool = False,
        **kwargs,
    ) -> "StructureBatch":
        """Places the ideal backbone (N, CA, C and optionally CB) of every residue with its frame:
        xyz[b,l,a] = R[b,l] @ ideal[a] + t[b,l]; the remaining slots up to 15 are zero and masked out
        (reference protstruc.py:263-319).  One kernel launch; the mask is fp32 like the reference's."""
        orientations, translations = _always_tensor(orientations), _always_tensor(translations)
        if orientations.ndim != 4 or tuple(orientations.shape[2:]) != (3, 3):
            raise ValueError(f"`orientations` must have shape (batch, residues, 3, 3), got {tuple(orientations.shape)}")
        B, L = orientations.shape[:2]
        if tuple(translations.shape) != (B, L, 3):
            raise ValueError(f"`translations` must have shape ({B}, {L}, 3), got {tuple(translations.shape)}")
        dev = _target_device(orientations, kwargs.pop("device", None))
        if dev.type != "cuda":
            raise _cabi.NativeLibraryError("from_backbone_orientations_translations needs a CUDA device: "
                                           "there is no CPU fallback")
        lib = _cabi.load()
        A = MAX_N_ATOMS_PER_RESIDUE
        rot = orientations.to(device=dev, dtype=torch.float32).contiguous()
        tr = translations.to(device=dev, dtype=torch.float32).contiguous()
        ideal = constants.ideal_backbone(include_cb).to(dev).contiguous()
        xyz = torch.empty(B, L, A, 3, dtype=torch.float32, device=dev)
        mask = torch.empty(B, L, A, dtype=torch.float32, device=dev)
        with _cabi.on_device(dev):
            rc = lib.ps_frames_to_backbone(rot.data_ptr(), tr.data_ptr(), ideal.data_ptr(), ideal.shape[0], B, L, A,
                                           xyz.data_ptr(), mask.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ps_frames_to_backbone")
        return cls(xyz, mask, _always_tensor(chain_idx), chain_ids, seq, _always_tensor(residue_idx), device=dev,
                   **kwargs)

    @classmethod
    def from_dihedrals(cls, dihedrals, chain_idx=None, chain_ids=None, **kwargs):
        """Present for API compatibility: the reference's method is an unimplemented stub that returns None
        (reference protstruc.py:321-339)."""
        return None

    # ------------------------------------------------------------------------------------- getters
    def get_batch_size(self) -> int:
        return self.batch_size

    def get_xyz(self) -> torch.Tensor:
        return self.xyz

    def get_atom_mask(self) -> torch.Tensor:
        return self.atom_mask

    def get_local_xyz(self) -> torch.Tensor:
        """Coordinates of every atom in its residue's backbone frame: R^T x minus the residue's (global)
        CA, exactly the expression of the reference (protstruc.py:347-362).  (B, L, A, 3)."""
        B, L, A = self._dims()
        if self._is_empty():
            return torch.empty_like(self.xyz)
        lib = self._lib()
        if A <= int(ATOM.C):
            raise IndexError(f"index {int(ATOM.C)} is out of bounds for dimension 2 with size {A}")
        out = torch.empty_like(self.xyz)
        with _cabi.on_device(self.xyz.device):
            rc = lib.ps_local_xyz(self.xyz.data_ptr(), B, L, A, int(ATOM.N), int(ATOM.CA), int(ATOM.C), int(ATOM.CA),
                                  out.data_ptr(), self._stream())
        _cabi.check(rc, "ps_local_xyz")
        return out

    def get_residue_mask(self) -> torch.Tensor:
        """CA-slot mask as bool (reference protstruc.py:372-378; not the same as `self.residue_mask`)."""
        return self.atom_mask[:, :, ATOM.CA].bool()

    def get_chain_idx(self) -> torch.Tensor:
        return self.chain_idx.long()

    def get_chain_ids(self):
        return self.chain_ids

    def get_seq(self):
        return self.seq

    def get_seq_idx(self) -> torch.Tensor:
        """Integer-encoded sequence (B, L), UNK = 20 beyond each structure (reference protstruc.py:394-409)."""
        from .pdb_ingest import AA_INDEX

        seq_idx = torch.full((self.batch_size, self.n_residues), AA_INDEX["X"], dtype=torch.long)
        for i, (seqdict, chain_ids) in enumerate(zip(self.seq, self.chain_ids)):
            joined = "".join(seqdict[c] for c in chain_ids)
            seq_idx[i, : len(joined)] = torch.tensor([AA_INDEX[r] for r in joined], dtype=torch.long)
        return seq_idx.to(self.xyz.device)

    def get_total_lengths(self) -> torch.Tensor:
        return self.residue_mask.cumsum(dim=1).argmax(dim=1) + 1

    def get_max_n_residues(self) -> int:
        return self.n_residues

    def get_max_n_atoms_per_residue(self) -> int:
        return self.max_n_atoms_per_residue

    def get_n_terminal_mask(self) -> torch.Tensor:
        """True where the previous residue belongs to another chain (NaN-padded compare), times the
        residue mask (reference protstruc.py:435-443).  O(B*L) index bookkeeping, evaluated with
        torch on the device; the fused kernel recomputes it internally for backbone_dihedrals."""
        nan = torch.full_like(self.chain_idx[:, :1], float("nan"))
        padded = torch.cat([nan, self.chain_idx], dim=1)
        return (padded[:, :-1] != padded[:, 1:]).bool() * self.residue_mask

    def get_c_terminal_mask(self) -> torch.Tensor:
        """True where the next residue belongs to another chain (reference protstruc.py:445-453)."""
        nan = torch.full_like(self.chain_idx[:, :1], float("nan"))
        padded = torch.cat([self.chain_idx, nan], dim=1)
        return (padded[:, :-1] != padded[:, 1:]).bool() * self.residue_mask

    # ------------------------------------------------------------------------------ native plumbing
    def _lib(self):
        if not self.xyz.is_cuda:
            raise _cabi.NativeLibraryError(
                "StructureBatch holds CPU tensors (no CUDA device was available at construction); "
                "protstruc_b200 computes on B200 only and has no CPU fallback"
            )
        return _cabi.load()

    def _is_empty(self) -> bool:
        """No structures or no residues: every feature is an empty tensor and nothing is launched."""
        return self.batch_size == 0 or self.n_residues == 0 or self.max_n_atoms_per_residue == 0

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.xyz.device).cuda_stream

    def _dims(self) -> Tuple[int, int, int]:
        return self.batch_size, self.n_residues, self.max_n_atoms_per_residue

    def _mask_for_kernel(self, mask: torch.Tensor) -> Tuple[torch.Tensor, int]:
        """Returns (contiguous mask tensor the kernel can read, PS_MASK_* code)."""
        if mask.dtype == torch.bool:
            return mask.contiguous(), _cabi.PS_MASK_BOOL
        return mask.to(torch.float32).contiguous(), _cabi.PS_MASK_F32

    # ------------------------------------------------------------------------------------ features
    def pairwise_distance_matrix(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """All-atom pairwise distances and the pair mask (reference protstruc.py:455-484).

        Returns `dist` (B, L, L, A, A) fp32 — NOT masked, NaN coordinates give NaN distances — and
        `dist_mask` of the same shape and of `atom_mask`'s dtype (bool stays bool)."""
        if self.atom_mask is None:
            raise TypeError("'NoneType' object is not subscriptable (pairwise_distance_matrix needs atom_mask)")
        B, L, A = self._dims()
        dev = self.xyz.device
        if self._is_empty():
            return (torch.empty(B, L, L, A, A, dtype=torch.float32, device=dev),
                    torch.empty(B, L, L, A, A, dtype=self.atom_mask.dtype, device=dev))
        lib = self._lib()
        mask, code = self._mask_for_kernel(self.atom_mask)
        dist = torch.empty(B, L, L, A, A, dtype=torch.float32, device=dev)
        dist_mask = torch.empty(B, L, L, A, A, dtype=mask.dtype, device=dev)
        with _cabi.on_device(dev):
            rc = lib.ps_pair_dist_mask(self.xyz.data_ptr(), mask.data_ptr(), code, dist.data_ptr(),
                                       dist_mask.data_ptr(), B, L, A, self._stream())
        _cabi.check(rc, "ps_pair_dist_mask")
        if dist_mask.dtype != self.atom_mask.dtype:
            dist_mask = dist_mask.to(self.atom_mask.dtype)
        return dist, dist_mask

    def _slots(self, atoms_i: List[str], atoms_j: List[str]) -> Tuple[List[int], List[int]]:
        for atom in atoms_i + atoms_j:
            if not ATOM.is_valid(atom):
                raise ValueError(f"Atom {atom} is not valid.")
        return [int(ATOM[a]) for a in atoms_i], [int(ATOM[a]) for a in atoms_j]

    def _pair_angles(self, atoms_i: List[str], atoms_j: List[str], kind: int, need: int) -> torch.Tensor:
        si, sj = self._slots(atoms_i, atoms_j)
        if len(si) + len(sj) != need:
            raise ValueError(f"expected {need} atoms in total, got {len(si)} + {len(sj)}")
        B, L, A = self._dims()
        out = torch.empty(B, L, L, dtype=torch.float32, device=self.xyz.device)
        if self._is_empty():
            return out
        lib = self._lib()
        for s in si + sj:
            if s >= A:
                raise IndexError(f"index {s} is out of bounds for dimension 2 with size {A}")
        with _cabi.on_device(self.xyz.device):
            rc = lib.ps_pair_angles(self.xyz.data_ptr(), B, L, A, _cabi.int_array(si), len(si),
                                    _cabi.int_array(sj), len(sj), kind, out.data_ptr(), self._stream())
        _cabi.check(rc, "ps_pair_angles")
        return out

    def pairwise_dihedrals(self, atoms_i: List[str], atoms_j: List[str]) -> torch.Tensor:
        """Dihedral of (atoms_i of residue i, atoms_j of residue j), 4 atoms in total → (B, L, L)
        (reference protstruc.py:620-640)."""
        return self._pair_angles(atoms_i, atoms_j, _cabi.PS_ANGLE_DIHEDRAL, 4)

    def pairwise_planar_angles(self, atoms_i: List[str], atoms_j: List[str]) -> torch.Tensor:
        """Planar angle of 3 atoms split between residue i and residue j → (B, L, L)
        (reference protstruc.py:642-660)."""
        return self._pair_angles(atoms_i, atoms_j, _cabi.PS_ANGLE_PLANAR, 3)

    def trrosetta_angles(self, virtual_cb: bool = False) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """omega, theta, phi exactly as inter_residue_geometry defines them, in one pass
        (reference protstruc.py:810-815).  `virtual_cb=True` recomputes CB from N, CA, C with the
        ideal-geometry coefficients (reference geometry.py:217-221) instead of reading slot 4."""
        B, L, A = self._dims()
        dev = self.xyz.device
        omega = torch.empty(B, L, L, dtype=torch.float32, device=dev)
        theta = torch.empty_like(omega)
        phi = torch.empty_like(omega)
        if self._is_empty():
            return omega, theta, phi
        lib = self._lib()
        with _cabi.on_device(dev):
            rc = lib.ps_trrosetta_angles(self.xyz.data_ptr(), B, L, A, int(bool(virtual_cb)),
                                         omega.data_ptr(), theta.data_ptr(), phi.data_ptr(), self._stream())
        _cabi.check(rc, "ps_trrosetta_angles")
        return omega, theta, phi

    def inter_residue_geometry(self) -> Dict[str, torch.Tensor]:
        """trRosetta-style inter-residue geometry (reference protstruc.py:790-817): d_ca, d_cb, d_no
        (strided views of the full distance tensor, with masks) and omega / theta / phi, produced by
        ONE fused kernel launch (large batches of the 5- / 10-atom layouts: distance tiles, then the exact-sequence
        angle kernel — same bits, the fused tile kernel is angle-bound with so few atoms per residue)."""
        if self.atom_mask is None:
            raise TypeError("'NoneType' object is not subscriptable (inter_residue_geometry needs atom_mask)")
        B, L, A = self._dims()
        if self._is_empty():
            e = lambda dt: torch.empty(B, L, L, dtype=dt, device=self.xyz.device)  # noqa: E731
            md = self.atom_mask.dtype
            return {"d_ca": e(torch.float32), "d_ca_mask": e(md), "d_cb": e(torch.float32), "d_cb_mask": e(md),
                    "d_no": e(torch.float32), "d_no_mask": e(md), "omega": e(torch.float32),
                    "theta": e(torch.float32), "phi": e(torch.float32)}
        lib = self._lib()
        if A <= int(ATOM.CB):
            raise IndexError(f"index {int(ATOM.CB)} is out of bounds for dimension 2 with size {A}")
        dev = self.xyz.device
        mask, code = self._mask_for_kernel(self.atom_mask)
        dist = torch.empty(B, L, L, A, A, dtype=torch.float32, device=dev)
        dist_mask = torch.empty(B, L, L, A, A, dtype=mask.dtype, device=dev)
        omega, theta, phi = torch.empty(3, B, L, L, dtype=torch.float32, device=dev).unbind(0)  # one allocation
        with _cabi.on_device(dev):
            rc = lib.ps_inter_residue_geometry(self.xyz.data_ptr(), mask.data_ptr(), code, dist.data_ptr(),
                                               dist_mask.data_ptr(), omega.data_ptr(), theta.data_ptr(),
                                               phi.data_ptr(), B, L, A, self._stream())
        _cabi.check(rc, "ps_inter_residue_geometry")
        if dist_mask.dtype != self.atom_mask.dtype:
            dist_mask = dist_mask.to(self.atom_mask.dtype)
        # strided (B, L, L) views of the full tensors, like the reference's dist[:, :, :, a, c] (one as_strided each:
        # this method is latency-critical for single small structures)
        block = A * A
        shape, strides = (B, L, L), (L * L * block, L * block, block)

        def pick(t: torch.Tensor, a: int, c: int) -> torch.Tensor:
            return t.as_strided(shape, strides, a * A + c)

        n, ca, cb, o = int(ATOM.N), int(ATOM.CA), int(ATOM.CB), int(ATOM.O)
        return {"d_ca": pick(dist, ca, ca), "d_ca_mask": pick(dist_mask, ca, ca),
                "d_cb": pick(dist, cb, cb), "d_cb_mask": pick(dist_mask, cb, cb),
                "d_no": pick(dist, n, o), "d_no_mask": pick(dist_mask, n, o),
                "omega": omega, "theta": theta, "phi": phi}

    def inter_residue_geometry_compact(self, out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """`inter_residue_geometry` with the six (B, L, L) float features as DENSE planes of one contiguous
        (6, B, L, L) buffer — omega, theta, phi, d_ca, d_cb, d_no — written by the same single fused launch (extension;
        the reference returns the distance planes as strided views, protstruc.py:801-808).  This is the buffer the
        optional multi-GPU exchange gathers over NVLink (`sharding.gather_compact_features`).  `out` may supply the
        buffer (e.g. a slot of a reused ring).  The dict also carries `dist`, `dist_mask` and `compact` (the buffer)."""
        if self.atom_mask is None:
            raise TypeError("'NoneType' object is not subscriptable (inter_residue_geometry needs atom_mask)")
        B, L, A = self._dims()
        dev = self.xyz.device
        names = ("omega", "theta", "phi", "d_ca", "d_cb", "d_no")
        if out is None:
            out = torch.empty(6, B, L, L, dtype=torch.float32, device=dev)
        if tuple(out.shape) != (6, B, L, L) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
            raise ValueError(f"`out` must be a contiguous float32 tensor of shape (6, {B}, {L}, {L}) on {dev}")
        mask, code = self._mask_for_kernel(self.atom_mask)
        dist = torch.empty(B, L, L, A, A, dtype=torch.float32, device=dev)
        dist_mask = torch.empty(B, L, L, A, A, dtype=mask.dtype, device=dev)
        if not self._is_empty():
            lib = self._lib()
            if A <= int(ATOM.CB):
                raise IndexError(f"index {int(ATOM.CB)} is out of bounds for dimension 2 with size {A}")
            with _cabi.on_device(dev):
                rc = lib.ps_inter_residue_geometry_compact(self.xyz.data_ptr(), mask.data_ptr(), code, dist.data_ptr(),
                                                           dist_mask.data_ptr(), out.data_ptr(), B, L, A, self._stream())
            _cabi.check(rc, "ps_inter_residue_geometry_compact")
        if dist_mask.dtype != self.atom_mask.dtype:
            dist_mask = dist_mask.to(self.atom_mask.dtype)
        feats = {name: out[k] for k, name in enumerate(names)}
        feats.update(dist=dist, dist_mask=dist_mask, compact=out)
        return feats

    def _backbone(self, want_dihedrals: bool, frame_slots: Optional[Tuple[int, int, int]]):
        B, L, A = self._dims()
        dev = self.xyz.device
        if self._is_empty():
            return (torch.empty(B, L, 3, dtype=torch.float32, device=dev) if want_dihedrals else None,
                    torch.empty(B, L, 3, dtype=torch.bool, device=dev) if want_dihedrals else None,
                    torch.empty(B, L, 3, 3, dtype=torch.float32, device=dev) if frame_slots is not None else None)
        lib = self._lib()
        dihedrals = dihedral_mask = frames = None
        rm_ptr = ch_ptr = dh_ptr = dm_ptr = fr_ptr = None
        keep = []
        if want_dihedrals:
            if A < 3:
                raise IndexError(f"index 2 is out of bounds for dimension 2 with size {A}")
            rm = _as_bytes(self.residue_mask)
            ch = self.chain_idx.to(torch.float32).contiguous()
            keep += [rm, ch]
            dihedrals = torch.empty(B, L, 3, dtype=torch.float32, device=dev)
            dihedral_mask = torch.empty(B, L, 3, dtype=torch.bool, device=dev)
            rm_ptr, ch_ptr, dh_ptr, dm_ptr = rm.data_ptr(), ch.data_ptr(), dihedrals.data_ptr(), dihedral_mask.data_ptr()
        a1 = a2 = a3 = 0
        if frame_slots is not None:
            a1, a2, a3 = frame_slots
            frames = torch.empty(B, L, 3, 3, dtype=torch.float32, device=dev)
            fr_ptr = frames.data_ptr()
        with _cabi.on_device(dev):
            rc = lib.ps_backbone(self.xyz.data_ptr(), rm_ptr, ch_ptr, B, L, A, a1, a2, a3, dh_ptr, dm_ptr,
                                 fr_ptr, self._stream())
        _cabi.check(rc, "ps_backbone")
        return dihedrals, dihedral_mask, frames

    def backbone_dihedrals(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """phi, psi, omega per residue (B, L, 3) in radians and the bool validity mask (B, L, 3);
        zero at chain termini (reference protstruc.py:486-541)."""
        dihedrals, dihedral_mask, _ = self._backbone(True, None)
        return dihedrals, dihedral_mask

    def backbone_orientations(self, a1: str = "N", a2: str = "CA", a3: str = "C") -> torch.Tensor:
        """Gram-Schmidt frames (B, L, 3, 3), columns e1, e2, e3 (reference protstruc.py:543-571)."""
        slots = (int(ATOM[a1]), int(ATOM[a2]), int(ATOM[a3]))  # KeyError on unknown names, like the reference
        A = self.max_n_atoms_per_residue
        for s in slots:
            if s >= A:
                raise IndexError(f"index {s} is out of bounds for dimension 2 with size {A}")
        return self._backbone(False, slots)[2]

    def backbone_features(self, a1: str = "N", a2: str = "CA", a3: str = "C"):
        """dihedrals, dihedral_mask and frames from a single kernel launch (extension)."""
        slots = (int(ATOM[a1]), int(ATOM[a2]), int(ATOM[a3]))
        return self._backbone(True, slots)

    def backbone_translations(self, atom: str = "CA") -> torch.Tensor:
        """View `xyz[:, :, ATOM[atom]]` (B, L, 3); aliases xyz (reference protstruc.py:573-587)."""
        return self.xyz[:, :, ATOM[atom]]

    # ------------------------------------------------------------------------------------ mutators
    def translate(self, translation: torch.Tensor, atomwise: bool = False) -> None:
        """In-place translation by (B, L, 3) / (B, 1, 3) or atomwise (B, L, A, 3) tensors
        (reference protstruc.py:662-679).  One broadcast-add kernel; `translation` is read through its
        broadcast strides, nothing is materialised."""
        dev = self.xyz.device
        translation = translation.to(device=dev, dtype=torch.float32)
        if not atomwise:
            if translation.ndim != 3:
                raise ValueError(f"`translation` must have shape (batch, residues, 3), got {tuple(translation.shape)}")
            translation = translation.unsqueeze(-2)
        view = translation.expand(self.xyz.shape)  # raises like torch's `+=` if the shapes do not broadcast
        if self._is_empty():
            return
        lib = self._lib()
        if view.stride(-1) != 1:
            # the kernel reads the three coordinates of a translation vector at unit stride; a broadcast or strided
            # coordinate axis (e.g. a (B, L, 1) translation) is materialised with its last axis expanded
            view = translation.expand(translation.shape[:-1] + (3,)).contiguous().expand(self.xyz.shape)
        sb, sl, sa, _ = view.stride()
        B, L, A = self._dims()
        with _cabi.on_device(dev):
            rc = lib.ps_translate_bcast(self.xyz.data_ptr(), view.data_ptr(), sb, sl, sa, B, L, A, self.xyz.data_ptr(),
                                        self._stream())
        _cabi.check(rc, "ps_translate_bcast")

    def rotate(self, rotation: torch.Tensor) -> None:
        """Rotates every structure: (B, 3, 3) per structure or (3, 3) for all; rebinds `self.xyz`
        (reference protstruc.py:681-694)."""
        if rotation.ndim not in (2, 3) or tuple(rotation.shape[-2:]) != (3, 3):
            raise ValueError(f"`rotation` must have shape (batch, 3, 3) or (3, 3), got {tuple(rotation.shape)}")
        if rotation.ndim == 3 and rotation.shape[0] not in (1, self.batch_size):
            raise ValueError(f"`rotation` has {rotation.shape[0]} matrices for {self.batch_size} structures")
        if self._is_empty():  # nothing to rotate: a no-op like the reference's einsum over no atoms
            return
        lib = self._lib()
        dev = self.xyz.device
        rot = rotation.to(device=dev, dtype=torch.float32).reshape(-1, 3, 3).contiguous()
        B, L, A = self._dims()
        out = torch.empty_like(self.xyz)
        with _cabi.on_device(dev):
            rc = lib.ps_rotate(self.xyz.data_ptr(), rot.data_ptr(), rot.shape[0], B, L, A, out.data_ptr(), self._stream())
        _cabi.check(rc, "ps_rotate")
        self.xyz = out

    def standardize(self, atom_mask: Optional[torch.Tensor] = None,
                    residue_mask: Optional[torch.Tensor] = None) -> None:
        """Per-structure, per-axis masked standardisation; sets `mu`, `std` (B, 3)
        (reference protstruc.py:696-734)."""
        if atom_mask is not None and residue_mask is not None:
            raise ValueError("Only one of atom_mask and residue_mask can be specified.")
        if self._standardized:
            raise ValueError("Coordinates are already standardized.")
        if self.atom_mask is None:
            raise TypeError("standardize needs an atom_mask")
        dev = self.xyz.device
        if self._is_empty():  # mean / deviation over no atoms: 0 / 0 like the reference
            self.mu = torch.full((self.batch_size, 3), float("nan"), dtype=torch.float32, device=dev)
            self.std = self.mu.clone()
            self._standardized = True
            return
        lib = self._lib()
        if atom_mask is not None:
            use = atom_mask.to(dev) * self.atom_mask
        elif residue_mask is not None:
            use = residue_mask.to(dev).unsqueeze(-1) * self.atom_mask
        else:
            use = self.atom_mask
        mask, code = self._mask_for_kernel(use)
        B, L, A = self._dims()
        mu = torch.empty(B, 3, dtype=torch.float32, device=dev)
        sd = torch.empty(B, 3, dtype=torch.float32, device=dev)
        out = torch.empty_like(self.xyz)
        with _cabi.on_device(dev):
            rc = lib.ps_masked_stats(self.xyz.data_ptr(), mask.data_ptr(), code, B, L, A, mu.data_ptr(),
                                     sd.data_ptr(), out.data_ptr(), self._stream())
        _cabi.check(rc, "ps_masked_stats")
        self.mu, self.std = mu, sd
        self.xyz = out
        self._standardized = True

    def unstandardize(self) -> None:
        """Inverse of `standardize` (reference protstruc.py:736-744)."""
        if not self._standardized:
            raise ValueError("Cannot unstandardize structures that are not standardized.")
        if self._is_empty():
            self._standardized = False
            return
        lib = self._lib()
        B, L, A = self._dims()
        out = torch.empty_like(self.xyz)
        with _cabi.on_device(self.xyz.device):
            rc = lib.ps_scale_shift(self.xyz.data_ptr(), self.std.data_ptr(), self.mu.data_ptr(), B, L, A,
                                    out.data_ptr(), self._stream())
        _cabi.check(rc, "ps_scale_shift")
        self.xyz = out
        self._standardized = False

    def center_of_mass(self) -> torch.Tensor:
        """NaN-skipping mean of the CA coordinates over ALL residues, (B, 3); not mask-aware, like
        the reference (protstruc.py:746-757)."""
        B, L, A = self._dims()
        if self._is_empty():  # nanmean over nothing
            return torch.full((B, 3), float("nan"), dtype=torch.float32, device=self.xyz.device)
        lib = self._lib()
        if A <= int(ATOM.CA):
            raise IndexError(f"index {int(ATOM.CA)} is out of bounds for dimension 2 with size {A}")
        out = torch.empty(B, 3, dtype=torch.float32, device=self.xyz.device)
        with _cabi.on_device(self.xyz.device):
            rc = lib.ps_center_of_mass(self.xyz.data_ptr(), B, L, A, int(ATOM.CA), out.data_ptr(), self._stream())
        _cabi.check(rc, "ps_center_of_mass")
        return out

    def center_at(self, center: Optional[torch.Tensor] = None) -> None:
        """Translates every structure so that its CA centre sits at `center` ((B, 3), (3,) or None
        for the origin); in place (reference protstruc.py:759-788)."""
        if center is None:
            # the reference builds zeros(1, 3) here and then rejects it for B > 1 (its own shape
            # check, protstruc.py:769-779); the documented meaning is "the origin" for any B
            center = torch.zeros(3)
        if center.ndim > 2 or center.shape[-1] != 3:
            raise ValueError(f"`center` must have a shape of (batch_size, 3) or (3,), got {center.shape}.")
        if center.ndim == 2 and center.shape[0] != self.batch_size:
            raise ValueError(f"`center` must have a shape of (batch_size, 3) or (3,), got {center.shape}.")
        if center.ndim == 1:
            center = center.unsqueeze(0)
        if self._is_empty():
            return
        lib = self._lib()
        dev = self.xyz.device
        B, L, A = self._dims()
        translation = (center.to(device=dev, dtype=torch.float32) - self.center_of_mass()).contiguous()
        with _cabi.on_device(dev):
            rc = lib.ps_translate(self.xyz.data_ptr(), translation.data_ptr(), translation.shape[0], B, L, A,
                                  self.xyz.data_ptr(), self._stream())
        _cabi.check(rc, "ps_translate")

    def align(self, target: "StructureBatch", atom_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Superimposes every structure onto `target` (same batch size, or a single structure for all) with
        the optimal rigid motion over the atoms selected by `atom_mask` (default: atoms valid in both);
        rotates / translates ALL atoms in place like the reference (protstruc.py:880-918).  The per-structure
        Kabsch solves run as one kernel launch.  Returns the rotations (B, 3, 3)."""
        if target.get_batch_size() != 1 and self.batch_size != target.get_batch_size():
            raise ValueError("Batch size of the two structures must be the same.")
        if tuple(target.get_xyz().shape[1:]) != tuple(self.xyz.shape[1:]):
            raise ValueError("Source and target must have the same (residues, atoms) layout.")
        dev = self.xyz.device
        if self._is_empty():  # no atoms to superimpose: identity motions, nothing moves
            return torch.eye(3, dtype=torch.float32, device=dev).repeat(self.batch_size, 1, 1)
        lib = self._lib()
        if atom_mask is None:
            if self.atom_mask is None or target.get_atom_mask() is None:
                raise TypeError("align needs atom masks (or an explicit `atom_mask`)")
            atom_mask = self.atom_mask * target.get_atom_mask().to(dev)
        B, L, A = self._dims()
        mask = _as_bytes(atom_mask.to(dev).bool().expand(B, L, A).reshape(B, L * A))
        tgt = target.get_xyz().to(device=dev, dtype=torch.float32).contiguous()
        rot = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
        tr = torch.empty(B, 3, dtype=torch.float32, device=dev)
        with _cabi.on_device(dev):
            rc = lib.ps_kabsch(self.xyz.data_ptr(), tgt.data_ptr(), mask.data_ptr(), tgt.shape[0], B, L * A,
                               rot.data_ptr(), tr.data_ptr(), self._stream())
        _cabi.check(rc, "ps_kabsch")
        self.rotate(rot)
        self.translate(tr.unsqueeze(1))
        return rot

    def get_topk_nearest_residue_mask(self, query_xyz: torch.Tensor, k: int = 128,
                                      mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Boolean (1, L) mask of the k residues whose CA is closest to any of the query points
        (reference protstruc.py:819-862; single-structure batches only, like the reference)."""
        if self.batch_size > 1:
            raise ValueError("get_topk_nearest_residue_mask method is not defined "
                             "for a StructureBatch with batch size > 1.")
        dev = self.xyz.device
        B, L, A = self._dims()
        if self._is_empty():
            return torch.zeros(1, L, dtype=torch.bool, device=dev)
        lib = self._lib()
        valid = self.residue_mask[0]
        if mask is not None:
            valid = valid & mask.to(dev)
        k_eff = min(int(k), int(valid.sum().item()))
        query = query_xyz.to(device=dev, dtype=torch.float32).reshape(-1, 3).contiguous()
        valid_u8 = _as_bytes(valid)
        scratch = torch.empty(L, dtype=torch.float32, device=dev)
        out = torch.empty(L, dtype=torch.bool, device=dev)
        with _cabi.on_device(dev):
            rc = lib.ps_topk_nearest_residue_mask(self.xyz.data_ptr(), valid_u8.data_ptr(), query.data_ptr(),
                                                  query.shape[0], L, A, int(ATOM.CA), k_eff, scratch.data_ptr(),
                                                  out.data_ptr(), self._stream())
        _cabi.check(rc, "ps_topk_nearest_residue_mask")
        return out.unsqueeze(0)

    def residue_masked_select(self, mask: torch.Tensor) -> "StructureBatch":
        """New single-structure batch holding only the residues selected by `mask` (1, L)
        (reference protstruc.py:920-956; `chain_ids` / `seq` are carried over unchanged)."""
        if self.batch_size > 1:
            raise ValueError("residue_masked_select method is not defined for a StructureBatch with batch size > 1.")
        if mask.shape != self.residue_mask.shape:
            raise ValueError(f"Mask shape {mask.shape} does not match residue mask shape {self.residue_mask.shape}.")
        if mask.dtype != torch.bool:
            raise ValueError("Mask must be a boolean tensor.")
        mask = mask.to(self.xyz.device)
        xyz = self.xyz[mask].unsqueeze(0)
        atom_mask = self.atom_mask[mask].unsqueeze(0)
        chain_idx = self.chain_idx[mask].unsqueeze(0)
        if self.chain_ids is None:
            return StructureBatch(xyz, atom_mask, device=self.xyz.device)
        return StructureBatch(xyz, atom_mask, chain_idx.cpu(), self.chain_ids, self.seq, device=self.xyz.device)

    def diffuse_xyz(self, beta: torch.Tensor, noise: Optional[torch.Tensor] = None,
                    generator: Optional[torch.Generator] = None) -> None:
        """One forward-diffusion step xyz <- sqrt(1-beta) xyz + sqrt(beta) z, beta (B,)
        (reference protstruc.py:864-878).  Rebinds `self.xyz` to a new tensor.

        `noise` (extension): inject z explicitly — the result is then bit-identical to the
        reference evaluated with the same z.  Otherwise z comes from the kernel's own Philox
        stream, seeded by `generator` (or torch's global seed)."""
        dev = self.xyz.device
        B, L, A = self._dims()
        beta = beta.to(device=dev, dtype=torch.float32).contiguous()
        if beta.shape != (B,):
            raise ValueError(f"`beta` must have shape ({B},), got {tuple(beta.shape)}")
        if self._is_empty():
            self.xyz = self.xyz.clone()  # rebinds like the reference, nothing to diffuse
            return
        lib = self._lib()
        out = torch.empty_like(self.xyz)
        per_b = L * A * 3
        if noise is not None:
            if noise.shape != self.xyz.shape:
                raise ValueError(f"`noise` must have shape {tuple(self.xyz.shape)}, got {tuple(noise.shape)}")
            z = noise.to(device=dev, dtype=torch.float32).contiguous()
            with _cabi.on_device(dev):
                rc = lib.ps_diffuse(self.xyz.data_ptr(), beta.data_ptr(), z.data_ptr(), 0, 0, 0, out.data_ptr(),
                                    B, per_b, self._stream())
        else:
            seed, step = _philox.reserve(1, generator)
            with _cabi.on_device(dev):
                rc = lib.ps_diffuse(self.xyz.data_ptr(), beta.data_ptr(), None, seed, step,
                                    self._noise_elem_offset, out.data_ptr(), B, per_b, self._stream())
        _cabi.check(rc, "ps_diffuse")
        self.xyz = out

    def diffuse_xyz_steps(self, betas: torch.Tensor, generator: Optional[torch.Generator] = None) -> None:
        """T diffusion steps fused in one kernel (extension); `betas` is (T, B).  Bit-identical to
        calling `diffuse_xyz(betas[t])` for t = 0..T-1 on the same noise stream."""
        dev = self.xyz.device
        B, L, A = self._dims()
        betas = betas.to(device=dev, dtype=torch.float32).contiguous()
        if betas.ndim != 2 or betas.shape[1] != B:
            raise ValueError(f"`betas` must have shape (T, {B}), got {tuple(betas.shape)}")
        T = betas.shape[0]
        if T == 0 or self._is_empty():
            return
        lib = self._lib()
        out = torch.empty_like(self.xyz)
        seed, step0 = _philox.reserve(T, generator)
        with _cabi.on_device(dev):
            rc = lib.ps_diffuse_steps(self.xyz.data_ptr(), betas.data_ptr(), T, seed, step0,
                                      self._noise_elem_offset, out.data_ptr(), B, L * A * 3, self._stream())
        _cabi.check(rc, "ps_diffuse_steps")
        self.xyz = out

    def diffusion_loop(self, betas: torch.Tensor, return_trajectory: bool = False,
                       generator: Optional[torch.Generator] = None) -> Optional[torch.Tensor]:
        """The reference's diffusion loop — `for t in range(T): sb.diffuse_xyz(betas[t])` (README.md:131-146,
        docs/tutorials/diffusing_xyz_coordinates.ipynb) — without a Python round trip per step (extension).
        `betas` is (T, B) or (T,) (one schedule for all structures).  Rebinds `self.xyz` to the final state.

        `return_trajectory=False`: one fused kernel, the state stays in registers for all T steps.
        `return_trajectory=True`: returns the (T, B, L, A, 3) tensor of every intermediate state (slice t = after step
        t, what the tutorial collects for its animation); the T one-step launches are issued back to back by the
        native library.  Either way the result is bit-identical to the T single calls on the same noise stream."""
        dev = self.xyz.device
        B, L, A = self._dims()
        betas = betas.to(device=dev, dtype=torch.float32)
        if betas.ndim == 1:
            betas = betas[:, None].expand(betas.shape[0], B)
        betas = betas.contiguous()
        if betas.ndim != 2 or betas.shape[1] != B:
            raise ValueError(f"`betas` must have shape (T,) or (T, {B}), got {tuple(betas.shape)}")
        T = betas.shape[0]
        if not return_trajectory:
            self.diffuse_xyz_steps(betas, generator=generator)
            return None
        trajectory = torch.empty((T, B, L, A, 3), dtype=torch.float32, device=dev)
        if T == 0 or self._is_empty():
            return trajectory
        lib = self._lib()
        seed, step0 = _philox.reserve(T, generator)
        with _cabi.on_device(dev):
            rc = lib.ps_diffuse_trajectory(self.xyz.data_ptr(), betas.data_ptr(), T, seed, step0, self._noise_elem_offset,
                                           trajectory.data_ptr(), B, L * A * 3, self._stream())
        _cabi.check(rc, "ps_diffuse_trajectory")
        # a copy, not a view: in-place mutators (translate, center_at) must not edit the returned trajectory, and
        # dropping the trajectory must free it
        self.xyz = trajectory[T - 1].clone()
        return trajectory
