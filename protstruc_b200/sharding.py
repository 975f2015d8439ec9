"""Multi-GPU execution of the hot path: partition the batch, no collective on the data path.

Every output element of every hot-path function depends on ONE structure only (reference
protstruc/protstruc.py:477-483, 612-616, 512-539, 720-733, 875-878), so the batch dimension is
split into contiguous shards, one process per GPU (`torch.distributed`, NCCL over NVLink for the
plumbing).  The full (B, L, L, A, A) distance tensor stays sharded where it was produced — it is
written at HBM speed (~6.5 TB/s per GPU) while NVLink egress is ~0.8 TB/s per GPU, so gathering it
would only slow the job down.  The only optional exchange is an all-gather of COMPACT features
((B, L, L) angles and the d_ca / d_cb / d_no slices).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.distributed as dist

from .structure_batch import StructureBatch


def shard_bounds(batch_size: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `rank`'s structures; the first `batch_size % world_size` ranks
    get one extra structure.  Empty shards (start == stop) are legal when B < world_size."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"invalid rank {rank} for world size {world_size}")
    base, extra = divmod(batch_size, world_size)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start, stop


def shard_sizes(batch_size: int, world_size: int) -> List[int]:
    return [b - a for a, b in (shard_bounds(batch_size, world_size, r) for r in range(world_size))]


def shard_structure_batch(xyz, atom_mask=None, chain_idx=None, chain_ids=None, seq=None,
                          rank: Optional[int] = None, world_size: Optional[int] = None,
                          device=None) -> Optional[StructureBatch]:
    """Builds this rank's `StructureBatch` from the GLOBAL host arrays (each rank uploads only its own
    rows: host -> its GPU directly, no scatter collective).  Returns None for an empty shard.

    The diffusion noise stream is keyed by the GLOBAL element index, so `diffuse_xyz` on the shards
    gives exactly what it gives on the unsharded batch, whatever the number of GPUs (any shard offset: the kernels
    address the stream per element).  The noise key follows the state of torch's generator, like `torch.randn_like`
    in the reference: ranks of one job must seed alike (`protstruc_b200.manual_seed(s)` on every rank)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    B = xyz.shape[0]
    start, stop = shard_bounds(B, world_size, rank)
    if start == stop:
        return None
    pick = lambda t: None if t is None else t[start:stop]  # noqa: E731
    sb = StructureBatch.from_xyz(
        xyz[start:stop], pick(atom_mask), pick(chain_idx),
        None if chain_ids is None else chain_ids[start:stop],
        None if seq is None else seq[start:stop],
        device=device,
    )
    # first GLOBAL element of this shard: the kernels address the Philox stream per element
    # (counter = global index >> 2, lane = global index & 3), so any offset reproduces the global stream
    sb._noise_elem_offset = start * int(xyz.shape[1]) * int(xyz.shape[2]) * 3
    return sb


class _PendingGather:
    """Handle of an asynchronous `gather_compact_features`: `wait()` makes the current CUDA stream wait for the
    collectives and returns the gathered dict."""

    def __init__(self, works, finish):
        self._works, self._finish = works, finish

    def wait(self) -> Dict[str, torch.Tensor]:
        for w in self._works:
            w.wait()
        return self._finish()


def gather_compact_features(local: Dict[str, torch.Tensor], batch_size: int, group=None,
                            out: Optional[Dict[str, torch.Tensor]] = None, async_op: bool = False):
    """Optional exchange step: all-gathers per-structure features (leading dimension = this rank's structures) from
    all ranks in rank order, so every rank ends up with the (B, ...) tensors.

    One `all_gather_into_tensor` per feature, straight from the feature tensor into ONE preallocated
    (world * shard, ...) output per feature — no staging copy, no list of buffers, no `cat` when the batch divides
    evenly.  Dense inputs are sent as they are (the fused kernel writes the compact planes densely,
    `StructureBatch.inter_residue_geometry_compact`); only a strided view is made contiguous first.  With uneven shards
    the short ranks send a zero-padded copy and the padding rows are dropped from the result (one gather-copy).
    With NCCL the bytes move GPU-to-GPU over NVLink / NVSwitch.  `async_op=True` returns a handle at once (the
    collectives run on NCCL's own stream, so the next chunk's kernel overlaps them); call `.wait()` for the result.

    Every rank must call this with the same feature names — also a rank whose shard is EMPTY: pass zero-row
    tensors of the right trailing shape and dtype (an empty `StructureBatch` returns such tensors), so that it takes
    part in every collective."""
    if not dist.is_initialized():
        return _PendingGather([], lambda: dict(local)) if async_op else dict(local)
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(batch_size, world)
    biggest = max(sizes)
    even = all(n == biggest for n in sizes)
    works, gathered, keep = [], {}, []
    for name in sorted(local):
        t = local[name]
        if t.shape[0] != sizes[rank]:
            raise ValueError(f"feature {name!r} has {t.shape[0]} structures, this rank's shard has {sizes[rank]}")
        if not t.is_contiguous():
            t = t.contiguous()
        if t.shape[0] != biggest:  # short shard: zero-padded copy (only on uneven batches)
            pad = torch.zeros((biggest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[: t.shape[0]] = t
            t = pad
        target = None if out is None else out.get(name)
        want = (world * biggest,) + tuple(t.shape[1:])
        if target is None or tuple(target.shape) != want or target.dtype != t.dtype or not target.is_contiguous():
            target = torch.empty(want, dtype=t.dtype, device=t.device)
        works.append(dist.all_gather_into_tensor(target, t, group=group, async_op=True))
        gathered[name] = target
        keep.append(t)

    def finish() -> Dict[str, torch.Tensor]:
        keep.clear()  # the (padded / made-contiguous) send buffers lived until the collectives were waited for
        if even:
            return gathered
        rows = torch.cat([torch.arange(r * biggest, r * biggest + n) for r, n in enumerate(sizes)])
        return {name: g.index_select(0, rows.to(g.device)) for name, g in gathered.items()}

    handle = _PendingGather(works, finish)
    return handle if async_op else handle.wait()


COMPACT_FEATURES = ("omega", "theta", "phi", "d_ca", "d_cb", "d_no")


class FusedFeatureGather:
    """The optional exchange step FUSED into the feature kernel: every rank's fused `inter_residue_geometry` launch
    stores the six compact (B, L, L) features of its shard straight into ALL ranks' gathered buffers over NVLink /
    NVSwitch peer memory (`ps_inter_residue_geometry_push`), tile by tile while the distance tensor is being written —
    no separate collective, no staging buffer, nothing for NCCL to wait for.

    The gathered buffer is symmetric memory (`torch.distributed._symmetric_memory`): allocated once per (shard size, L),
    every rank's copy mapped into every other rank's address space; with NVSwitch multicast support one `multimem.st`
    per value reaches all ranks.  `run(sb)` launches the kernel on the current stream, then a symmetric-memory barrier
    across the ranks (device-side, on the same stream), and returns the (world * shard, L, L) tensors — structures in
    rank order; ranks whose shard is shorter than `shard` leave the tail rows of their slab untouched.  The returned
    tensors are views of the reused buffer: they are valid until the next `run` (which, with `guard_reuse`, first waits
    on the same stream until every rank has finished reading them).

    Needs the linear-sweep kernel: A = 15, L >= 32, bool or fp32 atom mask, at most 8 ranks."""

    def __init__(self, shard: int, L: int, group=None, use_multicast: bool = True, guard_reuse: bool = True):
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.shard, self.L = int(shard), int(L)
        # the gathered buffer is REUSED by every `run`: a fast rank's next launch would store into a peer's buffer while
        # that peer still reads the previous step's result.  With `guard_reuse` every `run` after the first starts with a
        # barrier across the ranks on the current stream (device-side, ~10 us), i.e. after each rank's own readers in
        # stream order.  Switch it off only if the caller orders the steps itself.
        self.guard_reuse, self._runs = bool(guard_reuse), 0
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buffer = symm_mem.empty((6, self.world, self.shard, self.L, self.L), dtype=torch.float32, device=dev)
        self.handle = symm_mem.rendezvous(self.buffer, self.group)
        self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
        mc = int(self.handle.multicast_ptr) if use_multicast else 0
        self.multicast_ptr = mc if mc != 0 else None

    def views(self) -> Dict[str, torch.Tensor]:
        n = self.world * self.shard
        return {name: self.buffer[k].view(n, self.L, self.L) for k, name in enumerate(COMPACT_FEATURES)}

    def run(self, sb: StructureBatch, dist_out: Optional[torch.Tensor] = None,
            dist_mask_out: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        import ctypes

        from . import _cabi

        B, L, A = sb._dims()
        if L != self.L or B > self.shard:
            raise ValueError(f"batch of {B} x {L} residues does not fit the gathered buffer ({self.shard} x {self.L})")
        dev = sb.xyz.device
        mask, code = sb._mask_for_kernel(sb.atom_mask)
        if dist_out is None:
            dist_out = torch.empty(B, L, L, A, A, dtype=torch.float32, device=dev)
        if dist_mask_out is None:
            dist_mask_out = torch.empty(B, L, L, A, A, dtype=mask.dtype, device=dev)
        if self.guard_reuse and self._runs > 0:
            self.handle.barrier()  # nobody still reads the buffer this launch is about to overwrite
        self._runs += 1
        if B > 0:
            peers = (ctypes.c_void_p * self.world)(*self.peer_ptrs)
            with _cabi.on_device(dev):
                rc = sb._lib().ps_inter_residue_geometry_push(
                    sb.xyz.data_ptr(), mask.data_ptr(), code, dist_out.data_ptr(), dist_mask_out.data_ptr(), peers,
                    self.world, self.rank, self.multicast_ptr, self.shard, B, L, A, sb._stream())
            _cabi.check(rc, "ps_inter_residue_geometry_push")
        self.handle.barrier()  # every rank's stores have landed before anybody reads
        out = self.views()
        out.update(dist=dist_out, dist_mask=dist_mask_out)
        return out
