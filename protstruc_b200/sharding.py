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
    gives exactly what it gives on the unsharded batch, whatever the number of GPUs."""
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


def gather_compact_features(local: Dict[str, torch.Tensor], batch_size: int,
                            group=None) -> Dict[str, torch.Tensor]:
    """Optional exchange step: all-gathers per-structure features (leading dimension = local batch)
    from all ranks, in rank order, so every rank ends up with the (B, ...) tensors.  Uses one
    `all_gather` per feature on padded equal-size buffers (NCCL needs equal counts); with NCCL the
    bytes move GPU-to-GPU over NVLink/NVSwitch."""
    if not dist.is_initialized():
        return dict(local)
    world = dist.get_world_size(group)
    sizes = shard_sizes(batch_size, world)
    biggest = max(sizes)
    out: Dict[str, torch.Tensor] = {}
    for name in sorted(local):
        t = local[name].contiguous()
        pad = torch.zeros((biggest,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        pad[: t.shape[0]] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(bufs, pad, group=group)
        out[name] = torch.cat([buf[:n] for buf, n in zip(bufs, sizes)], dim=0)
    return out
