#!/usr/bin/env python
"""Raw device->host ceiling of this box, per rank and aggregate: what bounds the end-to-end (`e2e`) arm of bench.py.

    python tools/d2h_ceiling.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/d2h_ceiling.py [--seconds 2] [--out gpurun_out/d2h_ceiling_N.json]

Every rank streams a 1.19 GB device buffer (the result bytes of four 512-residue structures) into host memory for
`--seconds`, all ranks at the same time (barrier before, max-over-ranks time after), with NO kernel in the loop.
Variants (each its own timed region):

  memcpy_pinned      one cudaMemcpyAsync per copy into cudaHostAlloc'ed memory (torch pin_memory)
  memcpy_pinned_x2   the same on two streams / two halves (two copy engines)
  memcpy_hugepage    destination = 2 MB-hugepage-backed anonymous mapping (madvise MADV_HUGEPAGE), cudaHostRegister'ed
  kernel_stores      a kernel writes the same bytes straight into the mapped pinned buffer (128-bit stores over PCIe)
  h2d_pinned         the opposite direction, for reference

Rank 0 prints one JSON object (and writes it to --out): per-variant aggregate GB/s, per-rank GB/s, min / max over ranks.
The bench's e2e arm moves 298 MB device->host per structure, so  aggregate GB/s / 0.298  is the structures/s ceiling.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import mmap
import os
import sys
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

BYTES = 4 * 298_157_568 // 16 * 16  # ~1.19 GB


def hugepage_buffer(nbytes: int):
    """Anonymous mapping with transparent huge pages requested, page-locked with cudaHostRegister.  Returns
    (uint8 tensor view, keepalive) or (None, reason)."""
    try:
        size = (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        mm = mmap.mmap(-1, size + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        buf = (ctypes.c_char * len(mm)).from_buffer(mm)
        addr = ctypes.addressof(buf)
        aligned = (addr + (2 << 20) - 1) // (2 << 20) * (2 << 20)
        libc = ctypes.CDLL("libc.so.6", use_errno=True)
        libc.madvise.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        rc = libc.madvise(ctypes.c_void_p(aligned), size, 14)  # MADV_HUGEPAGE
        ctypes.memset(aligned, 0, size)  # first touch by this rank's thread
        arr = (ctypes.c_uint8 * size).from_address(aligned)
        t = torch.frombuffer(arr, dtype=torch.uint8)
        err = torch.cuda.cudart().cudaHostRegister(aligned, size, 0)
        if int(err) != 0:
            return None, f"cudaHostRegister failed ({err})"
        return t[:nbytes], (mm, buf, arr, rc)
    except Exception as exc:  # noqa: BLE001 - diagnostic tool: report and go on
        return None, f"{type(exc).__name__}: {exc}"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=2.0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    import torch.distributed as dist

    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from protstruc_b200 import _cabi

    lib = _cabi.load()
    src = torch.empty(BYTES, dtype=torch.uint8, device=dev)
    src.random_(0, 255)
    pinned = torch.empty(BYTES, dtype=torch.uint8, pin_memory=True)
    pinned.zero_()
    huge, huge_keep = hugepage_buffer(BYTES)
    s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    half = BYTES // 2 // 16 * 16

    def run_variant(issue):
        issue()  # warm-up
        torch.cuda.synchronize()
        barrier()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < args.seconds:
            issue()
            torch.cuda.synchronize()
            n += 1
        dt = time.perf_counter() - t0
        return n * BYTES / dt / 1e9

    def memcpy_pinned():
        pinned.copy_(src, non_blocking=True)

    def memcpy_x2():
        with torch.cuda.stream(s0):
            pinned[:half].copy_(src[:half], non_blocking=True)
        with torch.cuda.stream(s1):
            pinned[half:].copy_(src[half:], non_blocking=True)

    def memcpy_huge():
        huge.copy_(src, non_blocking=True)

    def kernel_stores():
        _cabi.check(lib.ps_debug_fill_pattern(pinned.data_ptr(), BYTES // 4, 2, torch.cuda.current_stream().cuda_stream),
                    "ps_debug_fill_pattern")

    def h2d():
        src.copy_(pinned, non_blocking=True)

    variants = [("memcpy_pinned", memcpy_pinned), ("memcpy_pinned_x2", memcpy_x2)]
    if huge is not None:
        variants.append(("memcpy_hugepage", memcpy_huge))
    variants += [("kernel_stores", kernel_stores), ("h2d_pinned", h2d)]
    results = {}
    for name, fn in variants:
        try:
            gbs = run_variant(fn)
        except Exception as exc:  # noqa: BLE001
            gbs = float("nan")
            if rank == 0:
                print(f"{name}: {type(exc).__name__}: {exc}", file=sys.stderr)
        t = torch.tensor([gbs], dtype=torch.float64, device=dev)
        if world > 1:
            allv = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allv, t)
            per_rank = [float(v.item()) for v in allv]
        else:
            per_rank = [gbs]
        results[name] = {"aggregate_gbs": sum(per_rank), "per_rank_gbs": per_rank, "min": min(per_rank), "max": max(per_rank)}
    if rank == 0:
        out = {"n_gpus": world, "bytes_per_copy": BYTES, "seconds": args.seconds, "host_cpus": os.cpu_count(),
               "hugepage_buffer": "ok" if huge is not None else str(huge_keep), "variants": results,
               "structures_per_s_ceiling": {k: v["aggregate_gbs"] / 0.298157568 for k, v in results.items()}}
        text = json.dumps(out, indent=1)
        print(text)
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(text)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
