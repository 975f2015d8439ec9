#!/usr/bin/env python
"""Benchmark of the geometric-feature hot path (BASELINE.json: structures/s on 512-residue x 15-atom
pairwise features, and achieved HBM GB/s against the measured peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload m|c5] [--with-gather]

Workload `m` (default, the metric shape): a "step" is one pass of the full pairwise feature set (distance matrix +
pair mask + omega/theta/phi, i.e. StructureBatch.inter_residue_geometry) over one batch of synthetic structures of
L = 512, A = 15.  One fused kernel launch per step.  N > 1: launched by torchrun, one rank per GPU, the batch
dimension is partitioned (WEAK scaling: every GPU gets the same per-step batch), no collective on the data path.

Workload `c5` (BASELINE config 5): 4096 structures of L = 384, A = 15, STRONG scaling — the 4096 structures are
partitioned over the ranks and each rank streams its shard in chunks of 128 through one reused output buffer
(687 GB of results in total); a step is one pass over all 4096 structures.

`--with-gather` adds the optional exchange step of SURVEY 8(e): the six compact (B, L, L) features are written
densely by the fused kernel and all-gathered over NVLink (NCCL), alone and overlapped with the next step's kernel.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM; `e2e` = the same
metric through the host-buffer API (pinned host inputs, every result byte copied back to the host inside the timed
region) with the box's raw device->host ceiling measured beside it; `roofline` = algorithmic bytes / CUDA-event time
of the fused kernel against the measured HBM peak; `cpu_baseline` = the CPU oracle port (same ATen / numpy op
sequence as the reference) timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's CPU algorithm (the oracle port; the reference is pure Python, there is
nothing to compile into oracle/_ref) on the same config / metric / unit.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

N_ATOM = 15
UNIT = "structures/s"
FALLBACK_HBM_GBS = 6650.0
# the fused A = 15 kernel a launch can take (ps_pair_dist_last_plan says which): name as ncu prints it
KERNEL_NAMES = {1: "pair_sweep_kernel<0, 0, 1>",          # linear sweep: dist + bool mask, ftz sqrt, fused angles
                0: "pair_tiles_kernel<15, 0, 0, 1, 2>"}   # column strips: A = 15, dist + bool mask, ftz sqrt, angles, 2 warps / tile

WORKLOADS = {
    "m": {"L": 512, "metric": "structures/sec, 512-res x 15-atom pairwise features (dist + mask + omega/theta/phi)",
          "scaling": "weak"},
    "c5": {"L": 384, "metric": "structures/sec, batch-sharded 4096 x 384-res x 15-atom full feature extraction "
                               "(dist + mask + omega/theta/phi; BASELINE config 5)", "scaling": "strong", "total": 4096},
}


def bytes_per_structure(L: int) -> int:
    """L^2 A^2 (4 + 1) + 3 L^2 4 written, L A 13 read (298,157,568 B at L = 512)."""
    return L * L * (N_ATOM * N_ATOM * 5 + 12) + L * N_ATOM * 13


def synthetic_structures(B: int, L: int, seed: int, device):
    """xyz ~ 10 N(0,1) A, Bernoulli(0.5) bool mask, masked slots NaN."""
    g = torch.Generator(device=device).manual_seed(seed)
    xyz = 10.0 * torch.randn(B, L, N_ATOM, 3, device=device, generator=g)
    mask = torch.rand(B, L, N_ATOM, device=device, generator=g) < 0.5
    xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
    return xyz.contiguous(), mask.contiguous()


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons every 20 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def hbm_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def ncu_traffic_per_structure(launched_kernel: str, allow_stale: bool):
    """DRAM bytes (read + write) per structure of the headline kernel, parsed at run time from the NEWEST committed
    `ncu --set full` summary of it under profiles/ (`*k1*_ncu_summary.txt`).  Fails loudly when that capture is of
    another kernel than the one this bench launched — a stale number cannot go unnoticed (`--allow-stale-profile`
    turns the failure into `traffic: null` for runs made while a new kernel is being developed)."""
    def order(path: Path):
        m = re.match(r"r(\d+)([a-z]*)_", path.name)
        return (int(m.group(1)), m.group(2)) if m else (0, "")

    files = sorted(REPO.glob("profiles/*k1*_ncu_summary.txt"), key=order)
    if not files:
        return None, "no profiles/*k1*_ncu_summary.txt committed"
    f = files[-1]
    text = f.read_text()
    kernel = re.search(r"^kernel: (.*)$", text, re.M)
    if kernel is None or launched_kernel not in kernel.group(1):
        msg = (f"{f}: the newest K1 ncu summary is of `{kernel.group(1) if kernel else '?'}`, but this run launched "
               f"`{launched_kernel}` — re-capture the profile (tools/ncu_summary.py)")
        if allow_stale:
            return None, "STALE PROFILE: " + msg
        raise SystemExit(msg)
    per_launch = re.search(r"structures per launch: (\d+)", text)
    structures = int(per_launch.group(1)) if per_launch else 16  # the round-1 captures were taken at 16 per launch
    total = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(rf"{re.escape(key)}\s+([0-9.]+)\s+(\w+)", text)
        if m is None:
            raise SystemExit(f"{f}: no {key} in the summary")
        total += float(m.group(1)) * _SCALE[m.group(2)]
    age_h = (time.time() - f.stat().st_mtime) / 3600.0
    return total / structures, f"{f.relative_to(REPO)} (ncu --set full, {structures} structures per launch; file age {age_h:.1f} h)"


def cpu_reference_throughput(n_structs: int, threads: int, L: int):
    """The CPU oracle port of inter_residue_geometry on `n_structs` structures of the bench shape,
    one structure at a time (the reference needs ~17 B of temporaries per distance element)."""
    from oracle import feature_oracle as orc

    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1)
    xyz = 10.0 * torch.randn(n_structs, L, N_ATOM, 3, generator=g)
    mask = torch.rand(n_structs, L, N_ATOM, generator=g) < 0.5
    xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
    t0 = time.perf_counter()
    for b in range(n_structs):
        out = orc.inter_residue_geometry(xyz[b:b + 1], mask[b:b + 1])
        del out
    dt = time.perf_counter() - t0
    return n_structs / dt, dt


def workload_label(args, B: int) -> str:
    L = WORKLOADS[args.workload]["L"]
    if args.workload == "c5":
        return (f"inter_residue_geometry (full pairwise feature set), BASELINE config 5: 4096 structures of L={L}, "
                f"A={N_ATOM}, batch-sharded, chunks of {args.c5_chunk} through a reused buffer")
    return (f"inter_residue_geometry (full pairwise feature set), L={L}, A={N_ATOM}, "
            f"{B} structures per GPU per step")


def bench_config(args, B: int, L: int, world: int) -> dict:
    """The `config` object: identical in both arms (the reference arm times a bounded sample of the same workload)."""
    out_gb = B * bytes_per_structure(L) / 1e9
    return {"workload": workload_label(args, B), "bytes_per_structure": bytes_per_structure(L),
            "l2_policy": f"outputs of one launch ({out_gb:.2f} GB per GPU) exceed the 126 MB L2; no explicit flush",
            "parallelism": f"batch-sharded x{world}, no collective on the data path"}


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    L = WORKLOADS[args.workload]["L"]
    # a step of the GPU arm covers `--batch` structures; the CPU arm is a per-structure loop whose throughput does
    # not depend on the batch, so each of its steps is a bounded SAMPLE of that batch (same shape, same generator),
    # sized so that `--steps 20 --warmup 5` ends within a few minutes
    per_step = min(args.batch, 16)
    for _ in range(args.warmup):
        cpu_reference_throughput(per_step, threads, L)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_throughput(per_step, threads, L)
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = (f"{per_step} structure(s) of L={L}, A={N_ATOM} per step (bounded sample of the {args.batch}-structure step), "
              f"{args.steps} steps, torch CPU {threads} threads")
    line = {
        "impl": "reference", "metric": WORKLOADS[args.workload]["metric"], "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": WORKLOADS[args.workload]["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": bench_config(args, args.batch, L, args.gpus),
        "note": "CPU oracle port: same ATen/numpy op sequence as the pure-Python reference (bit-identical to it, "
                "tests/golden/MANIFEST.json); a per-structure loop, so its throughput does not depend on the batch: " + sample,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="m")
    ap.add_argument("--batch", type=int, default=32,
                    help="structures per GPU per step (workload m); 32 = 9.5 GB of output per launch: the sweep by batch\n"
                         "(profiles/r4e_bench_by_batch.txt) peaks there — 16: 22.8 k, 32: 22.9 k, 64: 22.0-22.5 k, 96: 21.5 k")
    ap.add_argument("--c5-chunk", type=int, default=64, help="structures per launch (workload c5; 10.7 GB of output)")
    ap.add_argument("--e2e-batch", type=int, default=4, help="structures per GPU per end-to-end call")
    ap.add_argument("--e2e-structures", type=int, default=200, help="structures per GPU timed end to end (>= 1 s at N = 1)")
    ap.add_argument("--e2e-seconds", type=float, default=2.5, help="N > 1: every rank keeps making calls at least this long")
    ap.add_argument("--cpu-sample", type=int, default=96, help="structures timed for the CPU baseline (~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--with-gather", action="store_true", help="also time the optional NVLink gather of compact features")
    ap.add_argument("--allow-stale-profile", action="store_true",
                    help="do not fail when the committed ncu capture is of another kernel than the one launched")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import torch.distributed as dist

    import protstruc_b200 as ps  # noqa: F401
    from protstruc_b200 import _cabi, sharding
    from protstruc_b200.host_pipeline import HostFeaturePipeline, bind_host_thread_near_gpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()
    wl = WORKLOADS[args.workload]
    L = wl["L"]
    per_struct = bytes_per_structure(L)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---------------------------------------------------------------- device-resident arm (`value`)
    if args.workload == "m":
        B = args.batch                     # structures per launch
        n_local = B                        # structures this rank processes per step
        launches_per_step = 1
        structures_per_step_all_ranks = world * B
    else:
        lo, hi = sharding.shard_bounds(wl["total"], world, rank)
        n_local = hi - lo
        B = min(args.c5_chunk, max(n_local, 1))
        launches_per_step = (n_local + B - 1) // B
        structures_per_step_all_ranks = wl["total"]
    xyz, mask = synthetic_structures(max(n_local, 1), L, seed=1000 + rank, device=dev)
    dist_t = torch.empty(B, L, L, N_ATOM, N_ATOM, dtype=torch.float32, device=dev)
    mask_t = torch.empty(B, L, L, N_ATOM, N_ATOM, dtype=torch.bool, device=dev)
    omega = torch.empty(B, L, L, dtype=torch.float32, device=dev)
    theta = torch.empty_like(omega)
    phi = torch.empty_like(omega)
    stream = torch.cuda.current_stream(dev)

    def launch(first: int, n: int):
        rc = lib.ps_inter_residue_geometry(xyz[first:].data_ptr(), mask[first:].data_ptr(), _cabi.PS_MASK_BOOL,
                                           dist_t.data_ptr(), mask_t.data_ptr(), omega.data_ptr(), theta.data_ptr(),
                                           phi.data_ptr(), n, L, N_ATOM, stream.cuda_stream)
        _cabi.check(rc, "ps_inter_residue_geometry")

    def step():
        for first in range(0, n_local, B):
            launch(first, min(B, n_local - first))

    for _ in range(args.warmup):
        step()
    plan = _cabi.last_pair_dist_plan()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    events[0].record(stream)
    for k in range(args.steps):
        step()
        events[k + 1].record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = events[0].elapsed_time(events[-1])
    per_step_ms = [events[k].elapsed_time(events[k + 1]) for k in range(args.steps)]
    max_ms = max_over_ranks(total_ms)
    value = structures_per_step_all_ranks * args.steps / (max_ms / 1e3)

    # ---------------------------------------------------------------- end-to-end arm (`e2e`)
    # pinned host inputs -> GPU -> EVERY result byte back in pinned host memory, through the C-ABI host entry
    Be = min(args.e2e_batch, max(n_local, 1))
    all_cpus = os.sched_getaffinity(0)
    host_cpus = bind_host_thread_near_gpu(dev.index)  # host buffers next to this GPU's PCIe link (multi-rank runs)
    pipe = HostFeaturePipeline(chunk=2, L=L, A=N_ATOM, device=dev)
    xyz_h = xyz[:Be].cpu().pin_memory()
    mask_h = mask[:Be].cpu().pin_memory()
    out_h = HostFeaturePipeline.allocate_host_outputs(Be, L, N_ATOM, pinned=True)
    pipe.run(xyz_h, mask_h, out_h)  # warm-up (page-locks, first touch)
    barrier()
    # N = 1: `--e2e-structures` structures (>= 1 s).  N > 1: every rank makes whole calls for `--e2e-seconds`:
    # on a multi-GPU box the ranks' PCIe paths differ (this pool: 7.7 vs 18.5 GB/s per GPU when all eight stream), so a
    # fixed, equal number of calls per rank would time the slowest link while the others idle; value = all structures
    # of all ranks / the longest rank time, i.e. the aggregate throughput of the box.
    t0 = time.perf_counter()
    e2e_calls = 0
    while True:
        pipe.run(xyz_h, mask_h, out_h)
        e2e_calls += 1
        if (world == 1 and e2e_calls * Be >= args.e2e_structures) or \
                (world > 1 and time.perf_counter() - t0 >= args.e2e_seconds):
            break
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_total = sum_over_ranks(float(Be * e2e_calls))
    e2e_value = e2e_total / e2e_s
    # the end-to-end result on the host must be the device arm's result: EVERY byte of every tensor is compared
    launch(0, Be)
    torch.cuda.synchronize()
    for name, dev_t in (("dist", dist_t), ("dist_mask", mask_t), ("omega", omega), ("theta", theta), ("phi", phi)):
        host_bits = out_h[name][:Be].to(dev).contiguous().view(torch.uint8)
        if not torch.equal(host_bits, dev_t[:Be].contiguous().view(torch.uint8)):
            raise SystemExit(f"end-to-end host result `{name}` differs from the device result")
        del host_bits
    # raw device->host ceiling of this box right now: every rank streams the same bytes with plain cudaMemcpyAsync
    # (no kernel) into the same pinned buffer, all ranks at once (tools/d2h_ceiling.py is the long form)
    flat_h = out_h["dist"].view(-1)
    src = dist_t.view(-1)[: flat_h.numel()]
    flat_h.copy_(src, non_blocking=True)
    barrier()
    t0, copies = time.perf_counter(), 0
    while time.perf_counter() - t0 < 0.7:
        flat_h.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        copies += 1
    ceil_s = max_over_ranks(time.perf_counter() - t0)
    ceiling_gbs = sum_over_ranks(copies * flat_h.numel() * 4 / 1e9) / ceil_s
    os.sched_setaffinity(0, all_cpus)  # the CPU baseline below uses every core again
    d2h_per_structure = pipe.d2h_bytes(1)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes(Be), "d2h_bytes_per_step": pipe.d2h_bytes(Be),
           "steps": e2e_calls, "batch": Be, "seconds": e2e_s, "structures": e2e_total,
           "mode": "rank 0 made `steps` calls of `batch` structures; N > 1: every rank makes whole calls for >= e2e-seconds, "
                   "value = structures of all ranks / longest rank time",
           "d2h_gbs": e2e_value * d2h_per_structure / 1e9,
           "ceiling_gbs": ceiling_gbs, "frac_of_ceiling": e2e_value * d2h_per_structure / 1e9 / ceiling_gbs,
           "ceiling": "aggregate device->host rate of plain cudaMemcpyAsync into pinned memory, all ranks at once, no kernel "
                      "(measured in this run; tools/d2h_ceiling.py / profiles/r2*_d2h_ceiling*.json for the long form)",
           "verified": "every byte of dist, dist_mask, omega, theta, phi on the host equals the device arm's result",
           "host_cpus": "all" if host_cpus is None else f"{len(host_cpus)} local to the GPU (NVML affinity)",
           "api": "C-ABI ps_host_inter_residue_geometry via protstruc_b200.host_pipeline.HostFeaturePipeline.run "
                  "(pinned host in, pinned host out, chunks double-buffered on two streams)"}
    del out_h, flat_h, pipe

    # ---------------------------------------------------------------- optional: NVLink gather of compact features
    gather = None
    if args.with_gather:
        Bg = min(B, n_local)
        ring = [torch.empty(6, Bg, L, L, dtype=torch.float32, device=dev) for _ in range(2)]
        gathered = {"compact": torch.empty(world * 6 * Bg, L, L, dtype=torch.float32, device=dev)}

        def compact_launch(buf):
            rc = lib.ps_inter_residue_geometry_compact(xyz.data_ptr(), mask.data_ptr(), _cabi.PS_MASK_BOOL, dist_t.data_ptr(),
                                                       mask_t.data_ptr(), buf.data_ptr(), Bg, L, N_ATOM, stream.cuda_stream)
            _cabi.check(rc, "ps_inter_residue_geometry_compact")

        def do_gather(buf, async_op=False):
            # the six planes travel as ONE (6 Bg, L, L) tensor: a single all_gather_into_tensor per step
            return sharding.gather_compact_features({"compact": buf.view(6 * Bg, L, L)}, world * 6 * Bg, out=gathered,
                                                    async_op=async_op)

        reps = 10
        compact_launch(ring[0])
        do_gather(ring[0])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            compact_launch(ring[0])
        e1.record(stream)
        barrier()
        kernel_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        barrier()
        e0.record(stream)
        for _ in range(reps):
            do_gather(ring[0])
        e1.record(stream)
        barrier()
        gather_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        barrier()
        # overlapped: the gather of step k-1's planes runs on a SIDE stream (NCCL orders a collective after the stream it
        # is called on, so calling it on the compute stream would serialise it behind step k's kernel) while step k's
        # kernel fills the other ring slot; a slot is only rewritten after its gather has finished
        side = torch.cuda.Stream(dev)
        produced = [torch.cuda.Event() for _ in range(2)]
        gathered_ev = [torch.cuda.Event() for _ in range(2)]
        e0.record(stream)
        for k in range(reps + 1):
            slot = k % 2
            if k < reps:
                if k >= 2:
                    stream.wait_event(gathered_ev[slot])   # the gather that last read this slot is done
                compact_launch(ring[slot])
                produced[slot].record(stream)
            if k >= 1:
                prev = (k - 1) % 2
                with torch.cuda.stream(side):
                    side.wait_event(produced[prev])
                    do_gather(ring[prev])
                    gathered_ev[prev].record(side)
        stream.wait_stream(side)
        e1.record(stream)
        barrier()
        overlapped_ms = max_over_ranks(e0.elapsed_time(e1) / reps)
        # the same pipeline with SMs left free for NCCL: the fused kernel is persistent (one CTA per SM holding most of its
        # registers and shared memory), so the collective's CTAs can only run beside it on SMs it does not occupy
        overlapped_reserved = {}
        for reserve in (8, 16, 32):
            lib.ps_reserve_sms(reserve)
            barrier()
            e0.record(stream)
            for k in range(reps + 1):
                slot = k % 2
                if k < reps:
                    if k >= 2:
                        stream.wait_event(gathered_ev[slot])
                    compact_launch(ring[slot])
                    produced[slot].record(stream)
                if k >= 1:
                    prev = (k - 1) % 2
                    with torch.cuda.stream(side):
                        side.wait_event(produced[prev])
                        do_gather(ring[prev])
                        gathered_ev[prev].record(side)
            stream.wait_stream(side)
            e1.record(stream)
            barrier()
            overlapped_reserved[str(reserve)] = max_over_ranks(e0.elapsed_time(e1) / reps)
        lib.ps_reserve_sms(0)
        # the all-gather FUSED into the feature kernel: the angle warp stores the six compact values of its pair straight
        # into every rank's gathered buffer over NVLink peer memory (multicast through the NVSwitch where available)
        fused_push = {}
        if world > 1:
            import protstruc_b200 as ps_pkg

            sb_local = ps_pkg.StructureBatch.from_xyz(xyz[:Bg], mask[:Bg])
            for label, use_mc in (("multicast", True), ("unicast", False)):
                try:
                    fg = sharding.FusedFeatureGather(Bg, L, use_multicast=use_mc)
                    if use_mc and fg.multicast_ptr is None:
                        fused_push[label] = "no multicast support on this box"
                        continue
                    fg.run(sb_local, dist_out=dist_t[:Bg], dist_mask_out=mask_t[:Bg])
                    barrier()
                    e0.record(stream)
                    for _ in range(reps):
                        fg.run(sb_local, dist_out=dist_t[:Bg], dist_mask_out=mask_t[:Bg])
                    e1.record(stream)
                    barrier()
                    ms = max_over_ranks(e0.elapsed_time(e1) / reps)
                    do_gather(ring[0])   # reference result of the same inputs over NCCL
                    torch.cuda.synchronize()
                    want = gathered["compact"].view(world, 6, Bg, L, L).permute(1, 0, 2, 3, 4).contiguous()
                    same = torch.equal(want.view(torch.int32), fg.buffer.view(torch.int32))
                    fused_push[label] = {"ms_per_step": ms, "equals_nccl_gather": bool(same)}
                    del fg, want
                except Exception as exc:  # noqa: BLE001 - report what the box answered
                    fused_push[label] = f"{type(exc).__name__}: {exc}"
        sent = 6 * Bg * L * L * 4
        gather = {"structures_per_rank": Bg, "bytes_sent_per_rank": sent, "bytes_received_per_rank": sent * (world - 1),
                  "fused_kernel_with_compact_planes_ms": kernel_ms, "gather_alone_ms": gather_ms,
                  "kernel_plus_gather_overlapped_ms_per_step": overlapped_ms,
                  "overlapped_ms_per_step_by_sms_left_to_nccl": overlapped_reserved,
                  "fused_kernel_with_in_kernel_all_gather": fused_push,
                  "busbw_gbs": sent * (world - 1) / (gather_ms / 1e3) / 1e9 if world > 1 else None,
                  "nvlink_peak_gbs": {"nominal_per_direction": 900.0, "measured_peer_copy": 770.0},
                  "frac_of_nominal": sent * (world - 1) / (gather_ms / 1e3) / 1e9 / 900.0 if world > 1 else None,
                  "api": "StructureBatch.inter_residue_geometry_compact layout (6, B, L, L) -> "
                         "sharding.gather_compact_features (one ncclAllGather per step, NCCL's own stream)"}
        del ring, gathered

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peak, peak_src = hbm_peak()
    avg_step_ms = sum(per_step_ms) / len(per_step_ms)
    avg_launch_ms = avg_step_ms / launches_per_step
    achieved = n_local * per_struct / (avg_step_ms / 1e3) / 1e9
    # write-only ceiling on this GPU for context: a plain fill of the same distance buffer (library kernel)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist_t.zero_()
    e0.record()
    for _ in range(3):
        dist_t.zero_()
    e1.record()
    torch.cuda.synchronize()
    fill_gbs = 3 * dist_t.numel() * 4 / (e0.elapsed_time(e1) / 1e3) / 1e9
    headline_kernel = KERNEL_NAMES[1 if plan.get("sweep") else 0]
    traffic_per_structure, traffic_src = ncu_traffic_per_structure(headline_kernel, args.allow_stale_profile)
    if args.workload != "m" and traffic_per_structure is not None:
        traffic_per_structure, traffic_src = None, "no ncu capture at this shape (the committed one is at L = 512)"
    roofline = {
        "bound": "hbm", "kernel": f"{headline_kernel} (fused inter_residue_geometry)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": None if traffic_per_structure is None else traffic_per_structure * B, "traffic_source": traffic_src,
        "algorithmic_bytes_per_launch": B * per_struct, "structures_per_launch": B,
        "avg_launch_ms": avg_launch_ms, "best_step_ms": min(per_step_ms), "launches_per_step": launches_per_step,
        "fill_ceiling_gbs": fill_gbs,
        "schedule": {"lockstep": bool(plan["lockstep"]), "ctas": plan["ctas"], "tile_buffers": plan["tile_buffers"],
                     "active_buffers": plan["active_buffers"], "path": plan["path"],
                     "kernel": "linear sweep (pair_sweep.cu)" if plan.get("sweep") else "column strips (pair_dist.cu)"},
    }
    if plan["path"] != 0 or plan["launches"] != 1:
        raise SystemExit(f"the bench shape did not take ONE launch of the staged fused kernel: {plan}")

    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_reference_throughput(args.cpu_sample, threads, L)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{args.cpu_sample} structures of L={L}, A={N_ATOM} (oracle port of "
                                  f"inter_residue_geometry, torch CPU {threads} threads, {dt:.1f} s)"}

    line = {
        "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, B, L, world),
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e,
        "gpu_launches": args.steps * launches_per_step, "clocks": clocks, "impl": "ours",
    }
    if gather is not None:
        line["gather"] = gather
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
