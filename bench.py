#!/usr/bin/env python
"""Benchmark of the geometric-feature hot path (BASELINE.json: structures/s on 512-residue x 15-atom
pairwise features, and achieved HBM GB/s against the measured peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the full pairwise feature set (distance matrix + pair mask + omega/theta/phi,
i.e. StructureBatch.inter_residue_geometry) over one batch of synthetic structures of L = 512, A = 15.
One fused kernel launch per step.  N > 1: launched by torchrun, one rank per GPU, the batch dimension
is partitioned (weak scaling: every GPU gets the same per-step batch), no collective on the data path.

Prints ONE JSON line (rank 0).  `value` = device-timed throughput with inputs resident in HBM;
`e2e` = the same metric through the host-buffer API (pinned host inputs, every result copied back to
host inside the timed region); `roofline` = algorithmic bytes / CUDA-event time of the fused kernel
against the measured HBM peak; `cpu_baseline` = the CPU oracle port (same ATen / numpy op sequence as
the reference) timed on this box's host cores on a bounded sample.

`--impl reference` times the reference's CPU algorithm (the oracle port; the reference is pure
Python, there is nothing to compile into oracle/_ref) on the same config / metric / unit.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

L_RES, N_ATOM = 512, 15
BYTES_PER_STRUCT = L_RES * L_RES * (N_ATOM * N_ATOM * 5 + 12) + L_RES * N_ATOM * 13  # 298,157,568
METRIC = "structures/sec, 512-res x 15-atom pairwise features (dist + mask + omega/theta/phi)"
UNIT = "structures/s"
FALLBACK_HBM_GBS = 6650.0


def synthetic_structures(B: int, seed: int, device):
    """xyz ~ 10 N(0,1) A, Bernoulli(0.5) bool mask, masked slots NaN."""
    g = torch.Generator(device=device).manual_seed(seed)
    xyz = 10.0 * torch.randn(B, L_RES, N_ATOM, 3, device=device, generator=g)
    mask = torch.rand(B, L_RES, N_ATOM, device=device, generator=g) < 0.5
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
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def hbm_peak():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except (KeyError, ValueError):
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def cpu_reference_throughput(n_structs: int, threads: int):
    """The CPU oracle port of inter_residue_geometry on `n_structs` structures of the bench shape,
    one structure at a time (the reference needs ~17 B of temporaries per distance element)."""
    from oracle import feature_oracle as orc

    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1)
    xyz = 10.0 * torch.randn(n_structs, L_RES, N_ATOM, 3, generator=g)
    mask = torch.rand(n_structs, L_RES, N_ATOM, generator=g) < 0.5
    xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
    t0 = time.perf_counter()
    for b in range(n_structs):
        out = orc.inter_residue_geometry(xyz[b:b + 1], mask[b:b + 1])
        del out
    dt = time.perf_counter() - t0
    return n_structs / dt, dt


def run_reference_arm(args, rank: int, world: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = 4
    for _ in range(args.warmup):
        cpu_reference_throughput(per_step, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_throughput(per_step, threads)
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = f"{per_step} structure(s) of L={L_RES}, A={N_ATOM} per step, {args.steps} steps, torch CPU {threads} threads"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"inter_residue_geometry L={L_RES} A={N_ATOM}, batch {per_step}/step (bounded sample)",
                   "note": "CPU oracle port: same ATen/numpy op sequence as the pure-Python reference "
                           "(bit-identical to it, tests/golden/MANIFEST.json)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=16, help="structures per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=4, help="structures per GPU per end-to-end step")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=96, help="structures timed for the CPU baseline (~10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    import protstruc_b200 as ps
    from protstruc_b200 import _cabi
    from protstruc_b200.host_pipeline import HostFeaturePipeline, bind_host_thread_near_gpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident arm (`value`)
    B = args.batch
    xyz, mask = synthetic_structures(B, seed=1000 + rank, device=dev)
    dist_t = torch.empty(B, L_RES, L_RES, N_ATOM, N_ATOM, dtype=torch.float32, device=dev)
    mask_t = torch.empty(B, L_RES, L_RES, N_ATOM, N_ATOM, dtype=torch.bool, device=dev)
    omega = torch.empty(B, L_RES, L_RES, dtype=torch.float32, device=dev)
    theta = torch.empty_like(omega)
    phi = torch.empty_like(omega)
    stream = torch.cuda.current_stream(dev)

    def step():
        rc = lib.ps_inter_residue_geometry(xyz.data_ptr(), mask.data_ptr(), _cabi.PS_MASK_BOOL, dist_t.data_ptr(),
                                           mask_t.data_ptr(), omega.data_ptr(), theta.data_ptr(), phi.data_ptr(),
                                           B, L_RES, N_ATOM, stream.cuda_stream)
        _cabi.check(rc, "ps_inter_residue_geometry")

    for _ in range(args.warmup):
        step()
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
    per_launch_ms = [events[k].elapsed_time(events[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * B * args.steps / (max_ms / 1e3)

    # ---------------------------------------------------------------- end-to-end arm (`e2e`)
    Be = args.e2e_batch
    all_cpus = os.sched_getaffinity(0)
    host_cpus = bind_host_thread_near_gpu(dev.index)  # host buffers next to this GPU's PCIe link (multi-rank runs)
    pipe = HostFeaturePipeline(chunk=2, L=L_RES, A=N_ATOM, device=dev)
    xyz_h = xyz[:Be].cpu().pin_memory()
    mask_h = mask[:Be].cpu().pin_memory()
    out_h = HostFeaturePipeline.allocate_host_outputs(Be, L_RES, N_ATOM, pinned=True)
    pipe.run(xyz_h, mask_h, out_h)  # warm-up (page-locks, first-touch)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        pipe.run(xyz_h, mask_h, out_h)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * args.e2e_steps / float(te.item())
    # the end-to-end result on the host must be the kernel's result
    same = torch.equal(torch.nan_to_num(out_h["dist"][0, :4, :4]), torch.nan_to_num(dist_t[0, :4, :4].cpu()))
    if not same:
        raise SystemExit("end-to-end host result differs from the device result")
    os.sched_setaffinity(0, all_cpus)  # the CPU baseline below uses every core again

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peak, peak_src = hbm_peak()
    avg_launch_ms = sum(per_launch_ms) / len(per_launch_ms)
    achieved = B * BYTES_PER_STRUCT / (avg_launch_ms / 1e3) / 1e9
    # write-only ceiling on this GPU for context: a plain fill of the same distance buffer (library kernel)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist_t.zero_()
    e0.record()
    for _ in range(3):
        dist_t.zero_()
    e1.record()
    torch.cuda.synchronize()
    fill_gbs = 3 * dist_t.numel() * 4 / (e0.elapsed_time(e1) / 1e3) / 1e9
    # DRAM bytes per launch from the committed `ncu --set full` capture of this command
    # (profiles/r1t_k1_ncu_summary.txt: dram__bytes_write 4.709095 GB + dram__bytes_read 4.61696 MB at 16 structures)
    traffic = (4.709095e9 + 4.61696e6) * B / 16 if B == 16 else None
    roofline = {
        "bound": "hbm", "kernel": "pair_tiles_kernel<15, dist+boolmask, angles> (fused inter_residue_geometry)",
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
        "traffic": traffic, "traffic_source": "profiles/r1t_k1_ncu_summary.txt (ncu --set full, per launch)",
        "algorithmic_bytes_per_launch": B * BYTES_PER_STRUCT,
        "avg_launch_ms": avg_launch_ms, "best_launch_ms": min(per_launch_ms),
        "fill_ceiling_gbs": fill_gbs,
    }

    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, dt = cpu_reference_throughput(args.cpu_sample, threads)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{args.cpu_sample} structures of L={L_RES}, A={N_ATOM} (oracle port of "
                                  f"inter_residue_geometry, torch CPU {threads} threads, {dt:.1f} s)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": max_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"inter_residue_geometry (full pairwise feature set), L={L_RES}, A={N_ATOM}, "
                               f"{B} structures per GPU per step",
                   "bytes_per_structure": BYTES_PER_STRUCT,
                   "l2_policy": f"outputs of one step ({B * BYTES_PER_STRUCT / 1e9:.2f} GB per GPU) exceed the 126 MB L2; "
                                "no explicit flush", "parallelism": f"batch-sharded x{world}, no collective"},
        "roofline": roofline, "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes(Be),
                "d2h_bytes_per_step": pipe.d2h_bytes(Be), "steps": args.e2e_steps, "batch": Be,
                "host_cpus": "all" if host_cpus is None else f"{len(host_cpus)} local to the GPU (NVML affinity)",
                "api": "C-ABI ps_host_inter_residue_geometry via protstruc_b200.host_pipeline.HostFeaturePipeline.run "
                       "(pinned host in, pinned host out, chunks double-buffered on two streams)"},
        "gpu_launches": args.steps, "clocks": clocks, "impl": "ours",
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
