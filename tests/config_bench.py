#!/usr/bin/env python
"""Times the five BASELINE.json configurations through the public API (run on the GPU box).

    python tests/config_bench.py [--no-cpu] > gpurun_out/config_bench.json
    python -m torch.distributed.run --nproc-per-node N ... tests/config_bench.py --only c5   # sharded config 5

Every GPU number is CUDA-event time of the façade call(s) with inputs resident; every CPU number is the oracle
port (same ATen/numpy ops as the reference) on a bounded sample of the same workload, normalised per structure.
"""
import argparse
import json
import os
import statistics
import sys
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import protstruc_b200 as ps  # noqa: E402
from protstruc_b200 import _cabi  # noqa: E402
from protstruc_b200.sharding import shard_bounds  # noqa: E402
from oracle import feature_oracle as orc  # noqa: E402
from tests import helpers as H  # noqa: E402


def gpu_time(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return {"best_ms": min(ts), "median_ms": statistics.median(ts)}


def cpu_time(fn, repeats=2):
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best * 1e3


def synthetic(B, L, A, seed, dev, nan_masked=True):
    g = torch.Generator(device=dev).manual_seed(seed)
    xyz = 10.0 * torch.randn(B, L, A, 3, device=dev, generator=g)
    mask = torch.rand(B, L, A, device=dev, generator=g) < 0.5
    if nan_masked:
        xyz = torch.where(mask[..., None], xyz, torch.full_like(xyz, float("nan")))
    return xyz.contiguous(), mask


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    torch.set_num_threads(os.cpu_count() or 1)
    out = {"device": torch.cuda.get_device_name(dev), "world_size": world, "host_cores": os.cpu_count(), "configs": {}}
    want = lambda name: (not args.only) or name in args.only.split(",")  # noqa: E731

    if want("c1") and rank == 0:
        g = H.load_golden("real_1a6v_HL")
        sb = ps.StructureBatch.from_xyz(g["xyz"], g["atom_mask"], g["chain_idx"], [["L", "H"]])
        r = {"what": "real structure tests/1a6v_HL.pdb (L=229, 1734 atoms): pairwise_distance_matrix + backbone_dihedrals"}
        r["gpu"] = gpu_time(lambda: (sb.pairwise_distance_matrix(), sb.backbone_dihedrals()))
        r["gpu_inter_residue_geometry"] = gpu_time(lambda: sb.inter_residue_geometry())
        # the same two calls captured once in a CUDA graph and replayed (no Python between the launches)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            sb.pairwise_distance_matrix(), sb.backbone_dihedrals()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            captured = (sb.pairwise_distance_matrix(), sb.backbone_dihedrals())  # noqa: F841 - keeps the outputs alive
        r["gpu_cuda_graph_replay"] = gpu_time(graph.replay)
        if not args.no_cpu:
            xyz, mask, ch = H.t(g["xyz"]), H.t(g["atom_mask"]), H.t(g["chain_idx"])
            r["cpu_ms"] = cpu_time(lambda: (orc.pair_distances(xyz, mask), orc.backbone_dihedrals(xyz, ch, mask.any(-1))))
            r["cpu_inter_residue_geometry_ms"] = cpu_time(lambda: orc.inter_residue_geometry(xyz, mask))
        out["configs"]["c1"] = r

    if want("c2") and rank == 0:
        B, L, A = 64, 256, 15
        xyz, mask = synthetic(B, L, A, 2, dev)
        sb = ps.StructureBatch.from_xyz(xyz, mask)
        r = {"what": "synthetic 64 x 256 x 15: pairwise_distance_matrix (dist + bool mask)", "bytes": B * (L * L * A * A * 5 + L * A * 13)}
        r["gpu"] = gpu_time(lambda: sb.pairwise_distance_matrix())
        r["gpu_GBps"] = r["bytes"] / r["gpu"]["best_ms"] / 1e6
        r["gpu_structures_per_s"] = B / r["gpu"]["best_ms"] * 1e3
        if not args.no_cpu:
            n = 4
            x, m = xyz[:n].cpu(), mask[:n].cpu()
            ms = cpu_time(lambda: orc.pair_distances(x, m))
            r["cpu_structures_per_s"] = n / ms * 1e3
            r["cpu_sample"] = f"{n} structures"
        del sb
        out["configs"]["c2"] = r

    if want("c3") and rank == 0:
        B, L, A = 256, 512, 5
        xyz, mask = synthetic(B, L, A, 3, dev, nan_masked=False)
        sb = ps.StructureBatch.from_xyz(xyz, mask)
        r = {"what": "synthetic 256 x 512 backbone (N,CA,C,O,CB): omega + theta + phi", "bytes": B * (L * L * 12 + L * A * 12)}
        r["gpu_fused"] = gpu_time(lambda: sb.trrosetta_angles())
        r["gpu_three_calls"] = gpu_time(lambda: (sb.pairwise_dihedrals(["CA", "CB"], ["CA", "CB"]),
                                                 sb.pairwise_dihedrals(["N", "CA", "CB"], ["CB"]),
                                                 sb.pairwise_planar_angles(["CA", "CB"], ["CB"])))
        r["gpu_structures_per_s"] = B / r["gpu_fused"]["best_ms"] * 1e3
        r["gpu_GBps"] = r["bytes"] / r["gpu_fused"]["best_ms"] / 1e6
        r["binding_roof"] = "FP32/SFU issue"
        if not args.no_cpu:
            n = 2
            x = xyz[:n].cpu()
            ms = cpu_time(lambda: orc.trrosetta_angles(x), repeats=1)
            r["cpu_structures_per_s"] = n / ms * 1e3
            r["cpu_sample"] = f"{n} structures"
        del sb
        out["configs"]["c3"] = r

    if want("c4") and rank == 0:
        B, L, A, T = 1024, 128, 15, 300
        xyz, mask = synthetic(B, L, A, 4, dev)
        betas = orc.cosine_variance_schedule(T)[:T].to(dev)[:, None].repeat(1, B).contiguous()
        r = {"what": "standardize + 300 x diffuse_xyz (cosine schedule) on 1024 x 128 x 15"}

        def loop_api():
            sb = ps.StructureBatch.from_xyz(xyz, mask)
            sb.standardize()
            for t in range(T):
                sb.diffuse_xyz(betas[t])
            return sb

        def fused_api():
            sb = ps.StructureBatch.from_xyz(xyz, mask)
            sb.standardize()
            sb.diffuse_xyz_steps(betas)
            return sb
        r["gpu_300_calls"] = gpu_time(loop_api, iters=3, warmup=1)
        r["gpu_fused_steps"] = gpu_time(fused_api, iters=3, warmup=1)
        r["gpu_trajectories_per_s_300_calls"] = B / r["gpu_300_calls"]["best_ms"] * 1e3
        r["gpu_trajectories_per_s_fused"] = B / r["gpu_fused_steps"]["best_ms"] * 1e3
        final = fused_api().get_xyz()
        valid = final[mask]
        r["final_std_of_valid_atoms"] = float(valid.std())
        if not args.no_cpu:
            n = 16
            x, m, bt = xyz[:n].cpu(), mask[:n].cpu(), betas[:, :n].cpu()

            def cpu_loop():
                cur, _, _ = orc.standardize_per_structure(x, m)
                for t in range(T):
                    cur = orc.diffuse(cur, bt[t], torch.randn_like(cur))
            ms = cpu_time(cpu_loop, repeats=1)
            r["cpu_trajectories_per_s"] = n / ms * 1e3
            r["cpu_sample"] = f"{n} structures x {T} steps"
        out["configs"]["c4"] = r

    if want("c5"):
        B_total, L, A, chunk = 4096, 384, 15, 64  # bench.py's default chunk (profiles/r4e_bench_by_batch.txt)
        start, stop = shard_bounds(B_total, world, rank)
        n_local = stop - start
        lib = _cabi.load()
        xyz, mask = synthetic(chunk, L, A, 5 + rank, dev)
        dist = torch.empty(chunk, L, L, A, A, device=dev)
        dmask = torch.empty(chunk, L, L, A, A, dtype=torch.bool, device=dev)
        om = torch.empty(chunk, L, L, device=dev)
        th, ph = torch.empty_like(om), torch.empty_like(om)
        s = torch.cuda.current_stream().cuda_stream

        def one_chunk(n):
            _cabi.check(lib.ps_inter_residue_geometry(xyz.data_ptr(), mask.data_ptr(), 0, dist.data_ptr(), dmask.data_ptr(),
                                                      om.data_ptr(), th.data_ptr(), ph.data_ptr(), n, L, A, s), "c5")
        one_chunk(chunk)
        torch.cuda.synchronize()
        if world > 1:
            dist_mod = torch.distributed
            dist_mod.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done = 0
        while done < n_local:
            n = min(chunk, n_local - done)
            one_chunk(n)  # same synthetic chunk re-used as input; outputs stream through one reused buffer
            done += n
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        per_struct = L * L * (A * A * 5 + 12) + L * A * 13
        if rank == 0:
            out["configs"]["c5"] = {
                "what": f"4096 x 384 x 15 full pairwise feature set, batch-sharded over {world} GPU(s), chunks of {chunk} "
                        f"structures through a reused {chunk * L * L * A * A * 5 / 1e9:.1f} GB output buffer",
                "total_bytes": B_total * per_struct, "ms_max_over_ranks": float(ms.item()),
                "structures_per_s": B_total / float(ms.item()) * 1e3,
                "aggregate_GBps": B_total * per_struct / float(ms.item()) / 1e6,
            }
    if rank == 0:
        print(json.dumps(out, indent=1))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
