#!/usr/bin/env python
"""Sharded execution == unsharded execution, bit for bit, over NCCL (run under torchrun on >= 2 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py

Every rank builds its shard from the same seeded global arrays, computes the pairwise feature set, the backbone
features and two diffusion steps, all-gathers the compact features over NCCL and compares them with the result
of the whole batch computed on its own GPU.
"""
import json
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import protstruc_b200 as ps  # noqa: E402
from protstruc_b200 import sharding  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl")
    B, L, A = 2 * world + 1, 96, 15  # uneven shards on purpose
    g = torch.Generator().manual_seed(77)
    xyz = 10.0 * torch.randn(B, L, A, 3, generator=g)
    mask = torch.rand(B, L, A, generator=g) < 0.6
    mask[:, :, :5] = True
    chain_idx = torch.zeros(B, L)
    chain_idx[:, L // 2:] = 1.0
    ids = [["A", "B"]] * B
    beta = torch.linspace(0.01, 0.2, B)

    def features(sb, beta_rows):
        out = sb.inter_residue_geometry()
        dih, dmask = sb.backbone_dihedrals()
        feats = {k: out[k].contiguous() for k in ("omega", "theta", "phi", "d_ca", "d_cb", "d_no")}
        feats["dihedrals"], feats["dihedral_mask"] = dih, dmask.to(torch.uint8)
        feats["frames"] = sb.backbone_orientations()
        ps.manual_seed(2024)
        sb.diffuse_xyz(beta_rows.to("cuda"))
        sb.diffuse_xyz(beta_rows.to("cuda"))
        feats["diffused"] = sb.get_xyz().contiguous()
        return feats

    shard = sharding.shard_structure_batch(xyz, mask, chain_idx, ids, device="cuda")
    start, stop = sharding.shard_bounds(B, world, rank)
    local = features(shard, beta[start:stop])
    gathered = sharding.gather_compact_features(local, B)
    whole = features(ps.StructureBatch.from_xyz(xyz, mask, chain_idx, ids, device="cuda"), beta)
    report = {}
    # the all-gather fused into the feature kernel (symmetric memory over NVLink): needs A = 15, L >= 32
    try:
        fused = sharding.FusedFeatureGather(max(sharding.shard_sizes(B, world)), L)
        shard2 = sharding.shard_structure_batch(xyz, mask, chain_idx, ids, device="cuda")
        pushed = fused.run(shard2)
        torch.cuda.synchronize()
        rows = torch.cat([torch.arange(r * fused.shard, r * fused.shard + n) for r, n in enumerate(sharding.shard_sizes(B, world))])
        for k in sharding.COMPACT_FEATURES:
            got = pushed[k].index_select(0, rows.to("cuda"))
            report["fused_push_" + k] = bool(torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(whole[k].float(), nan=-7.0)))
        report["fused_push_multicast"] = fused.multicast_ptr is not None
        if fused.multicast_ptr is not None:  # and once more without the NVSwitch multicast (one store per peer)
            fused_uc = sharding.FusedFeatureGather(max(sharding.shard_sizes(B, world)), L, use_multicast=False)
            pushed_uc = fused_uc.run(shard2)
            torch.cuda.synchronize()
            report["fused_push_unicast_all"] = all(
                torch.equal(torch.nan_to_num(pushed_uc[k].index_select(0, rows.to("cuda")), nan=-7.0),
                            torch.nan_to_num(whole[k].float(), nan=-7.0)) for k in sharding.COMPACT_FEATURES)
    except Exception as exc:  # noqa: BLE001 - report, do not hide
        report["fused_push_error"] = f"{type(exc).__name__}: {exc}"
    report_ok = {k: v for k, v in report.items() if isinstance(v, bool) and k != "fused_push_multicast"}
    for k in sorted(whole):
        same = torch.equal(torch.nan_to_num(gathered[k].float(), nan=-7.0), torch.nan_to_num(whole[k].float(), nan=-7.0))
        report[k] = bool(same)
    report_ok.update({k: v for k, v in report.items() if isinstance(v, bool) and k != "fused_push_multicast"})
    ok = torch.tensor([int(all(report_ok.values()) and "fused_push_error" not in report)], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world_size": world, "backend": dist.get_backend(), "batch": B, "shards": sharding.shard_sizes(B, world),
                          "bit_identical": report, "all_ranks_ok": bool(ok.item())}))
    dist.destroy_process_group()
    sys.exit(0 if ok.item() else 1)


if __name__ == "__main__":
    main()
