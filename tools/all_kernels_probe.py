#!/usr/bin/env python
"""Launches every kernel family once at its BASELINE.json size (target for one `ncu --set full` pass).

    python tools/all_kernels_probe.py            # plain run
    ncu --set full --clock-control none -k regex:'backbone|center_of_mass|diffuse_|kabsch|local_xyz|masked_stats|pair_|rotate_kernel|scale_shift|translate|trrosetta' -o gpurun_out/all_kernels python tools/all_kernels_probe.py
"""
import sys
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))

import protstruc_b200 as ps  # noqa: E402

DEV = "cuda"
g = torch.Generator(device=DEV).manual_seed(0)


def batch(B, L, A, p=0.5):
    xyz = 10.0 * torch.randn(B, L, A, 3, device=DEV, generator=g)
    mask = torch.rand(B, L, A, device=DEV, generator=g) < p
    mask[:, :, :5] = True
    return ps.StructureBatch.from_xyz(xyz, mask)


# K1 fused (bench shape, smaller batch to keep the capture short), K1 distances + mask (config 2 shape, 16 structures)
sb = batch(16, 512, 15)
sb.inter_residue_geometry()
sb.pairwise_distance_matrix()
# any-A tile kernel: A = 25 (the reference tests' atom count; unrolled instantiation), A = 37 (atom37), A = 20 (run-time A)
batch(24, 256, 25).pairwise_distance_matrix()
if "--all-atom-counts" in sys.argv:  # (left out by default: the .ncu-rep of the whole list must stay below 64 MB)
    batch(12, 256, 37).pairwise_distance_matrix()
    batch(40, 256, 20).pairwise_distance_matrix()
# K2f / K2 (config 3: backbone + CB)
c3 = batch(256, 512, 5, p=1.1)
c3.trrosetta_angles()
c3.pairwise_dihedrals(["N", "CA", "C"], ["N"])
c3.pairwise_planar_angles(["CA", "CB"], ["CB"])
# backbone-only batch through the distance / fused calls: 5-atom strip kernel (eight tile buffers), then the split
# dispatch of inter_residue_geometry (distance tiles + exact-sequence angle kernel)
c3.pairwise_distance_matrix()
c3.inter_residue_geometry()
# K3 / K4 / f1 / f2 at 256 x 512 x 15
big = batch(256, 512, 15, p=0.7)
big.backbone_dihedrals()
big.backbone_orientations()
big.center_of_mass()
big.get_local_xyz()
big.rotate(torch.eye(3, device=DEV).expand(256, 3, 3).contiguous())
big.translate(torch.ones(256, 1, 3, device=DEV))
big.align(batch(256, 512, 15, p=0.7))
big.standardize()
big.unstandardize()
# K5 (config 4)
c4 = batch(1024, 128, 15)
c4.standardize()
beta = torch.full((1024,), 0.02, device=DEV)
c4.diffuse_xyz(beta)
c4.diffuse_xyz(beta, noise=torch.randn_like(c4.get_xyz()))
c4.diffuse_xyz_steps(torch.full((300, 1024), 0.02, device=DEV))
torch.cuda.synchronize()
print("ok")
