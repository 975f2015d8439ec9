"""Host-buffer entry to the full pairwise feature set (the call a user of the reference makes).

`HostFeaturePipeline.run(xyz_host, atom_mask_host, out)` takes HOST arrays (pinned for full
speed), computes `inter_residue_geometry` on the GPU and leaves every result — the full
(B, L, L, A, A) distance tensor, its mask, omega / theta / phi — in HOST buffers, like the reference
does.  Structures are streamed through the GPU in chunks on two CUDA streams so that the
host->device copy, the kernel and the device->host copy of consecutive chunks overlap; the
device->host copy (~298 MB per 512-residue structure over PCIe) is what bounds this path.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import _cabi


class HostFeaturePipeline:
    def __init__(self, chunk: int, L: int, A: int = 15, device: Optional[torch.device] = None, n_slots: int = 2):
        if not torch.cuda.is_available():
            raise _cabi.NativeLibraryError("HostFeaturePipeline needs a CUDA device: there is no CPU fallback")
        self.lib = _cabi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.chunk, self.L, self.A = chunk, L, A
        dev = self.device
        self.slots = []
        for _ in range(n_slots):
            self.slots.append({
                "stream": torch.cuda.Stream(device=dev),
                "xyz": torch.empty(chunk, L, A, 3, dtype=torch.float32, device=dev),
                "mask": torch.empty(chunk, L, A, dtype=torch.bool, device=dev),
                "dist": torch.empty(chunk, L, L, A, A, dtype=torch.float32, device=dev),
                "dist_mask": torch.empty(chunk, L, L, A, A, dtype=torch.bool, device=dev),
                "omega": torch.empty(chunk, L, L, dtype=torch.float32, device=dev),
                "theta": torch.empty(chunk, L, L, dtype=torch.float32, device=dev),
                "phi": torch.empty(chunk, L, L, dtype=torch.float32, device=dev),
            })
        self.launches = 0

    @staticmethod
    def allocate_host_outputs(B: int, L: int, A: int = 15, pinned: bool = True) -> Dict[str, torch.Tensor]:
        mk = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=pinned)  # noqa: E731
        return {
            "dist": mk((B, L, L, A, A), torch.float32), "dist_mask": mk((B, L, L, A, A), torch.bool),
            "omega": mk((B, L, L), torch.float32), "theta": mk((B, L, L), torch.float32),
            "phi": mk((B, L, L), torch.float32),
        }

    def h2d_bytes(self, B: int) -> int:
        return B * self.L * self.A * (12 + 1)

    def d2h_bytes(self, B: int) -> int:
        return B * self.L * self.L * (self.A * self.A * 5 + 12)

    def run(self, xyz_host: torch.Tensor, mask_host: torch.Tensor, out: Dict[str, torch.Tensor]) -> None:
        """xyz_host (B, L, A, 3) fp32, mask_host (B, L, A) bool, out: dict from allocate_host_outputs.
        Returns after every result byte is in host memory."""
        B = xyz_host.shape[0]
        L, A = self.L, self.A
        if tuple(xyz_host.shape[1:]) != (L, A, 3) or mask_host.dtype != torch.bool:
            raise ValueError("host inputs must be (B, L, A, 3) fp32 and (B, L, A) bool for this pipeline")
        with torch.cuda.device(self.device):
            for k, start in enumerate(range(0, B, self.chunk)):
                n = min(self.chunk, B - start)
                slot = self.slots[k % len(self.slots)]
                s = slot["stream"]
                with torch.cuda.stream(s):
                    slot["xyz"][:n].copy_(xyz_host[start:start + n], non_blocking=True)
                    slot["mask"][:n].copy_(mask_host[start:start + n], non_blocking=True)
                    rc = self.lib.ps_inter_residue_geometry(
                        slot["xyz"].data_ptr(), slot["mask"].data_ptr(), _cabi.PS_MASK_BOOL,
                        slot["dist"].data_ptr(), slot["dist_mask"].data_ptr(), slot["omega"].data_ptr(),
                        slot["theta"].data_ptr(), slot["phi"].data_ptr(), n, L, A, s.cuda_stream)
                    _cabi.check(rc, "ps_inter_residue_geometry")
                    self.launches += 1
                    for name in ("dist", "dist_mask", "omega", "theta", "phi"):
                        out[name][start:start + n].copy_(slot[name][:n], non_blocking=True)
            for slot in self.slots:
                slot["stream"].synchronize()
