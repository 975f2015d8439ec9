"""Host-buffer entry to the full pairwise feature set (the call a user of the reference makes).

`HostFeaturePipeline.run(xyz_host, atom_mask_host, out)` takes HOST tensors (pinned for full speed),
computes `inter_residue_geometry` on the GPU and leaves every result — the full (B, L, L, A, A) distance
tensor, its mask, omega / theta / phi — in HOST buffers, like the reference does.  It is a thin wrapper
over the native pipeline behind `ps_host_inter_residue_geometry` (protstruc_b200/csrc/host_pipeline.cu):
chunks double-buffered on two CUDA streams so uploads, the fused kernel and downloads overlap; the
device->host copy (~298 MB per 512-residue structure over PCIe) is what bounds this path.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, List, Optional

import torch

from . import _cabi


def bind_host_thread_near_gpu(device_index: Optional[int] = None) -> Optional[List[int]]:
    """Restricts the calling thread to the CPUs NVML reports as local to the GPU (same NUMA node / PCIe root
    complex).  Pinned host buffers allocated afterwards are first-touched on that node, so the device->host stream
    of this path (298 MB per structure) does not cross the socket interconnect — which matters once several ranks
    of one host stream results at the same time.  Opt-in: call it once per rank before allocating host buffers.
    Returns the CPU list, or None when NVML gives no usable answer (then nothing is changed)."""
    try:
        import pynvml

        index = torch.cuda.current_device() if device_index is None else int(device_index)
        props = torch.cuda.get_device_properties(index)
        bus_id = f"{props.pci_domain_id:08x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus_id.encode())
        allowed = sorted(os.sched_getaffinity(0))
        words = (max(allowed) // 64) + 1
        masks = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        local = [c for c in allowed if (int(masks[c // 64]) >> (c % 64)) & 1]
        if not local or len(local) == len(allowed):
            return None
        os.sched_setaffinity(0, local)
        return local
    except Exception:  # noqa: BLE001 - no NVML, no affinity support, container restrictions: leave things alone
        return None


class HostFeaturePipeline:
    def __init__(self, chunk: int, L: int, A: int = 15, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _cabi.NativeLibraryError("HostFeaturePipeline needs a CUDA device: there is no CPU fallback")
        self.lib = _cabi.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.chunk, self.L, self.A = chunk, L, A
        handle = ctypes.c_void_p()
        with _cabi.on_device(self.device):
            _cabi.check(self.lib.ps_host_pipeline_create(chunk, L, A, ctypes.byref(handle)), "ps_host_pipeline_create")
        self._handle = handle

    def close(self) -> None:
        if getattr(self, "_handle", None):
            self.lib.ps_host_pipeline_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.ps_host_pipeline_launches(self._handle))

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
        """xyz_host (B, L, A, 3) fp32, mask_host (B, L, A) bool, out: dict from allocate_host_outputs — all on the
        host.  Returns after every result byte is in host memory."""
        B = xyz_host.shape[0]
        L, A = self.L, self.A
        if xyz_host.is_cuda or mask_host.is_cuda or tuple(xyz_host.shape[1:]) != (L, A, 3) or \
                xyz_host.dtype != torch.float32 or mask_host.dtype != torch.bool or tuple(mask_host.shape) != (B, L, A):
            raise ValueError("host inputs must be CPU tensors of shape (B, L, A, 3) fp32 and (B, L, A) bool")
        xyz_host, mask_host = xyz_host.contiguous(), mask_host.contiguous()
        full, plane = (L, L, A, A), (L, L)
        for name, dt, tail in (("dist", torch.float32, full), ("dist_mask", torch.bool, full),
                               ("omega", torch.float32, plane), ("theta", torch.float32, plane),
                               ("phi", torch.float32, plane)):
            t = out[name]
            if t.is_cuda or t.dtype != dt or not t.is_contiguous() or t.ndim != 1 + len(tail) or t.shape[0] < B or \
                    tuple(t.shape[1:]) != tail:
                raise ValueError(f"out[{name!r}] must be a contiguous CPU {dt} tensor of shape (>= {B}, "
                                 f"{', '.join(map(str, tail))}), got {tuple(t.shape)}")
        with _cabi.on_device(self.device):
            rc = self.lib.ps_host_inter_residue_geometry(
                self._handle, xyz_host.data_ptr(), mask_host.data_ptr(), B, out["dist"].data_ptr(),
                out["dist_mask"].data_ptr(), out["omega"].data_ptr(), out["theta"].data_ptr(), out["phi"].data_ptr())
        _cabi.check(rc, "ps_host_inter_residue_geometry")
