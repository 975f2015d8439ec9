"""Free-function geometry on point lists, CUDA-backed.

Mirrors the reference's `protstruc.geometry` surface for the hot path — `dot`, `norm`, `unit`,
`angle`, `dihedral`, `gram_schmidt` (reference protstruc/geometry.py:24-124, 413-439) — including
the type contract of its `with_tensor` adaptor (reference protstruc/decorator.py:5-53):
numpy-only inputs give numpy outputs, any torch input gives torch outputs, numpy floats become fp32.

Every function launches a hand-written kernel through the C-ABI (ps_geom_dot / ps_geom_norm / ps_geom_unit /
ps_geom_angle / ps_geom_dihedral / ps_geom_gram_schmidt).  Nothing here computes on the CPU: inputs are moved to the
current CUDA device, and without a GPU (or without the built library) the calls raise.
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch

from . import _cabi

ArrayLike = Union[np.ndarray, torch.Tensor]


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise _cabi.NativeLibraryError(
            "protstruc_b200.geometry needs a CUDA device: there is no CPU fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def _ingest(args):
    """Applies the adaptor rules; returns (tensors_on_gpu, saw_tensor, out_device)."""
    saw_tensor = False
    out_device = None
    prepared = []
    for a in args:
        if isinstance(a, np.ndarray):
            t = torch.tensor(a)
            if a.dtype in (np.float32, np.float64):
                t = t.float()
            prepared.append(t)
        elif isinstance(a, torch.Tensor):
            saw_tensor = True
            if out_device is None:
                out_device = a.device
            prepared.append(a)
        else:
            raise TypeError(f"expected numpy.ndarray or torch.Tensor, got {type(a).__name__}")
    dev = next((t.device for t in prepared if t.is_cuda), None) or _device()
    return [t.to(dev) for t in prepared], saw_tensor, dev


def _egress(out: torch.Tensor, saw_tensor: bool):
    return out if saw_tensor else out.cpu().numpy()


def _points(tensors):
    """Broadcasts point arrays against each other and flattens to contiguous fp32 (n, 3)."""
    for t in tensors:
        if t.shape[-1] != 3:
            raise ValueError(f"points must have a trailing dimension of 3, got shape {tuple(t.shape)}")
    shape = torch.broadcast_shapes(*[t.shape for t in tensors])
    flat = [t.to(torch.float32).expand(shape).reshape(-1, 3).contiguous() for t in tensors]
    return flat, shape[:-1]


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _rows(tensors):
    """Broadcasts arrays against each other and flattens to contiguous fp32 (n, D) rows."""
    shape = torch.broadcast_shapes(*[t.shape for t in tensors])
    if len(shape) == 0:
        raise ValueError("expected at least one dimension")
    flat = [t.to(torch.float32).expand(shape).reshape(-1, shape[-1]).contiguous() for t in tensors]
    return flat, shape


def dot(x: ArrayLike, y: ArrayLike):
    """Inner product over the last axis, keepdim (reference geometry.py:24-26)."""
    tensors, saw, dev = _ingest([x, y])
    (fx, fy), shape = _rows(tensors)
    out = torch.empty(fx.shape[0], dtype=torch.float32, device=dev)
    if shape[-1] == 0:
        out.zero_()
    else:
        with _cabi.on_device(dev):
            rc = _cabi.load().ps_geom_dot(fx.data_ptr(), fy.data_ptr(), fx.shape[0], shape[-1], out.data_ptr(), _stream(dev))
        _cabi.check(rc, "ps_geom_dot")
    out = out.reshape(shape[:-1] + (1,))
    kind = torch.result_type(tensors[0], tensors[1])
    if not kind.is_floating_point:  # integer inputs give an integer product sum, like `(x * y).sum`
        out = out.round().to(torch.int64 if kind != torch.bool else torch.int64)
    return _egress(out, saw)


def norm(x: ArrayLike):
    """Euclidean norm over the last axis, keepdim (reference geometry.py:29-31)."""
    tensors, saw, dev = _ingest([x])
    (fx,), shape = _rows(tensors)
    out = torch.empty(fx.shape[0], dtype=torch.float32, device=dev)
    if shape[-1] == 0:
        out.zero_()
    else:
        with _cabi.on_device(dev):
            rc = _cabi.load().ps_geom_norm(fx.data_ptr(), fx.shape[0], shape[-1], out.data_ptr(), _stream(dev))
        _cabi.check(rc, "ps_geom_norm")
    return _egress(out.reshape(shape[:-1] + (1,)), saw)


def unit(x: ArrayLike):
    """x / |x| (reference geometry.py:34-36)."""
    tensors, saw, dev = _ingest([x])
    (fx,), shape = _rows(tensors)
    out = torch.empty_like(fx)
    if shape[-1] != 0:
        with _cabi.on_device(dev):
            rc = _cabi.load().ps_geom_unit(fx.data_ptr(), fx.shape[0], shape[-1], out.data_ptr(), _stream(dev))
        _cabi.check(rc, "ps_geom_unit")
    return _egress(out.reshape(shape), saw)


def angle(a: ArrayLike, b: ArrayLike, c: ArrayLike, to_degree: bool = False):
    """Planar angle a-b-c in [0, pi] (or degrees); arccos without clamping, like the reference
    (geometry.py:39-71).  Shapes (*, 3) -> (*)."""
    tensors, saw, dev = _ingest([a, b, c])
    (pa, pb, pc), lead = _points(tensors)
    n = pa.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with _cabi.on_device(dev):
        rc = _cabi.load().ps_geom_angle(pa.data_ptr(), pb.data_ptr(), pc.data_ptr(), n,
                                        int(bool(to_degree)), out.data_ptr(), _stream(dev))
    _cabi.check(rc, "ps_geom_angle")
    return _egress(out.reshape(lead), saw)


def dihedral(a: ArrayLike, b: ArrayLike, c: ArrayLike, d: ArrayLike, to_degree: bool = False):
    """Dihedral angle of a-b-c-d in [-pi, pi] (or degrees), atan2 of the reference's sine/cosine
    terms (geometry.py:74-124).  Shapes (*, 3) -> (*).  Always returns fp32."""
    tensors, saw, dev = _ingest([a, b, c, d])
    (pa, pb, pc, pd), lead = _points(tensors)
    n = pa.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=dev)
    with _cabi.on_device(dev):
        rc = _cabi.load().ps_geom_dihedral(pa.data_ptr(), pb.data_ptr(), pc.data_ptr(), pd.data_ptr(),
                                           n, int(bool(to_degree)), out.data_ptr(), _stream(dev))
    _cabi.check(rc, "ps_geom_dihedral")
    return _egress(out.reshape(lead), saw)


def gram_schmidt(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
    """Orthonormal frame from (c - b) and (a - b); columns e1, e2, e3 = e1 x e2
    (reference geometry.py:413-439; the cross product is always taken along the last axis,
    DESIGN.md quirk Q2).  Shapes (*, 3) -> (*, 3, 3)."""
    tensors, _, dev = _ingest([a, b, c])
    (pa, pb, pc), lead = _points(tensors)
    n = pa.shape[0]
    out = torch.empty(n, 3, 3, dtype=torch.float32, device=dev)
    with _cabi.on_device(dev):
        rc = _cabi.load().ps_geom_gram_schmidt(pa.data_ptr(), pb.data_ptr(), pc.data_ptr(), n,
                                               out.data_ptr(), _stream(dev))
    _cabi.check(rc, "ps_geom_gram_schmidt")
    return out.reshape(*lead, 3, 3)
