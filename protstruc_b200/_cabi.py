"""ctypes binding of libprotstruc_b200.so (declarations: include/protstruc_b200.h).

There is deliberately NO fallback here: if the shared library is missing or a launch fails,
an exception is raised.  PyTorch is only used by the callers for device memory and streams;
nothing torch-typed crosses this boundary (raw device pointers, ints, a cudaStream_t as void*).
"""
from __future__ import annotations

import contextlib
import ctypes
import threading
from ctypes import c_char_p, c_float, c_int, c_int64, c_uint8, c_uint64, c_void_p, POINTER
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libprotstruc_b200.so"

PS_MASK_BOOL = 0
PS_MASK_F32 = 1
PS_ANGLE_DIHEDRAL = 0
PS_ANGLE_PLANAR = 1

STATUS_NAMES = {
    0: "PS_OK",
    -1: "PS_ERR_BAD_SHAPE",
    -2: "PS_ERR_NULL_POINTER",
    -3: "PS_ERR_BAD_DTYPE",
    -4: "PS_ERR_BAD_SLOT",
    -5: "PS_ERR_MISALIGNED",
    -6: "PS_ERR_CUDA",
}

_fp = c_void_p  # device pointers travel as void*

# name -> (restype, argtypes); must list every symbol include/protstruc_b200.h declares.
SIGNATURES = {
    "ps_abi_version": (c_int, []),
    "ps_build_info": (c_char_p, []),
    "ps_last_error_string": (c_char_p, []),
    "ps_device_sm_count": (c_int, [c_int]),
    "ps_reserve_sms": (c_int, [c_int]),
    "ps_pair_dist_mask": (c_int, [_fp, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_void_p]),
    "ps_pair_dist_mask_ex": (c_int, [_fp, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_int, c_void_p]),
    "ps_pair_angles": (c_int, [_fp, c_int, c_int, c_int, POINTER(c_int), c_int, POINTER(c_int), c_int,
                               c_int, _fp, c_void_p]),
    "ps_pair_angles_ex": (c_int, [_fp, c_int, c_int, c_int, POINTER(c_int), c_int, POINTER(c_int), c_int,
                                  c_int, _fp, c_int, c_void_p]),
    "ps_trrosetta_angles": (c_int, [_fp, c_int, c_int, c_int, c_int, _fp, _fp, _fp, c_void_p]),
    "ps_trrosetta_angles_ex": (c_int, [_fp, c_int, c_int, c_int, c_int, _fp, _fp, _fp, c_int, c_void_p]),
    "ps_inter_residue_geometry": (c_int, [_fp, _fp, c_int, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int,
                                          c_void_p]),
    "ps_inter_residue_geometry_compact": (c_int, [_fp, _fp, c_int, _fp, _fp, _fp, c_int, c_int, c_int, c_void_p]),
    "ps_inter_residue_geometry_push": (c_int, [_fp, _fp, c_int, _fp, _fp, POINTER(c_void_p), c_int, c_int, c_void_p, c_int,
                                               c_int, c_int, c_int, c_void_p]),
    "ps_inter_residue_geometry_ex": (c_int, [_fp, _fp, c_int, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int,
                                             c_int, c_void_p]),
    "ps_backbone": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int, _fp, _fp, _fp,
                            c_void_p]),
    "ps_masked_stats": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, _fp, _fp, _fp, c_void_p]),
    "ps_masked_stats_ex": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, _fp, _fp, _fp, c_int, c_void_p]),
    "ps_scale_shift": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, _fp, c_void_p]),
    "ps_center_of_mass": (c_int, [_fp, c_int, c_int, c_int, c_int, _fp, c_void_p]),
    "ps_translate": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, _fp, c_void_p]),
    "ps_local_xyz": (c_int, [_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _fp, c_void_p]),
    "ps_rotate": (c_int, [_fp, _fp, c_int, c_int, c_int, c_int, _fp, c_void_p]),
    "ps_frames_to_backbone": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, _fp, _fp, c_void_p]),
    "ps_translate_bcast": (c_int, [_fp, _fp, c_int64, c_int64, c_int64, c_int, c_int, c_int, _fp, c_void_p]),
    "ps_kabsch": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, _fp, _fp, c_void_p]),
    "ps_topk_nearest_residue_mask": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_int, c_int, _fp, _fp, c_void_p]),
    "ps_host_pdb_parse": (c_int, [c_char_p, c_int64, c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp, POINTER(c_int)]),
    "ps_host_pipeline_create": (c_int, [c_int, c_int, c_int, POINTER(c_void_p)]),
    "ps_host_pipeline_destroy": (c_int, [c_void_p]),
    "ps_host_inter_residue_geometry": (c_int, [c_void_p, _fp, _fp, c_int, _fp, _fp, _fp, _fp, _fp]),
    "ps_host_pipeline_launches": (c_int64, [c_void_p]),
    "ps_pair_dist_last_plan": (c_int, [POINTER(c_int64), c_int]),
    "ps_debug_fill_pattern": (c_int, [_fp, c_int64, c_int, c_void_p]),
    "ps_diffuse": (c_int, [_fp, _fp, _fp, c_uint64, c_uint64, c_uint64, _fp, c_int, c_int64, c_void_p]),
    "ps_diffuse_steps": (c_int, [_fp, _fp, c_int, c_uint64, c_uint64, c_uint64, _fp, c_int, c_int64,
                                 c_void_p]),
    "ps_diffuse_trajectory": (c_int, [_fp, _fp, c_int, c_uint64, c_uint64, c_uint64, _fp, c_int, c_int64, c_void_p]),
    "ps_philox_normal": (c_int, [_fp, c_int64, c_uint64, c_uint64, c_uint64, c_void_p]),
    "ps_geom_dot": (c_int, [_fp, _fp, c_int64, c_int, _fp, c_void_p]),
    "ps_geom_norm": (c_int, [_fp, c_int64, c_int, _fp, c_void_p]),
    "ps_geom_unit": (c_int, [_fp, c_int64, c_int, _fp, c_void_p]),
    "ps_geom_angle": (c_int, [_fp, _fp, _fp, c_int64, c_int, _fp, c_void_p]),
    "ps_geom_dihedral": (c_int, [_fp, _fp, _fp, _fp, c_int64, c_int, _fp, c_void_p]),
    "ps_geom_gram_schmidt": (c_int, [_fp, _fp, _fp, c_int64, _fp, c_void_p]),
}


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing or a native call failed.  Never swallowed, never a fallback."""


_lock = threading.Lock()
_lib = None


def load(path: Path | None = None) -> ctypes.CDLL:
    """Loads the shared library once and attaches prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    with _lock:
        if _lib is not None and path is None:
            return _lib
        p = Path(path) if path is not None else LIB_PATH
        if not p.exists():
            raise NativeLibraryError(
                f"{p} not found: the CUDA extension is not built. Run `python -m protstruc_b200.build` "
                "(nvcc, sm_100a). protstruc_b200 has no CPU fallback."
            )
        lib = ctypes.CDLL(str(p))
        for name, (restype, argtypes) in SIGNATURES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:  # pragma: no cover - build/ABI mismatch
                raise NativeLibraryError(f"{p} does not export {name}") from exc
            fn.restype = restype
            fn.argtypes = argtypes
        if path is None:
            _lib = lib
        return lib


def last_error() -> str:
    return load().ps_last_error_string().decode("utf-8", "replace")


_NO_SWITCH = contextlib.nullcontext()


def on_device(device):
    """Context that makes `device` the current CUDA device for a launch; free when it already is (the
    usual case — `torch.cuda.device` costs ~10 us of driver calls per entry even then)."""
    import torch

    index = device.index if device.index is not None else torch.cuda.current_device()
    return _NO_SWITCH if torch.cuda.current_device() == index else torch.cuda.device(index)


def check(rc: int, what: str) -> None:
    """Maps a ps_status to an exception (RuntimeError family, message from the library)."""
    if rc == 0:
        return
    raise NativeLibraryError(f"{what} failed with {STATUS_NAMES.get(rc, rc)}: {last_error()}")


PLAN_FIELDS = ("path", "lockstep", "ctas", "tile_buffers", "active_buffers", "strip_stride", "tile_pairs", "launches",
               "sweep")


def last_pair_dist_plan() -> dict:
    """What the most recent K1 launch of this thread chose (ps_pair_dist_last_plan)."""
    out = (c_int64 * len(PLAN_FIELDS))()
    check(load().ps_pair_dist_last_plan(out, len(PLAN_FIELDS)), "ps_pair_dist_last_plan")
    return dict(zip(PLAN_FIELDS, (int(v) for v in out)))


def int_array(values):
    arr = (c_int * max(len(values), 1))(*values)
    return arr
