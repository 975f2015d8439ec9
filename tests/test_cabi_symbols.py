"""CPU suite, part 2: the C-ABI library builds, loads, and exports exactly what include/*.h declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from protstruc_b200 import _cabi, build

REPO = Path(__file__).resolve().parent.parent
HEADER = REPO / "include" / "protstruc_b200.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(ps_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for required in ("ps_pair_dist_mask", "ps_pair_angles", "ps_trrosetta_angles", "ps_inter_residue_geometry",
                     "ps_backbone", "ps_masked_stats", "ps_center_of_mass", "ps_diffuse", "ps_diffuse_steps",
                     "ps_last_error_string"):
        assert required in names


def test_library_builds_for_sm_100a_and_exports_every_declared_symbol(native_lib):
    assert build.LIB_PATH.exists()
    names = declared_functions()
    for name in names:
        assert hasattr(native_lib, name), f"{name} is declared in the header but not exported"
    assert sorted(_cabi.SIGNATURES) == names, "ctypes prototypes and header are out of sync"
    out = subprocess.run(["nm", "-D", "--defined-only", str(build.LIB_PATH)], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (ps_[a-z0-9_]+)", out))
    assert exported == set(names), f"exported {sorted(exported ^ set(names))} differ from the header"


def test_library_identifies_itself_without_a_gpu(native_lib):
    assert native_lib.ps_abi_version() == 1
    info = native_lib.ps_build_info().decode()
    assert "sm_100a" in info and "nvcc" in info
    assert isinstance(native_lib.ps_last_error_string(), bytes)


def test_library_contains_sm_100a_code_with_bulk_copy_and_packed_fp32():
    """The distance kernel must be real Blackwell code: TMA bulk stores (UBLKCP) and f32x2 math."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    build.build()
    res = subprocess.run([cuobjdump, "-lelf", str(build.LIB_PATH)], capture_output=True, text=True)
    assert "sm_100a" in res.stdout
    sass = subprocess.run([cuobjdump, "-sass", str(build.LIB_DIR / "obj" / "pair_dist.o")], capture_output=True,
                          text=True).stdout
    assert "UBLKCP" in sass, "no TMA bulk store in the distance kernel"
    assert "FFMA2" in sass and "FADD2" in sass, "no packed f32x2 arithmetic in the distance kernel"


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(_cabi.NativeLibraryError, match="no CPU fallback"):
        _cabi.load(tmp_path / "libprotstruc_b200.so")


def test_status_codes_are_mapped_to_exceptions(native_lib):
    with pytest.raises(_cabi.NativeLibraryError, match="PS_ERR_BAD_SHAPE"):
        _cabi.check(-1, "unit test")
    _cabi.check(0, "unit test")


def test_argument_validation_happens_before_any_cuda_call(native_lib):
    """Every entry point validates shapes / pointers / codes first and reports through the status code and the
    thread-local message — no launch, so this runs without a GPU (compute calls would need one)."""
    lib = native_lib
    null = None
    fake = 4096  # a non-NULL pointer value that is never dereferenced: validation fails first

    def message():
        return lib.ps_last_error_string().decode()

    assert lib.ps_pair_dist_mask(fake, fake, 0, fake, fake, 0, 8, 15, null) == -1 and "must be > 0" in message()
    assert lib.ps_pair_dist_mask(null, fake, 0, fake, fake, 1, 8, 15, null) == -2 and "xyz is NULL" in message()
    assert lib.ps_pair_dist_mask(fake, fake, 7, fake, fake, 1, 8, 15, null) == -3 and "mask_dtype" in message()
    assert lib.ps_pair_dist_mask(fake, null, 0, fake, fake, 1, 8, 15, null) == -2  # mask in without mask out
    assert lib.ps_pair_dist_mask(fake, null, 0, null, null, 1, 8, 15, null) == -2 and "nothing to compute" in message()
    # the trRosetta triple needs the CB slot
    assert lib.ps_inter_residue_geometry(fake, fake, 0, fake, fake, fake, fake, fake, 1, 8, 4, null) == -1
    assert "CB slot" in message()
    assert lib.ps_center_of_mass(fake, 1, 8, 15, 15, fake, null) == -4 and "slot 15" in message()
    assert lib.ps_translate(fake, fake, 3, 2, 8, 15, fake, null) == -1 and "t_rows" in message()
    assert lib.ps_rotate(fake, fake, 1, 1, 8, 15, fake, null) == -5  # in-place rotation is refused
    # (elem_offset may be ANY value since round 2: the Philox stream is addressed per element)
    assert lib.ps_diffuse(fake, fake, null, 1, 0, 2, fake, 0, 360, null) == -1 and "must be > 0" in message()
    slots = _cabi.int_array([0, 1, 2])
    assert lib.ps_pair_angles(fake, 1, 8, 15, slots, 3, slots, 3, 0, fake, null) == -1 and "needs 4 atoms" in message()
    bad = _cabi.int_array([0, 1, 99])
    assert lib.ps_pair_angles(fake, 1, 8, 15, bad, 3, slots, 1, 0, fake, null) == -4 and "slot 99" in message()
