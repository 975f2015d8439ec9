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
