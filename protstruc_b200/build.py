"""Builds libprotstruc_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m protstruc_b200.build [--force] [--verbose]

The library is written next to this file (protstruc_b200/lib/) so that it travels with the
repository snapshot to the GPU box; it is git-ignored (*.so).
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_DIR = PKG_DIR / "lib"
LIB_PATH = LIB_DIR / "libprotstruc_b200.so"
STAMP_PATH = LIB_DIR / "libprotstruc_b200.stamp"
INCLUDE_DIR = PKG_DIR.parent / "include"

SOURCES = ["cabi.cu", "pair_dist.cu", "pair_sweep.cu", "pair_angles.cu", "backbone.cu", "stats.cu", "diffuse.cu", "align.cu", "pdb_ingest.cu", "host_pipeline.cu"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    # IEEE division / sqrt and no flush-to-zero: NaN / inf semantics must match the reference.
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set $NVCC or put it on PATH)")


def source_digest() -> str:
    h = hashlib.sha256()
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(INCLUDE_DIR.glob("*.h"))
    for f in files:
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    return LIB_PATH.exists() and STAMP_PATH.exists() and STAMP_PATH.read_text().strip() == source_digest()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and is_current():
        return LIB_PATH
    nvcc = find_nvcc()
    LIB_DIR.mkdir(exist_ok=True)
    obj_dir = LIB_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    objs = []
    for src in SOURCES:
        obj = obj_dir / (Path(src).stem + ".o")
        objs.append(str(obj))
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE_DIR), "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log_lines = []
    failed = False
    for src, cmd, proc in procs:
        out, _ = proc.communicate()
        log_lines.append(f"$ {' '.join(cmd)}\n{out}")
        if proc.returncode != 0:
            failed = True
            sys.stderr.write(f"nvcc failed for {src}:\n{out}\n")
    (LIB_DIR / "build.log").write_text("\n".join(log_lines))
    if failed:
        raise RuntimeError("nvcc compilation failed (see protstruc_b200/lib/build.log)")
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB_PATH), *objs]
    res = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout)
        raise RuntimeError("nvcc link failed")
    STAMP_PATH.write_text(source_digest())
    if verbose:
        sys.stdout.write("\n".join(log_lines))
    return LIB_PATH


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    path = build(force=args.force, verbose=args.verbose)
    print(path)


if __name__ == "__main__":
    main()
