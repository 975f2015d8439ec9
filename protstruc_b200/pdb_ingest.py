"""PDB files -> padded host arrays, through the native parser.

`read_pdb_arrays(path)` parses one file with `ps_host_pdb_parse` (C++, no biotite / pandas);
`read_pdb_batch(paths)` parses several files concurrently (the call releases the GIL) and pads them like
`StructureBatch.from_pdb` of the reference does (protstruc/protstruc.py:171-187): zero coordinates, False
mask and NaN chain / residue indices beyond each structure's length.
"""
from __future__ import annotations

import ctypes
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Sequence

import numpy as np

from . import _cabi

N_SLOTS = 15
# one-letter code -> index (alphabetical one-letter order, 'X' = UNK; reference general.py:127-134)
AA_INDEX = {c: i for i, c in enumerate("ACDEFGHIKLMNPQRSTVWY")}
AA_INDEX["X"] = 20


def read_pdb_arrays(path) -> Dict[str, object]:
    lib = _cabi.load()
    text = Path(path).read_bytes()
    n = ctypes.c_int(0)
    _cabi.check(lib.ps_host_pdb_parse(text, len(text), 0, None, None, None, None, None, None, None, ctypes.byref(n)),
                "ps_host_pdb_parse")
    L = n.value
    xyz = np.empty((L, N_SLOTS, 3), dtype=np.float32)
    mask = np.empty((L, N_SLOTS), dtype=np.uint8)
    chain_idx = np.empty(L, dtype=np.int32)
    resseq = np.empty(L, dtype=np.int32)
    chain_id = np.empty(L, dtype="S1")
    icode = np.empty(L, dtype="S1")
    aa1 = np.empty(L, dtype="S1")
    if L:
        ptr = lambda a: a.ctypes.data  # noqa: E731
        _cabi.check(lib.ps_host_pdb_parse(text, len(text), L, ptr(xyz), ptr(mask), ptr(chain_idx), ptr(chain_id),
                                          ptr(resseq), ptr(icode), ptr(aa1), ctypes.byref(n)), "ps_host_pdb_parse")
    chain_chars = [c.decode() for c in chain_id]
    chain_ids: List[str] = []
    for c in chain_chars:
        if c not in chain_ids:
            chain_ids.append(c)
    letters = "".join(c.decode() for c in aa1)
    seq = {cid: "".join(ch for ch, c in zip(letters, chain_chars) if c == cid) for cid in chain_ids}
    return {"xyz": xyz, "atom_mask": mask.astype(bool), "chain_idx": chain_idx, "chain_ids": chain_ids, "seq": seq,
            "residue_number": resseq, "insertion_code": [c.decode().strip("\x00") for c in icode], "one_letter": letters}


def read_pdb_batch(paths: Sequence, max_workers: int = 8) -> Dict[str, object]:
    paths = list(paths)
    if len(paths) > 1:
        with ThreadPoolExecutor(max_workers=min(max_workers, len(paths))) as pool:
            parts = list(pool.map(read_pdb_arrays, paths))
    else:
        parts = [read_pdb_arrays(p) for p in paths]
    B, L = len(parts), max((len(p["xyz"]) for p in parts), default=0)
    xyz = np.zeros((B, L, N_SLOTS, 3), dtype=np.float32)
    mask = np.zeros((B, L, N_SLOTS), dtype=bool)
    chain_idx = np.full((B, L), np.nan, dtype=np.float32)
    residue_idx = np.full((B, L), np.nan, dtype=np.float32)
    for b, p in enumerate(parts):
        n = len(p["xyz"])
        xyz[b, :n] = p["xyz"]
        mask[b, :n] = p["atom_mask"]
        chain_idx[b, :n] = p["chain_idx"]
        residue_idx[b, :n] = np.arange(n)
    return {"xyz": xyz, "atom_mask": mask, "chain_idx": chain_idx, "residue_idx": residue_idx,
            "chain_ids": [p["chain_ids"] for p in parts], "seq": [p["seq"] for p in parts]}
