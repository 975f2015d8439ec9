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
    # One parse in the common case: PDB lines are 81 bytes and an all-atom file has ~8 ATOM lines per residue, so
    # one row per four lines is a generous first guess; files with long numbering gaps (UNK placeholder rows) or
    # CA-only traces trigger the second call.
    capacity = len(text) // (81 * 4) + 64
    while True:
        xyz = np.empty((capacity, N_SLOTS, 3), dtype=np.float32)
        mask = np.empty((capacity, N_SLOTS), dtype=np.uint8)
        chain_idx = np.empty(capacity, dtype=np.int32)
        resseq = np.empty(capacity, dtype=np.int32)
        chain_id = np.empty(capacity, dtype="S1")
        icode = np.empty(capacity, dtype="S1")
        aa1 = np.empty(capacity, dtype="S1")
        ptr = lambda a: a.ctypes.data  # noqa: E731
        rc = lib.ps_host_pdb_parse(text, len(text), capacity, ptr(xyz), ptr(mask), ptr(chain_idx), ptr(chain_id),
                                   ptr(resseq), ptr(icode), ptr(aa1), ctypes.byref(n))
        if n.value <= capacity:
            _cabi.check(rc, "ps_host_pdb_parse")
            break
        capacity = n.value  # the row count is reported even when the rows did not fit
    L = n.value
    xyz, mask, chain_idx, resseq = xyz[:L], mask[:L], chain_idx[:L], resseq[:L]
    chain_id, icode, aa1 = chain_id[:L], icode[:L], aa1[:L]
    # per-residue characters as flat strings (one decode each instead of a Python loop over residues)
    chain_chars = chain_id.tobytes().decode("latin-1")
    letters = aa1.tobytes().decode("latin-1")
    chain_ids: List[str] = list(dict.fromkeys(chain_chars))  # order of first appearance
    if len(chain_ids) == 1:
        seq = {chain_ids[0]: letters}
    else:
        codes = np.frombuffer(chain_id.tobytes(), dtype=np.uint8)
        letter_codes = np.frombuffer(aa1.tobytes(), dtype=np.uint8)
        seq = {cid: letter_codes[codes == ord(cid)].tobytes().decode("latin-1") for cid in chain_ids}
    return {"xyz": xyz, "atom_mask": mask.view(np.bool_), "chain_idx": chain_idx, "chain_ids": chain_ids, "seq": seq,
            "residue_number": resseq, "insertion_code": ["" if ch == "\x00" else ch for ch in icode.tobytes().decode("latin-1")],
            "one_letter": letters}


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
