"""PDB text -> padded (B, L, 15, 3) arrays — TEST INFRASTRUCTURE (builds real-data test inputs).

The reference ingests PDB files through biotite (reference protstruc/pdb.py:4,8), which is not
installed here and cannot be.  This module restates, with the standard library + numpy only, the
part of the ingest that determines the tensors the hot path consumes:

  * biotite defaults used by the reference (pdb.py:66): first MODEL only, first alternate location;
  * `tidy_structure` (pdb.py:24-40): non-standard residue names are substituted, only the 20
    canonical amino acids and atoms with a standard heavy-atom name are kept;
  * `PDB._initialize_lookup` (pdb.py:82-130): residues in file order, gaps in the residue numbering
    inside a chain are filled with UNK placeholders, chain index = order of first appearance;
  * `PDB._compute_atom_xyz` (pdb.py:132-151): NaN-initialised (L, 15, 3) coordinates, slot =
    position of the atom name in the residue type's heavy-atom list, boolean mask;
  * `StructureBatch.from_pdb` padding (protstruc.py:171-187): zero coordinates, False mask and NaN
    chain index beyond each structure's length.

Pinned against the reference's own tests: 6dc4 (chains H, L... all) -> L = 437
(tests/test_AntibodyStructureBatch.py:13 uses the antibody reader; the plain reader gives the
lengths asserted in tests/test_StructureBatch.py:127,163 for 1REX = 130 and 4EOT = 184, and
tests/test_geometry.py:214 for 15c8_HL = 229).  It is NOT part of the product.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, List, Sequence, Tuple

import numpy as np

N_SLOTS = 15

# Heavy-atom slot names per residue type (the AlphaFold atom14-style table the reference copies,
# protstruc/general.py:149-171), written compactly: side-chain names after N, CA, C, O.
_SIDE_CHAINS = {
    "ALA": "CB", "ARG": "CB CG CD NE CZ NH1 NH2", "ASN": "CB CG OD1 ND2", "ASP": "CB CG OD1 OD2",
    "CYS": "CB SG", "GLN": "CB CG CD OE1 NE2", "GLU": "CB CG CD OE1 OE2", "GLY": "",
    "HIS": "CB CG ND1 CD2 CE1 NE2", "ILE": "CB CG1 CG2 CD1", "LEU": "CB CG CD1 CD2",
    "LYS": "CB CG CD CE NZ", "MET": "CB CG SD CE", "PHE": "CB CG CD1 CD2 CE1 CE2 CZ",
    "PRO": "CB CG CD", "SER": "CB OG", "THR": "CB OG1 CG2",
    "TRP": "CB CG CD1 CD2 NE1 CE2 CE3 CZ2 CZ3 CH2", "TYR": "CB CG CD1 CD2 CE1 CE2 CZ OH",
    "VAL": "CB CG1 CG2",
}


def _slot_table() -> Dict[str, List[str]]:
    table = {}
    for res, side in _SIDE_CHAINS.items():
        names = ["N", "CA", "C", "O"] + side.split()
        if res == "GLY":
            names = ["N", "CA", "C", "O", ""]  # glycine has an empty CB slot
        names = names + [""] * (N_SLOTS - 1 - len(names)) + ["OXT"]
        assert len(names) == N_SLOTS
        table[res] = names
    return table


SLOT_NAMES = _slot_table()
HEAVY_ATOM_NAMES = {n for names in SLOT_NAMES.values() for n in names if n}

# Non-standard -> standard residue substitutions (the OpenMM/PDBFixer table the reference embeds,
# protstruc/general.py:109-124), keyed by target to keep this file short.
_SUBSTITUTIONS_BY_TARGET = {
    "ALA": "AIB ALM AYA BNN CHG CSD DAL DHA DNP FLA HAC MAA PRR TIH TPQ",
    "ARG": "ACL AGM ARM DAR HAR HMR",
    "ASN": "MEN",
    "ASP": "2AS ASA ASB ASK ASL ASQ BHD DAS DSP IAS",
    "CYS": "BCS BUC C5C C6C CAS CCS CEA CME CSO CSP CSS CSW CSX CY1 CY3 CYG CYM CYQ DCY EFC OCS PEC PR3 "
           "PYX SCH SCS SCY SHC SMC SOC",
    "GLN": "DGN",
    "GLU": "5HP CGU DGL GGL GMA PCA",
    "GLY": "GL3 GLZ GSC MPQ MSA NMC SAR",
    "HIS": "3AH DHI HIC HIP MHS NEM NEP",
    "ILE": "DIL IIL",
    "LEU": "BUG CLE DLE MLE NLE NLN NLP",
    "LYS": "ALY DLY KCX LLP LLY LYM LYZ SHR TRG",
    "MET": "CXM FME MSE OMT",
    "PHE": "DAH DPN HPQ PHI PHL",
    "PRO": "DPR HYP",
    "SER": "DSN MIS OAS SAC SEL SEP SET SVA",
    "THR": "ALO BMT DTH TPO",
    "TRP": "DTR HTR LTR TPL TRO",
    "TYR": "DTY IYR PAQ PTR STY TYB TYI TYQ TYS TYY",
    "VAL": "DIV DVA MVA",
}
SUBSTITUTIONS = {src: dst for dst, srcs in _SUBSTITUTIONS_BY_TARGET.items() for src in srcs.split()}


def _parse_atoms(text: str):
    """Yields (chain, resseq, icode, resname, atomname, altloc, xyz) of model 1, fixed PDB columns."""
    in_first_model = True
    seen_model = False
    for line in text.splitlines():
        rec = line[:6]
        if rec.startswith("MODEL"):
            if seen_model:
                in_first_model = False
            seen_model = True
            continue
        if rec.startswith("ENDMDL"):
            in_first_model = False
            continue
        if not in_first_model or rec not in ("ATOM  ", "HETATM"):
            continue
        yield (line[21], int(line[22:26]), line[26].strip(), line[17:20].strip(), line[12:16].strip(),
               line[16].strip(), (float(line[30:38]), float(line[38:46]), float(line[46:54])))


def read_structure(path) -> Tuple[np.ndarray, np.ndarray, np.ndarray, List[str]]:
    """One PDB file -> (xyz (L,15,3) fp32 NaN-filled, mask (L,15) bool, chain_idx (L,) int64, chain ids)."""
    residues = []  # [(chain, resseq, icode, resname, {atomname: xyz})] in file order
    first_altloc: Dict[Tuple, str] = {}
    for chain, resseq, icode, resname, atomname, altloc, xyz in _parse_atoms(Path(path).read_text()):
        resname = SUBSTITUTIONS.get(resname, resname)
        if resname not in SLOT_NAMES or atomname not in HEAVY_ATOM_NAMES:
            continue
        key = (chain, resseq, icode, resname)
        if altloc:  # keep only the first alternate location seen in this residue
            if first_altloc.setdefault(key, altloc) != altloc:
                continue
        if not residues or residues[-1][:4] != key:
            residues.append((chain, resseq, icode, resname, {}))
        residues[-1][4].setdefault(atomname, xyz)

    rows = []  # (chain, resname or "UNK", atoms)
    cur_chain, cur_num = None, None
    for chain, resseq, icode, resname, atoms in residues:
        if cur_chain is None or cur_chain != chain:
            cur_chain, cur_num = chain, resseq
        while cur_num + 1 < resseq:  # fill numbering gaps with placeholder residues
            rows.append((cur_chain, "UNK", {}))
            cur_num += 1
        rows.append((chain, resname, atoms))
        cur_chain, cur_num = chain, resseq

    L = len(rows)
    xyz = np.full((L, N_SLOTS, 3), np.nan, dtype=np.float32)
    mask = np.zeros((L, N_SLOTS), dtype=bool)
    chain_ids: List[str] = []
    chain_idx = np.zeros(L, dtype=np.int64)
    for r, (chain, resname, atoms) in enumerate(rows):
        if chain not in chain_ids:
            chain_ids.append(chain)
        chain_idx[r] = chain_ids.index(chain)
        if resname == "UNK":
            continue
        names = SLOT_NAMES[resname]
        for atomname, coord in atoms.items():
            if atomname in names:
                slot = names.index(atomname)
                xyz[r, slot] = coord
                mask[r, slot] = True
    return xyz, mask, chain_idx, chain_ids


def read_batch(paths: Sequence) -> Dict[str, np.ndarray]:
    """Several PDB files -> the padded arrays `StructureBatch.from_pdb` would hold."""
    parts = [read_structure(p) for p in paths]
    B, L = len(parts), max(len(p[0]) for p in parts)
    xyz = np.zeros((B, L, N_SLOTS, 3), dtype=np.float32)
    mask = np.zeros((B, L, N_SLOTS), dtype=bool)
    chain_idx = np.full((B, L), np.nan, dtype=np.float32)
    for b, (x, m, c, _) in enumerate(parts):
        xyz[b, : len(x)] = x
        mask[b, : len(m)] = m
        chain_idx[b, : len(c)] = c
    return {"xyz": xyz, "atom_mask": mask, "chain_idx": chain_idx, "chain_ids": [p[3] for p in parts]}
