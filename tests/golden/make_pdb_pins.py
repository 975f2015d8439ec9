#!/usr/bin/env python
"""Independent pins for the PDB ingest (row f3): per reference PDB file, how many residues and heavy atoms survive the
reference's documented filters, and the sum of their coordinates — computed from the PDB TEXT by a third code path
(regular expressions over the fixed columns, counters, sets), sharing nothing with the native parser
(protstruc_b200/csrc/pdb_ingest.cu) or the Python restatement (oracle/pdb_fixture_reader.py).

biotite is not installable in the build container, so the reference's real parser (protstruc/pdb.py:55-151) has never
run here; what pins the ingest is (a) the four residue counts the reference's own tests assert, (b) these text-level
counts, (c) agreement of two independently written parsers on every file.

    python tests/golden/make_pdb_pins.py          # needs /root/reference; writes tests/golden/pdb_pins.json

Filters (reference protstruc/pdb.py:24-40, 66, 132-151; protstruc/general.py:109-171): first MODEL, first alternate
location of an atom, residue name substituted by the non-standard table then restricted to the 20 canonical amino acids,
atom name in the union of heavy-atom names AND in the residue type's own slot list, hydrogens and hetero groups that
do not map to an amino acid dropped.
"""
import json
import re
import sys
from collections import OrderedDict
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "pdb_pins.json"

SIDE = """ALA CB|ARG CB CG CD NE CZ NH1 NH2|ASN CB CG OD1 ND2|ASP CB CG OD1 OD2|CYS CB SG|GLN CB CG CD OE1 NE2|GLU CB CG CD OE1 OE2|
GLY|HIS CB CG ND1 CD2 CE1 NE2|ILE CB CG1 CG2 CD1|LEU CB CG CD1 CD2|LYS CB CG CD CE NZ|MET CB CG SD CE|
PHE CB CG CD1 CD2 CE1 CE2 CZ|PRO CB CG CD|SER CB OG|THR CB OG1 CG2|TRP CB CG CD1 CD2 NE1 CE2 CE3 CZ2 CZ3 CH2|
TYR CB CG CD1 CD2 CE1 CE2 CZ OH|VAL CB CG1 CG2""".replace("\n", "")
ALLOWED = {}
for item in SIDE.split("|"):
    parts = item.split()
    ALLOWED[parts[0]] = {"N", "CA", "C", "O", "OXT", *parts[1:]}

RECORD = re.compile(r"^(ATOM  |HETATM)")


def substitutions():
    """The non-standard residue table, read from the reference's own source text (not imported: biotite)."""
    text = (REF / "protstruc" / "general.py").read_text()
    block = text[text.index("non_standard_residue_substitutions = {"):]
    block = block[:block.index("}")]
    return dict(re.findall(r"'(\w{3})': '(\w{3})'", block))


def pins_for(path: Path, subst: dict) -> dict:
    seen_atoms = set()
    residues = OrderedDict()
    n_atoms, total = 0, 0.0
    for line in path.read_text().splitlines():
        if line.startswith("ENDMDL"):
            break  # first model only
        if not RECORD.match(line):
            continue
        name, altloc, resname = line[12:16].strip(), line[16], line[17:20].strip()
        chain, resnum, icode = line[21], int(line[22:26]), line[26]
        resname = subst.get(resname, resname)
        if resname not in ALLOWED or name not in ALLOWED[resname]:
            continue
        key = (chain, resnum, icode, name)
        if key in seen_atoms:
            continue  # a later alternate location of the same atom
        seen_atoms.add(key)
        residues.setdefault((chain, resnum, icode), resname)
        n_atoms += 1
        total += float(line[30:38]) + float(line[38:46]) + float(line[46:54])
    # residues after gap filling: inside a chain, a jump of the residue number by more than one inserts placeholders
    filled, prev_chain, prev_num = 0, None, None
    for (chain, resnum, _icode) in residues:
        if prev_chain == chain and resnum > prev_num + 1:
            filled += resnum - prev_num - 1
        prev_chain, prev_num = chain, resnum
    return {"residues_with_atoms": len(residues), "residues_with_gap_fill": len(residues) + filled, "heavy_atoms": n_atoms,
            "coordinate_sum": round(total, 3), "chains": "".join(OrderedDict.fromkeys(c for (c, _, _) in residues))}


def main():
    if not REF.exists():
        sys.exit("needs /root/reference")
    subst = substitutions()
    files = sorted(list((REF / "tests").glob("*.pdb")) + list((REF / "docs" / "tutorials").glob("*.pdb")))
    pins = {f"{f.parent.name}/{f.name}": pins_for(f, subst) for f in files}
    OUT.write_text(json.dumps(pins, indent=1, sort_keys=True) + "\n")
    for k, v in pins.items():
        print(k, v)


if __name__ == "__main__":
    main()
