"""Atom-slot vocabulary of the hot path.

Mirrors the slot numbering the reference uses for the first five heavy-atom slots of every residue
(`ATOM`, reference protstruc/general.py:4-23) and its per-residue slot count
(`MAX_N_ATOMS_PER_RESIDUE`, reference protstruc/constants/__init__.py:1).  The kernels take slot
indices; names only exist on the Python side.
"""
import enum

MAX_N_ATOMS_PER_RESIDUE = 15


class ATOM(enum.IntEnum):
    """Backbone (+CB) slot indices.  Lower/mixed-case spellings are enum aliases, as in the reference."""

    N = 0
    CA = 1
    C = 2
    O = 3  # noqa: E741
    CB = 4
    # aliases (same values -> IntEnum aliases, reachable through ATOM["ca"] but not listed as members)
    n = 0
    Ca = 1
    ca = 1
    c = 2
    o = 3
    Cb = 4
    cb = 4

    @classmethod
    def is_valid(cls, value: str) -> bool:
        """Case-insensitive membership test against the canonical names (N, CA, C, O, CB)."""
        return value.upper() in cls._member_names_

    def __str__(self) -> str:
        return self.name
