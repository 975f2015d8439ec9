"""protstruc_b200 — B200-native (sm_100a) implementation of protstruc's batched geometric-feature
hot path behind the reference's `StructureBatch` API.  See DESIGN.md."""
from . import geometry  # noqa: F401
from .general import ATOM, MAX_N_ATOMS_PER_RESIDUE  # noqa: F401
from .structure_batch import StructureBatch, manual_seed  # noqa: F401

__all__ = ["StructureBatch", "ATOM", "MAX_N_ATOMS_PER_RESIDUE", "geometry", "manual_seed"]
__version__ = "0.1.0"
