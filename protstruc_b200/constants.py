"""Ideal backbone geometry used by the frame -> coordinate constructors.

Values are the reference's ideal bond lengths / angles (reference protstruc/constants/ideal.py:1-50);
only the ones the hot path and its neighbours need are listed.
"""
import math

import torch

NA = 1.458   # N - CA bond length (Angstrom)
AC = 1.523   # CA - C bond length
AB = 1.522   # CA - CB bond length
NAC = 1.937  # N - CA - C angle (rad)

# virtual-CB coefficients (reference protstruc/geometry.py:221)
CB_COEFF_A, CB_COEFF_B, CB_COEFF_C = -0.58273431, 0.56802827, -0.54067466


def ideal_backbone(include_cb: bool = False) -> torch.Tensor:
    """(3, 3) or (4, 3) fp32 table: N, CA, C (and CB) of the ideal residue with CA at the origin and CA->C
    along x (reference protstruc/geometry.py:206-224).  A 12-number constant table evaluated on the host."""
    ca = torch.zeros(3)
    c = torch.tensor([AC, 0.0, 0.0])
    n = torch.tensor([NA * math.cos(NAC), NA * math.sin(NAC), 0.0])
    if not include_cb:
        return torch.stack([n, ca, c])
    b, cc = ca - n, c - ca
    a = torch.linalg.cross(b, cc)
    cb = CB_COEFF_A * a + CB_COEFF_B * b + CB_COEFF_C * cc + ca
    return torch.stack([n, ca, c, cb])
