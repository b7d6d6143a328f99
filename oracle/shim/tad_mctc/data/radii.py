"""
Radii tables of tad-mctc (`tad_mctc.data.radii`), restated.

COV_D3: covalent radii of Pyykko & Atsumi (Chem. Eur. J. 15, 2009, 188),
metals decreased by 10 % (the D3 convention), in Angstrom, converted to Bohr
and scaled by k2 = 4/3.  Values recalled from the published D3 parameter set;
H, C, S, I, Pb and Bi are validated against the reference's pinned CN values
(`test/test_model/samples.py:626-639,834-856`, see tests/test_oracle_golden.py).

VDW_PAIRWISE: the D3 pairwise r0ab table (4465 unique values) is NOT available
in this container.  `VDW_PAIRWISE()` here returns a SYNTHETIC symmetric
stand-in (flagged by `VDW_PAIRWISE_IS_SYNTHETIC`); every parity test passes
`rvdw=` explicitly so both sides see the same numbers.
"""
from __future__ import annotations

import torch

ANGSTROM_TO_BOHR = 1.0 / 0.529177210903

# fmt: off
_COV_2009 = [
    0.00,  # dummy
    0.32, 0.46,  # H, He
    1.20, 0.94, 0.77, 0.75, 0.71, 0.63, 0.64, 0.67,  # Li-Ne
    1.40, 1.25, 1.13, 1.04, 1.10, 1.02, 0.99, 0.96,  # Na-Ar
    1.76, 1.54,  # K, Ca
    1.33, 1.22, 1.21, 1.10, 1.07, 1.04, 1.00, 0.99, 1.01, 1.09,  # Sc-Zn
    1.12, 1.09, 1.15, 1.10, 1.14, 1.17,  # Ga-Kr
    1.89, 1.67,  # Rb, Sr
    1.47, 1.39, 1.32, 1.24, 1.15, 1.13, 1.13, 1.08, 1.15, 1.23,  # Y-Cd
    1.28, 1.26, 1.26, 1.23, 1.32, 1.31,  # In-Xe
    2.09, 1.76,  # Cs, Ba
    1.62, 1.47, 1.58, 1.57, 1.56, 1.55, 1.51,  # La-Eu
    1.52, 1.51, 1.50, 1.49, 1.49, 1.48, 1.53,  # Gd-Yb
    1.46, 1.37, 1.31, 1.23, 1.18, 1.16, 1.11, 1.12, 1.13, 1.32,  # Lu-Hg
    1.30, 1.30, 1.36, 1.31, 1.38, 1.42,  # Tl-Rn
    2.01, 1.81,  # Fr, Ra
    1.67, 1.58, 1.52, 1.53, 1.54, 1.55, 1.49,  # Ac-Am
    1.49, 1.51, 1.51, 1.48, 1.50, 1.56, 1.58,  # Cm-No
    1.45, 1.41, 1.34, 1.29, 1.27, 1.21, 1.16, 1.15, 1.09, 1.22,  # Lr-Cn
    1.36, 1.43, 1.46, 1.58, 1.48, 1.57,  # Nh-Og
]
# fmt: on

VDW_PAIRWISE_IS_SYNTHETIC = True


def COV_D3(dtype=None, device=None) -> torch.Tensor:
    t = torch.tensor(_COV_2009, dtype=torch.double, device=device)
    t = t * ANGSTROM_TO_BOHR * (4.0 / 3.0)
    return t.to(dtype if dtype is not None else torch.get_default_dtype())


def VDW_PAIRWISE(dtype=None, device=None) -> torch.Tensor:
    """SYNTHETIC [Zmax+1, Zmax+1] symmetric table in Bohr (see module doc)."""
    r = torch.tensor(_COV_2009, dtype=torch.double, device=device)
    r = r[:104] * ANGSTROM_TO_BOHR
    z = torch.arange(104, dtype=torch.double, device=device)
    wob = 0.05 * torch.sin(0.37 * z.unsqueeze(0) * z.unsqueeze(1) + 0.11 * (z.unsqueeze(0) + z.unsqueeze(1)))
    t = 1.15 * (r.unsqueeze(0) + r.unsqueeze(1)) + 2.1 + wob
    t[0, :] = 0.0
    t[:, 0] = 0.0
    return t.to(dtype if dtype is not None else torch.get_default_dtype())
