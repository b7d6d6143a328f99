"""
Atomic data of the D3 model (layer L1 of the reference: `data/r4r2.py`,
`reference.py:32-161`, and the radii that the reference takes from tad-mctc).

* `R4R2(dtype, device)`            -- sqrt(0.5 <r4>/<r2> sqrt(Z)), `[119]`
  (reference `data/r4r2.py:31-85`; evaluated in the requested dtype, as there).
* `COV_D3(dtype, device)`          -- 4/3-scaled Pyykko-Atsumi covalent radii in
  Bohr, `[119]` (tad-mctc `data.radii.COV_D3`, reference call `disp.py:142`).
* `VDW_PAIRWISE(dtype, device)`    -- D3 pairwise r0ab table `[104, 104]`
  (tad-mctc `data.radii.VDW_PAIRWISE`, reference call `disp.py:144`).  The table
  is not available offline: it is read from `$TAD_DFTD3_B200_R0AB` / `r0ab.pt`
  next to this file, or from an installed tad-mctc; otherwise a clear error
  tells the caller to pass `rvdw=`.
* `REFCN(dtype, device)`           -- reference coordination numbers `[104, 7]`,
  -1 = reference absent (reference `reference.py:32-161`).
"""
from __future__ import annotations

import os

import numpy as np
import torch

__all__ = ["R4R2", "COV_D3", "VDW_PAIRWISE", "REFCN", "COV_2009_ANGSTROM", "ANGSTROM_TO_BOHR",
           "set_vdw_pairwise"]

ANGSTROM_TO_BOHR = 1.0 / 0.529177210903

_HERE = os.path.dirname(os.path.abspath(__file__))
_TABLES = np.load(os.path.join(_HERE, "d3_tables.npz"))

# Covalent radii (Angstrom) of Pyykko & Atsumi, Chem. Eur. J. 15 (2009) 188,
# metals reduced by 10 % as in all D3 implementations; index = atomic number.
# fmt: off
COV_2009_ANGSTROM = [
    0.00,
    0.32, 0.46,
    1.20, 0.94, 0.77, 0.75, 0.71, 0.63, 0.64, 0.67,
    1.40, 1.25, 1.13, 1.04, 1.10, 1.02, 0.99, 0.96,
    1.76, 1.54,
    1.33, 1.22, 1.21, 1.10, 1.07, 1.04, 1.00, 0.99, 1.01, 1.09,
    1.12, 1.09, 1.15, 1.10, 1.14, 1.17,
    1.89, 1.67,
    1.47, 1.39, 1.32, 1.24, 1.15, 1.13, 1.13, 1.08, 1.15, 1.23,
    1.28, 1.26, 1.26, 1.23, 1.32, 1.31,
    2.09, 1.76,
    1.62, 1.47, 1.58, 1.57, 1.56, 1.55, 1.51,
    1.52, 1.51, 1.50, 1.49, 1.49, 1.48, 1.53,
    1.46, 1.37, 1.31, 1.23, 1.18, 1.16, 1.11, 1.12, 1.13, 1.32,
    1.30, 1.30, 1.36, 1.31, 1.38, 1.42,
    2.01, 1.81,
    1.67, 1.58, 1.52, 1.53, 1.54, 1.55, 1.49,
    1.49, 1.51, 1.51, 1.48, 1.50, 1.56, 1.58,
    1.45, 1.41, 1.34, 1.29, 1.27, 1.21, 1.16, 1.15, 1.09, 1.22,
    1.36, 1.43, 1.46, 1.58, 1.48, 1.57,
]
# fmt: on


def _dt(dtype):
    return dtype if dtype is not None else torch.get_default_dtype()


def R4R2(dtype: torch.dtype | None = None, device: torch.device | None = None) -> torch.Tensor:
    raw = torch.tensor(_TABLES["r4_over_r2"], device=device, dtype=_dt(dtype))
    sqrtz = torch.sqrt(torch.arange(raw.shape[0], device=device, dtype=_dt(dtype)))
    return torch.sqrt(0.5 * (raw * sqrtz))


def COV_D3(dtype: torch.dtype | None = None, device: torch.device | None = None) -> torch.Tensor:
    t = torch.tensor(COV_2009_ANGSTROM, dtype=torch.double, device=device)
    return (t * ANGSTROM_TO_BOHR * (4.0 / 3.0)).to(_dt(dtype))


def REFCN(dtype: torch.dtype | None = None, device: torch.device | None = None) -> torch.Tensor:
    return torch.tensor(_TABLES["refcn"], device=device, dtype=_dt(dtype))


_registered_vdw = None


def set_vdw_pairwise(table: torch.Tensor | None) -> None:
    """Register the `[>=104, >=104]` D3 r0ab table that `VDW_PAIRWISE()` returns
    (the table is not bundled; see the module docstring)."""
    global _registered_vdw
    if table is not None and (table.dim() != 2 or table.shape[0] < 104 or table.shape[0] != table.shape[1]):
        raise ValueError("expected a square [>=104, >=104] table")
    _registered_vdw = None if table is None else table.detach().clone()
    try:  # drop cached device copies
        from .. import disp

        for k in [k for k in disp._table_cache if k[0] == "VDW_PAIRWISE"]:
            del disp._table_cache[k]
    except ImportError:  # pragma: no cover
        pass


def VDW_PAIRWISE(dtype: torch.dtype | None = None, device: torch.device | None = None) -> torch.Tensor:
    if _registered_vdw is not None:
        return _registered_vdw[:104, :104].to(device=device, dtype=_dt(dtype)).contiguous()
    path = os.environ.get("TAD_DFTD3_B200_R0AB", os.path.join(_HERE, "r0ab.pt"))
    if os.path.isfile(path):
        return torch.load(path, map_location=device, weights_only=True).to(_dt(dtype))
    try:  # a real tad-mctc install carries the table
        from tad_mctc.data import radii  # type: ignore

        if not getattr(radii, "VDW_PAIRWISE_IS_SYNTHETIC", False):
            return radii.VDW_PAIRWISE(dtype=_dt(dtype), device=device)
    except ImportError:
        pass
    raise FileNotFoundError(
        "The D3 pairwise van-der-Waals radii (r0ab, needed only for the ATM "
        "three-body term) are not bundled: pass `rvdw=` explicitly, install "
        "tad-mctc, or point $TAD_DFTD3_B200_R0AB to a [104,104] tensor file."
    )
