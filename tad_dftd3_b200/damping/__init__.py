"""
Damping schemes (mirror of the reference's `damping` package).

`rational_damping` is the marker object for the Becke-Johnson fast path
(reference damping/rational.py:38-68); `dispersion_atm` is the stage-level
ATM entry (reference damping/atm.py:43-155).
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import defaults

__all__ = ["rational_damping", "dispersion_atm"]


def rational_damping(order: int, distances: Tensor, qq: Tensor, param: dict) -> Tensor:
    """1 / (r^n + (a1 sqrt(qq) + a2)^n)."""
    dd = {"device": distances.device, "dtype": distances.dtype}
    a1 = param.get("a1", torch.tensor(defaults.A1, **dd))
    a2 = param.get("a2", torch.tensor(defaults.A2, **dd))
    return 1.0 / (distances.pow(order) + (a1 * torch.sqrt(qq) + a2).pow(order))


def dispersion_atm(numbers, positions, c6, rvdw, cutoff, s9=None, rs9=None, alp=None):
    from ..stages import dispersion_atm as _atm

    return _atm(numbers, positions, c6, rvdw, cutoff, s9, rs9, alp)
