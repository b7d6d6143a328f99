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

__all__ = ["rational_damping", "zero_damping", "dispersion_atm"]


def rational_damping(order: int, distances: Tensor, qq: Tensor, param: dict) -> Tensor:
    """1 / (r^n + (a1 sqrt(qq) + a2)^n)."""
    dd = {"device": distances.device, "dtype": distances.dtype}
    a1 = param.get("a1", torch.tensor(defaults.A1, **dd))
    a2 = param.get("a2", torch.tensor(defaults.A2, **dd))
    return 1.0 / (distances.pow(order) + (a1 * torch.sqrt(qq) + a2).pow(order))


def zero_damping(order: int, distances: Tensor, qq: Tensor, param: dict) -> Tensor:
    """Zero ("D3(0)"-type) damping in the signature of the reference's damping hook (disp.py:89-91,
    `damping_function(order, distances, qq, param)`):

        1 / (r^n (1 + 6 (rs_n sqrt(qq) / r)^alp_n)),   rs_6 = param["rs6"], rs_8 = param["rs8"] (default 1),
        alp_6 = param["alp"] (default 14), alp_8 = alp_6 + 2

    on the radius R0 = sqrt(qq) = sqrt(C8 / C6) -- the hook carries no pair table, so the tabulated cutoff radii
    of the original D3(0) cannot enter.  Passed as `damping_function=` it selects the kernel's second damping
    function (`d3b200_params.damping = 1`); called directly it is the plain elementwise formula, which is also
    what the unmodified reference runs when it is handed this callable (tests/golden fixtures `*_zero*`)."""
    dd = {"device": distances.device, "dtype": distances.dtype}
    rs = param.get("rs6" if order == 6 else "rs8", torch.tensor(1.0, **dd))
    alp = param.get("alp", torch.tensor(defaults.ALP, **dd)) + (order - 6)
    return 1.0 / (distances.pow(order) * (1.0 + 6.0 * (rs * torch.sqrt(qq) / distances).pow(alp)))


DAMPING_IDS = {rational_damping: 0, zero_damping: 1}   # callable -> kernel id


def dispersion_atm(numbers, positions, c6, rvdw, cutoff, s9=None, rs9=None, alp=None):
    from ..stages import dispersion_atm as _atm

    return _atm(numbers, positions, c6, rvdw, cutoff, s9, rs9, alp)
