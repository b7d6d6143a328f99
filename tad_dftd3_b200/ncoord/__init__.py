"""
Coordination number (mirror of the reference's `ncoord` re-exports,
`ncoord/__init__.py:23-24`; the arithmetic is tad-mctc's `cn_d3`/`exp_count`).

`exp_count` is the counting function object that selects the CUDA fast path in
`dftd3()`; called directly it is the plain elementwise formula.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import defaults

__all__ = ["exp_count", "erf_count", "cn_d3"]


def exp_count(r: Tensor, r0: Tensor, kcn: float = defaults.D3_KCN) -> Tensor:
    """Exponential counting function 1 / (1 + exp(-kcn (r0/r - 1)))."""
    return 1.0 / (1.0 + torch.exp(-kcn * (r0 / r - 1.0)))


def erf_count(r: Tensor, r0: Tensor, kcn: float = 7.5) -> Tensor:
    """Error-function counting function (1 + erf(-kcn (r/r0 - 1))) / 2 (tad-mctc `ncoord/count.py::erf_count`,
    the D4 / GFN-type coordination number).  Passed as `counting_function=` it selects the kernel's second
    counting function (`d3b200_params.counting = 1`); called directly it is the plain elementwise formula."""
    return 0.5 * (1.0 + torch.erf(-kcn * (r / r0 - 1.0)))


COUNTING_IDS = {exp_count: (0, defaults.D3_KCN), erf_count: (1, 7.5)}   # callable -> (kernel id, default kcn)


def cn_d3(numbers, positions, *, counting_function=exp_count, rcov=None, cutoff=None, **kwargs):
    """D3 coordination number on the GPU (stage-level drop-in)."""
    from ..stages import cn_d3 as _cn

    return _cn(numbers, positions, counting_function=counting_function, rcov=rcov, cutoff=cutoff, **kwargs)
