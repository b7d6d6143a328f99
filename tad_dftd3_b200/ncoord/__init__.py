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

__all__ = ["exp_count", "cn_d3"]


def exp_count(r: Tensor, r0: Tensor, kcn: float = defaults.D3_KCN) -> Tensor:
    """Exponential counting function 1 / (1 + exp(-kcn (r0/r - 1)))."""
    return 1.0 / (1.0 + torch.exp(-kcn * (r0 / r - 1.0)))


def cn_d3(numbers, positions, *, counting_function=exp_count, rcov=None, cutoff=None, **kwargs):
    """D3 coordination number on the GPU (stage-level drop-in)."""
    from ..stages import cn_d3 as _cn

    return _cn(numbers, positions, counting_function=counting_function, rcov=rcov, cutoff=cutoff, **kwargs)
