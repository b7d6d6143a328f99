"""
Dispersion model (mirror of the reference's `model` package):
`gaussian_weight`, `weight_references` (model/weights.py:58-176) and
`atomic_c6` (model/c6.py:42-96) as stage-level drop-ins.
"""
from __future__ import annotations

import torch
from torch import Tensor

from .. import defaults

__all__ = ["gaussian_weight", "weight_references", "atomic_c6"]


def gaussian_weight(dcn: Tensor, factor: float = defaults.WF) -> Tensor:
    """exp(-factor dcn^2); marker object for the CUDA fast path."""
    return torch.exp(-factor * dcn.pow(2))


def weight_references(numbers, cn, reference, weighting_function=gaussian_weight, **kwargs):
    from ..stages import weight_references as _w

    return _w(numbers, cn, reference, weighting_function, **kwargs)


def atomic_c6(numbers, weights, reference, chunk_size=None):
    from ..stages import atomic_c6 as _c6

    return _c6(numbers, weights, reference, chunk_size)
