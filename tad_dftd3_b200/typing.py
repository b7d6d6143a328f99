"""
Type aliases under the names the reference exports from `tad_dftd3.typing`
(typing/d3.py, typing/pytorch.py; there they come from tad-mctc).
"""
from __future__ import annotations

from typing import Callable, List, Tuple, TypedDict, Union

import torch
from torch import Size, Tensor

__all__ = [
    "DD", "CountingFunction", "DampingFunction", "WeightingFunction", "Molecule", "Size", "Tensor",
    "TensorOrTensors", "get_default_device", "get_default_dtype",
]


class DD(TypedDict):
    """Keyword arguments `device` / `dtype` of tensor factories."""

    device: Union[torch.device, None]
    dtype: torch.dtype


class Molecule(TypedDict):
    """Atomic numbers and positions of one structure."""

    numbers: Tensor
    positions: Tensor


CountingFunction = Callable[..., Tensor]
DampingFunction = Callable[..., Tensor]
WeightingFunction = Callable[[Tensor], Tensor]
TensorOrTensors = Union[List[Tensor], Tuple[Tensor, ...], Tensor]


def get_default_device() -> torch.device:
    return torch.tensor(1.0).device


def get_default_dtype() -> torch.dtype:
    return torch.get_default_dtype()
