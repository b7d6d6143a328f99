"""Type aliases re-exported by the reference (`typing/pytorch.py:21-31`)."""
from __future__ import annotations

from typing import Any, Callable, Dict, Protocol, TypedDict, Union

import torch
from torch import Size, Tensor

TensorOrTensors = Union[Tensor, "tuple[Tensor, ...]", "list[Tensor]"]
CountingFunction = Callable[..., Tensor]
DampingFunction = Callable[..., Tensor]


class DD(TypedDict):
    device: "torch.device | None"
    dtype: torch.dtype


class Molecule(TypedDict):
    numbers: Tensor
    positions: Tensor


def get_default_device() -> torch.device:
    return torch.tensor(1.0).device


def get_default_dtype() -> torch.dtype:
    return torch.get_default_dtype()

__all__ = [
    "Any", "Callable", "CountingFunction", "DD", "DampingFunction", "Dict", "Molecule",
    "Protocol", "Size", "Tensor", "TensorOrTensors", "TypedDict",
    "get_default_device", "get_default_dtype",
]
