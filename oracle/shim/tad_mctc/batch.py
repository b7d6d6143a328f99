"""Batch helpers of tad-mctc (`batch.pack`, `real_pairs`, `real_triples`)."""
from __future__ import annotations

import torch
from torch import Tensor


def real_atoms(numbers: Tensor) -> Tensor:
    return numbers != 0


def real_pairs(numbers: Tensor, mask_diagonal: bool = True) -> Tensor:
    real = numbers != 0
    mask = real.unsqueeze(-2) * real.unsqueeze(-1)
    if mask_diagonal:
        mask = mask * ~torch.diag_embed(torch.ones_like(real))
    return mask


def real_triples(
    numbers: Tensor, mask_diagonal: bool = True, mask_self: bool = True
) -> Tensor:
    real = real_pairs(numbers, mask_diagonal=False)
    mask = real.unsqueeze(-3) * real.unsqueeze(-2) * real.unsqueeze(-1)
    if mask_diagonal:
        mask = mask * ~torch.diag_embed(torch.ones_like(real))
    if mask_self:
        n = numbers.shape[-1]
        eye = torch.eye(n, dtype=torch.bool, device=numbers.device)
        mask = mask * ~eye.unsqueeze(-1)  # i == j
        mask = mask * ~eye.unsqueeze(-2)  # i == k
        mask = mask * ~eye.unsqueeze(-3)  # j == k
    return mask


def pack(tensors, axis: int = 0, value=0, size=None) -> Tensor:
    """Zero-pad a sequence of tensors to a common shape and stack them."""
    tensors = list(tensors)
    if size is None:
        size = torch.tensor([list(t.shape) for t in tensors]).max(0).values.tolist()
    padded = torch.full(
        (len(tensors), *size), value, dtype=tensors[0].dtype, device=tensors[0].device
    )
    for n, t in enumerate(tensors):
        padded[(n, *[slice(0, s) for s in t.shape])] = t
    if axis != 0:
        padded = padded.movedim(0, axis)
    return padded
