"""
"Safe" torch ops of tad-mctc (`storch`), restated.

Call sites in the reference: `disp.py:279`, `damping/atm.py:97,111`,
`model/weights.py:147`.
"""
from __future__ import annotations

import torch
from torch import Tensor


def sqrt(x: Tensor, *, eps: Tensor | float | None = None) -> Tensor:
    """sqrt(clamp(x, min=eps)), eps = finfo(dtype).eps."""
    if eps is None:
        eps = torch.finfo(x.dtype).eps
    return torch.sqrt(torch.clamp(x, min=eps))


def divide(x: Tensor, y: Tensor, *, eps: Tensor | float | None = None) -> Tensor:
    """x / where(y == 0, eps, y)."""
    if eps is None:
        eps = torch.finfo(x.dtype).eps
    if not isinstance(eps, Tensor):
        eps = torch.tensor(eps, device=y.device, dtype=y.dtype)
    return x / torch.where(y == 0, eps.to(y.dtype), y)


def cdist(x: Tensor, y: Tensor | None = None, p: int = 2) -> Tensor:
    """
    Pairwise distances.  For p == 2 tad-mctc uses the quadratic expansion
    |x|^2 + |y|^2 - 2 x.y followed by sqrt(clamp(., min=eps)).
    """
    if y is None:
        y = x
    eps = torch.finfo(x.dtype).eps
    if p == 2:
        xnorm = torch.einsum("...ij,...ij->...i", x, x)
        ynorm = torch.einsum("...ij,...ij->...i", y, y)
        n = xnorm.unsqueeze(-1) + ynorm.unsqueeze(-2)
        prod = torch.einsum("...ik,...jk->...ij", x, y)
        return torch.sqrt(torch.clamp(n - 2.0 * prod, min=eps))
    diff = torch.abs(x.unsqueeze(-2) - y.unsqueeze(-3))
    return torch.pow(torch.clamp(torch.sum(diff**p, -1), min=eps), 1.0 / p)
