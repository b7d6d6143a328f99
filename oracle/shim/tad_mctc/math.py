"""`tad_mctc.math.einsum` (reference call sites `model/c6.py:176-180,321-331`)."""
import torch


def einsum(equation, *operands, optimize=None):
    # tad-mctc forwards to opt_einsum / torch.einsum; the contraction path
    # only changes evaluation order, not the mathematical result.
    return torch.einsum(equation, *operands)
