"""Counting functions of tad-mctc (`ncoord/count.py`)."""
import torch

KCN_D3 = 16.0


def exp_count(r, r0, kcn: float = KCN_D3):
    """1 / (1 + exp(-kcn (r0/r - 1)))."""
    return 1.0 / (1.0 + torch.exp(-kcn * (r0 / r - 1.0)))
