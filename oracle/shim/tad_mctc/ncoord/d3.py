"""D3 coordination number of tad-mctc (`ncoord/d3.py`), restated.

Reference call site: `/root/reference/src/tad_dftd3/disp.py:150-152`.
"""
from __future__ import annotations

import torch

from .. import storch
from ..batch import real_pairs
from ..data import radii
from .count import exp_count

CUTOFF_D3 = 25.0


def cn_d3(
    numbers,
    positions,
    *,
    counting_function=exp_count,
    rcov=None,
    cutoff=None,
    **kwargs,
):
    dd = {"device": positions.device, "dtype": positions.dtype}
    if cutoff is None:
        cutoff = torch.tensor(CUTOFF_D3, **dd)
    if rcov is None:
        rcov = radii.COV_D3(**dd)[numbers]
    else:
        rcov = rcov.to(**dd)
    if numbers.shape != rcov.shape:
        raise ValueError(
            f"Shape of covalent radii {rcov.shape} is not consistent with "
            f"({numbers.shape})."
        )
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(
            f"Shape of positions ({positions.shape[:-1]}) is not consistent "
            f"with atomic numbers ({numbers.shape})."
        )

    mask = real_pairs(numbers, mask_diagonal=True)
    distances = torch.where(
        mask,
        storch.cdist(positions, positions, p=2),
        torch.tensor(torch.finfo(positions.dtype).eps, **dd),
    )
    rc = rcov.unsqueeze(-2) + rcov.unsqueeze(-1)
    cf = torch.where(
        mask * (distances <= cutoff),
        counting_function(distances, rc, **kwargs),
        torch.tensor(0.0, **dd),
    )
    return torch.sum(cf, dim=-1)
