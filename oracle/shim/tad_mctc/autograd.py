"""`tad_mctc.autograd` pieces the reference and its tests use."""
from __future__ import annotations

import torch
from torch import Tensor


def is_functorch_tensor(x: Tensor) -> bool:
    f = torch._C._functorch
    return bool(f.is_batchedtensor(x) or f.is_gradtrackingtensor(x))


jacrev = torch.func.jacrev


def hess_fn_rev(f, argnums=0):
    return torch.func.jacrev(torch.func.jacrev(f, argnums=argnums), argnums=argnums)
