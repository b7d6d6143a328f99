"""`tad_mctc.tools.memory` (reference call site `model/c6.py:128-131`), MB units."""
from __future__ import annotations

import torch


def memory_tensor(size, dtype) -> float:
    n = 1
    for s in size:
        n *= int(s)
    return n * torch.finfo(dtype).bits / 8 / 1024**2


def memory_device(device) -> tuple[float, float]:
    if device is not None and torch.device(device).type == "cuda":
        free, total = torch.cuda.mem_get_info(device)
    else:
        import psutil

        vm = psutil.virtual_memory()
        free, total = vm.available, vm.total
    return free / 1024**2, total / 1024**2
