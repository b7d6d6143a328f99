"""
`tad_dftd3.model.c6` (reference model/c6.py:42-480).

`atomic_c6` is the differentiable stage entry.  The reference's helpers
`_atomic_c6_full` / `_atomic_c6_chunked` and its autograd classes
`AtomicC6_V1` / `AtomicC6_V2` exist there because of the `N x N x 7 x 7` gather
and its memory; here they are aliases of the one CUDA-backed implementation
(the chunk size is accepted and ignored).
"""
from __future__ import annotations

from ..stages import _C6Bilinear, atomic_c6

__all__ = ["atomic_c6"]


def _atomic_c6_full(numbers, weights, reference):
    return atomic_c6(numbers, weights, reference)


def _atomic_c6_chunked(numbers, weights, reference, chunk_size=None):
    return atomic_c6(numbers, weights, reference, chunk_size)


def memory_device(device) -> tuple:
    """(free, total) memory of `device` in MB (the reference takes this from tad-mctc's tools.memory)."""
    import torch

    device = torch.device(device)
    if device.type == "cuda":
        free, total = torch.cuda.mem_get_info(device)
        return free / 1024**2, total / 1024**2
    import os

    pages, size = os.sysconf("SC_PHYS_PAGES"), os.sysconf("SC_PAGE_SIZE")
    avail = os.sysconf("SC_AVPHYS_PAGES")
    return avail * size / 1024**2, pages * size / 1024**2


def _check_memory(numbers, weights, chunk_size=None) -> None:
    """
    Memory check of the reference (model/c6.py:102-153): `MemoryError` if the estimate exceeds the
    device's total memory, `ResourceWarning` if it exceeds the free memory.  The estimate here is
    the dense `N x N` output of `atomic_c6` -- the reference's `N x N x 7 x 7` gather (and with it
    the reason for `chunk_size`) does not exist.
    """
    n = numbers.shape[-1]
    mem = n * n * weights.element_size() / 1024**2
    free, total = memory_device(numbers.device)
    if mem > total:
        raise MemoryError(
            f"Estimated memory usage exceeds total available memory: {mem:.2f} MB > {total:.2f} MB "
            f"for the dense ({n}, {n}) C6 matrix.")
    if mem > free:
        from warnings import warn

        warn(f"Estimated memory usage appears to exceed the available memory: {mem:.2f} MB > {free:.2f} MB.",
             ResourceWarning)


class _AtomicC6:
    """Call surface of the reference's `AtomicC6_V1/V2.apply(numbers, weights, reference, chunk_size)`."""

    @staticmethod
    def apply(numbers, weights, reference, chunk_size=None):
        return atomic_c6(numbers, weights, reference, chunk_size)


AtomicC6_V1 = AtomicC6_V2 = _AtomicC6
AtomicC6Base = _C6Bilinear
