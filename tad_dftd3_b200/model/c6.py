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


class _AtomicC6:
    """Call surface of the reference's `AtomicC6_V1/V2.apply(numbers, weights, reference, chunk_size)`."""

    @staticmethod
    def apply(numbers, weights, reference, chunk_size=None):
        return atomic_c6(numbers, weights, reference, chunk_size)


AtomicC6_V1 = AtomicC6_V2 = _AtomicC6
AtomicC6Base = _C6Bilinear
