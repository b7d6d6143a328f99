"""
Exceptions of the package (the reference's `tad_dftd3.exception`, exception.py:22-31,
tested by test/test_disp/test_exception.py).
"""
__all__ = ["DFTD3Error"]


class DFTD3Error(Exception):
    """Base class for exceptions raised by this package."""
