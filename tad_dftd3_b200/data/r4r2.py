"""`tad_dftd3.data.r4r2` (reference data/r4r2.py:31-85): sqrt(0.5 r4/r2 sqrt(Z)) per element."""
from . import R4R2

__all__ = ["R4R2"]
