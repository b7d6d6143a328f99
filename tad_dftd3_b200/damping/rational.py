"""`tad_dftd3.damping.rational` (reference damping/rational.py:38-68): Becke-Johnson damping."""
from . import rational_damping

__all__ = ["rational_damping"]
