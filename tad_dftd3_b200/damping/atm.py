"""`tad_dftd3.damping.atm` (reference damping/atm.py:43-155): the ATM stage entry."""
from . import dispersion_atm

__all__ = ["dispersion_atm"]
