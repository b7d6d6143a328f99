"""`tad_dftd3.model.weights` (reference model/weights.py:58-176)."""
from . import gaussian_weight, weight_references

__all__ = ["gaussian_weight", "weight_references"]
