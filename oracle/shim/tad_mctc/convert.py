"""`tad_mctc.convert.symbol_to_number`."""
import torch

from .data.pse import S2Z


def symbol_to_number(symbols):
    return torch.tensor([S2Z[s.capitalize()] for s in symbols])
