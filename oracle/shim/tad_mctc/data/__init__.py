from . import pse, radii
from .radii import COV_D3, VDW_PAIRWISE
