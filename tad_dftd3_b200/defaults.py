"""
Default scalars of DFT-D3 (reference `defaults.py:38-73`).

`D3_CN_CUTOFF`/`D3_KCN` are what the coordination number actually uses
(tad-mctc's 25.0 / 16.0 -- the `cutoff=` argument of `dftd3` reaches only the
dispersion sums, reference `disp.py:150-152` vs `156-165`).
"""
D3_CN_CUTOFF = 25.0
D3_DISP_CUTOFF = 50.0
D3_KCN = 16.0
A1 = 0.4
A2 = 5.0
S6 = 1.0
S8 = 1.0
S9 = 1.0
RS9 = 4.0 / 3.0
ALP = 14.0
MAX_ELEMENT = 104
# weighting factor of the Gaussian reference weights (reference `model/weights.py:58`)
WF = 4.0
