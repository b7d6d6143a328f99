"""
One-off extraction of the D3 model constants that the reference keeps inline
in Python source (`reference.py:32-161` reference CNs, `data/r4r2.py:55-85`
<r4>/<r2> expectation values) into `tad_dftd3_b200/data/d3_tables.npz`.

The numbers are published D3 model parameters (Grimme et al., JCP 132, 154104);
they are model DATA that "stays as it is" (BASELINE.json north_star), stored
here as arrays so that no reference source text is carried over.

Run in the build container only:  python tools/extract_tables.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from refimport import import_reference  # noqa: E402

import_reference()
from tad_dftd3.data import R4R2  # noqa: E402
from tad_dftd3.reference import _load_cn  # noqa: E402

import ast  # noqa: E402
import inspect  # noqa: E402

refcn = _load_cn(torch.double).numpy()

# the raw <r4>/<r2> list is a local of R4R2(); read it from the function's AST
tree = ast.parse(inspect.getsource(R4R2))
raw = None
for node in ast.walk(tree):
    if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") == "_r4_over_r2":
        raw = np.asarray(ast.literal_eval(node.value), dtype=np.float64)
assert raw is not None
z = np.arange(len(raw), dtype=np.float64)
# round-trip check of the defining formula sqrt(0.5 * raw * sqrt(Z))
assert np.allclose(np.sqrt(0.5 * (raw * np.sqrt(z))), R4R2(torch.double).numpy(), rtol=4e-16, atol=0)
out = os.path.join(ROOT, "tad_dftd3_b200", "data", "d3_tables.npz")
np.savez(out, refcn=refcn, r4_over_r2=raw)
print("wrote", out, refcn.shape, raw.shape)
