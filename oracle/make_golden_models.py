"""
Golden fixtures for the non-default model functions (`tests/golden/reference_models.npz`): the
UNMODIFIED reference driven through its own callable hooks (disp.py:89-91) with
`counting_function=erf_count` and / or `damping_function=zero_damping` -- the elementwise torch
formulae exported by `tad_dftd3_b200.ncoord` / `.damping`, which the reference evaluates on its
dense tensors and differentiates with autograd.  The CUDA kernels implement the same two functions
natively (`d3b200_params.counting / .damping`) and are compared against these outputs.

TEST INFRASTRUCTURE ONLY; needs /root/reference:   python oracle/make_golden_models.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from make_golden import d3, heavy_cluster, ref_tables, tens  # noqa: E402

from tad_dftd3_b200 import synthetic as syn  # noqa: E402
from tad_dftd3_b200.damping import rational_damping, zero_damping  # noqa: E402
from tad_dftd3_b200.ncoord import erf_count, exp_count  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_models.npz")
COUNT = {"exp": exp_count, "erf": erf_count}
DAMP = {"rational": rational_damping, "zero": zero_damping}


def run(numbers, positions, param, counting, damping, cutoff=None, hessian=False):
    dt = torch.double
    refobj, rcov_t, r4r2_t, rvdw_t = ref_tables(dt)
    pkeys = [k for k in ("s6", "s8", "a1", "a2", "rs6", "rs8") if k in param]
    par = {k: torch.tensor(v, dtype=dt, requires_grad=k in pkeys) for k, v in param.items()}
    kw = dict(ref=refobj, rcov=rcov_t[numbers], r4r2=r4r2_t[numbers],
              rvdw=rvdw_t[numbers.unsqueeze(-1), numbers.unsqueeze(-2)],
              counting_function=COUNT[counting], damping_function=DAMP[damping])
    if cutoff is not None:
        kw["cutoff"] = torch.tensor(cutoff, dtype=dt)
    pos = positions.clone().requires_grad_(True)
    e = d3.dftd3(numbers, pos, par, **kw)
    gen = torch.Generator().manual_seed(7)
    g = torch.rand(numbers.shape, dtype=dt, generator=gen) + 0.5
    grads = torch.autograd.grad((e * g).sum(), [pos] + [par[k] for k in pkeys])
    out = {"energy": e.detach().numpy(), "g": g.numpy(), "grad": grads[0].numpy(),
           "pgrad_keys": np.asarray(pkeys), "pgrad": np.asarray([float(x) for x in grads[1:]])}
    if hessian:
        pd = {k: v.detach() for k, v in par.items()}
        h = torch.func.jacrev(torch.func.jacrev(lambda p: d3.dftd3(numbers, p, pd, **kw).sum(-1)))(positions)
        out["hessian"] = h.numpy()
    return out


def main():
    store = {}

    def add(case, numbers, positions, param, counting, damping, **kw):
        out = run(numbers, positions, param, counting, damping, **kw)
        store[f"{case}/numbers"] = numbers.numpy().astype(np.int64)
        store[f"{case}/positions"] = positions.numpy()
        store[f"{case}/param_keys"] = np.asarray(list(param.keys()))
        store[f"{case}/param_vals"] = np.asarray([float(v) for v in param.values()])
        store[f"{case}/counting"] = np.asarray(counting)
        store[f"{case}/damping"] = np.asarray(damping)
        store[f"{case}/cutoff"] = np.asarray(50.0 if kw.get("cutoff") is None else kw["cutoff"])
        for k, v in out.items():
            store[f"{case}/{k}"] = v
        print(case, float(out["energy"].sum()))

    bj = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}
    zero = {"s6": 1.0, "s8": 1.2, "rs6": 1.1, "rs8": 0.95}
    mols = [tens(m) for m in syn.README_BATCH]
    z, x = syn.water_cluster(21, seed=8)
    hz, hx = heavy_cluster(natoms=30, nelem=12, seed=41, box=12.0, dmin=2.4)
    for name, (nz, nx) in {"pbh4": mols[0], "c6h5i": mols[1], "water63": (z, x), "heavy30": (hz, hx)}.items():
        small = name == "pbh4"
        add(f"{name}_erf_bj", nz, nx, bj, "erf", "rational", hessian=small)
        add(f"{name}_exp_zero", nz, nx, zero, "exp", "zero", hessian=small)
        add(f"{name}_erf_zero_atm_cut9", nz, nx, dict(zero, s9=1.0, alp=14.0), "erf", "zero", cutoff=9.0)
    np.savez_compressed(OUT, **store)
    print(OUT, os.path.getsize(OUT) // 1024, "kB")


if __name__ == "__main__":
    main()
