"""Helpers shared by the parity tests (inputs from golden cases, oracle calls)."""
from __future__ import annotations

import numpy as np
import torch

from tad_dftd3_b200 import synthetic as syn
from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

DTYPES = {"torch.float64": torch.double, "torch.float32": torch.float32}


def case_inputs(c, device="cpu"):
    """Tensors + parameter dict of a `reference_runs.npz` case."""
    dtype = DTYPES[str(c["dtype"])]
    numbers = torch.from_numpy(c["numbers"]).to(device)
    positions = torch.from_numpy(c["positions"]).to(device=device, dtype=dtype)
    g = torch.from_numpy(c["g"]).to(device=device, dtype=dtype)
    param = {str(k): torch.tensor(float(v), dtype=dtype, device=device)
             for k, v in zip(c["param_keys"], c["param_vals"])}
    cutoff = float(c["cutoff"])
    return dtype, numbers, positions, g, param, cutoff


def tables(dtype, device="cpu"):
    """(refcn, refc6, rcov[119], r4r2[119], rvdw[104,104]) synthetic/D3 tables."""
    return (
        REFCN(dtype, device),
        syn.synthetic_c6_table(dtype, device),
        COV_D3(dtype, device),
        R4R2(dtype, device),
        syn.synthetic_rvdw_table(dtype, device),
    )


def oracle_energy(numbers, positions, param, cutoff, dtype):
    """Oracle per-atom energies on CPU (differentiable w.r.t. positions/param)."""
    import d3_oracle as orc

    refcn, refc6, rcov, r4r2, rvdw = tables(dtype)
    return orc.dftd3(
        numbers, positions, param,
        refcn=refcn, refc6=refc6, rcov=rcov[numbers], r4r2=r4r2[numbers],
        rvdw=rvdw[numbers.unsqueeze(-1), numbers.unsqueeze(-2)], cutoff=cutoff,
    )


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-300)


def assert_close(a, b, rtol, atol, what=""):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    err = np.abs(a - b)
    bad = err > atol + rtol * np.abs(b)
    assert not bad.any(), (
        f"{what}: max abs err {err.max():.3e}, max rel err "
        f"{(err / np.maximum(np.abs(b), 1e-300))[np.abs(b) > atol].max() if (np.abs(b) > atol).any() else 0:.3e} "
        f"(rtol {rtol}, atol {atol}), {bad.sum()} of {bad.size} bad"
    )
