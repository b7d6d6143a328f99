"""
Pin the blocked fp64 C oracle (`oracle/d3_blocked.c`, the checker of the full-size
GPU tests) before anything trusts it:

* against energies and autograd gradients of the UNMODIFIED reference at the largest
  sizes it can hold (`tests/golden/reference_large.npz`, made by
  `oracle/make_golden_large.py`): 2001-atom water cluster and a 608-atom / 12-element
  cluster with ghost atoms (two-body + CN, default cutoffs), 300- and 152-atom ATM
  cases with cutoffs that cut most triples partly (ordered rule, atm.py:147-149);
* against every single-structure case of `reference_runs*.npz` (small molecules,
  cutoffs 6..20 Bohr, underflow fallback, close contacts);
* against the dense torch oracle on a seeded cluster (independent derivation of the
  derivatives: torch autograd there, analytic expressions here).

fp64 tolerances of BASELINE.json: 1e-10 relative / 1e-12 Hartree absolute.
"""
import numpy as np
import pytest
import torch

from d3_blocked import dftd3_blocked
from helpers import assert_close, case_inputs, oracle_energy, tables

RTOL, ATOL = 1e-10, 1e-12


@pytest.fixture(scope="module")
def large():
    from conftest import Golden

    return Golden("reference_large.npz")


def _blocked(c, numbers, positions, g, cutoff):
    refcn, refc6, rcov, r4r2, rvdw = tables(torch.double)
    param = {str(k): float(v) for k, v in zip(c["param_keys"], c["param_vals"])}
    return dftd3_blocked(numbers, positions, param, refcn=refcn, refc6=refc6, rcov=rcov[numbers], r4r2=r4r2[numbers],
                         rvdw_table=rvdw, g=g, cutoff=cutoff)


@pytest.mark.parametrize("name", ["water2001", "heavy600", "water300_atm", "heavy150_atm"])
def test_blocked_oracle_vs_reference_large(large, name):
    c = large.case(name)
    numbers, positions = torch.from_numpy(c["numbers"]), torch.from_numpy(c["positions"])
    out = _blocked(c, numbers, positions, c["g"], float(c["cutoff"]))
    assert_close(out["energy"], c["energy"], RTOL, ATOL, name + " energy")
    assert_close(out["grad"], c["grad"], RTOL, ATOL, name + " grad")
    ghosts = c["numbers"] == 0
    assert (out["energy"][ghosts] == 0).all() and (out["grad"][ghosts] == 0).all()


def test_blocked_oracle_vs_reference_small(runs):
    """Every fp64 single-structure case of the small fixtures (energies and VJPs of the unmodified reference)."""
    checked = 0
    for name in runs.cases:
        c = runs.case(name)
        if "energy" not in c or "grad" not in c or str(c["dtype"]) != "torch.float64" or c["numbers"].ndim != 1:
            continue
        if "rcov" in c or "r4r2" in c:        # cases with caller-supplied per-atom arrays are covered elsewhere
            continue
        _, numbers, positions, g, _, cutoff = case_inputs(c)
        out = _blocked(c, numbers, positions, g.numpy(), cutoff)
        assert_close(out["energy"], c["energy"], RTOL, ATOL, name + " energy")
        assert_close(out["grad"], c["grad"], RTOL, ATOL, name + " grad")
        checked += 1
    assert checked >= 10, checked


@pytest.mark.parametrize("s9,cutoff", [(0.0, 50.0), (1.0, 50.0), (1.0, 10.0)])
def test_blocked_oracle_vs_dense_oracle(s9, cutoff):
    from tad_dftd3_b200 import synthetic as syn

    dt = torch.double
    numbers, positions = syn.water_cluster(70, seed=9)
    gen = torch.Generator().manual_seed(5)
    g = torch.rand(numbers.shape, dtype=dt, generator=gen) + 0.5
    par = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}
    if s9:
        par["s9"] = s9
    x = positions.clone().requires_grad_(True)
    e = oracle_energy(numbers, x, {k: torch.tensor(v, dtype=dt) for k, v in par.items()}, cutoff, dt)
    (gr,) = torch.autograd.grad((e * g).sum(), x)
    refcn, refc6, rcov, r4r2, rvdw = tables(dt)
    out = dftd3_blocked(numbers, positions, par, refcn=refcn, refc6=refc6, rcov=rcov[numbers], r4r2=r4r2[numbers],
                        rvdw_table=rvdw, g=g, cutoff=cutoff)
    assert_close(out["energy"], e.detach(), RTOL, ATOL, "energy")
    assert_close(out["grad"], gr, RTOL, ATOL, "grad")
    # energy-only mode returns the same energies
    out2 = dftd3_blocked(numbers, positions, par, refcn=refcn, refc6=refc6, rcov=rcov[numbers], r4r2=r4r2[numbers],
                         rvdw_table=rvdw, g=g, cutoff=cutoff, want_grad=False)
    assert np.array_equal(out2["energy"], out["energy"])
