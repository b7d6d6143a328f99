"""
Pin the CPU oracle (`oracle/d3_oracle.py`) before anything trusts it:

* against golden values held by the reference's own tests
  (`reference_pins.npz`: cn, weights, disp2 -- extracted from
  `/root/reference/test/test_model/samples.py` and `test/test_disp/samples.py`);
* against outputs of the unmodified reference run in the build container
  (`reference_runs.npz`: energies, VJPs, parameter gradients, Hessians).

fp64 tolerances: 1e-10 relative / 1e-12 Hartree absolute (BASELINE.json).
"""
import numpy as np
import pytest
import torch

import d3_oracle as orc
from helpers import assert_close, case_inputs, oracle_energy, tables

RTOL, ATOL = 1e-10, 1e-12


@pytest.mark.parametrize("name", ["PbH4_BiH3", "C6H5I_CH3SH"])
def test_cn_pinned(pins, name):
    """CN from in-tree geometries vs `test/test_model/samples.py:626-639,834-856`."""
    c = pins.case(name)
    numbers = torch.from_numpy(c["numbers"])
    pos = torch.from_numpy(c["positions"])
    _, _, rcov, _, _ = tables(torch.double)
    cn = orc.cn_d3(numbers, pos, rcov[numbers])
    # README geometry of C6H5I-CH3SH carries fewer digits than the pinned run
    assert_close(cn, c["cn"], 0, 5e-11, name)


@pytest.mark.parametrize("name", ["SiH4", "MB16_43_01", "PbH4_BiH3", "C6H5I_CH3SH", "AmF3"])
def test_weights_pinned(pins, name):
    """cn -> weights vs `test/test_model/samples.py` (needs only refcn + Z)."""
    c = pins.case(name)
    numbers = torch.from_numpy(c["numbers"])
    refcn = tables(torch.double)[0]
    w = orc.weight_references(numbers, torch.from_numpy(c["cn"]), refcn)
    tol = 1e-7 if name == "AmF3" else 1e-13  # AmF3 pinned to float32 data precision
    assert_close(w, c["weights"].reshape(w.shape), 0, tol, name)


@pytest.mark.parametrize("name", ["PbH4_BiH3", "C6H5I_CH3SH"])
def test_disp2_pinned(pins, name):
    """dense pinned c6 -> disp2 with TPSS0-D3BJ, `test/test_disp/test_disp.py:34-46`."""
    c = pins.case(name)
    numbers = torch.from_numpy(c["numbers"])
    pos = torch.from_numpy(c["positions"])
    r4r2 = tables(torch.double)[3][numbers]
    e = orc.dispersion2(numbers, pos, torch.from_numpy(c["c6"]).reshape(len(numbers), -1), r4r2,
                        1.0, 1.2576, 0.3768, 4.5865, 50.0)
    assert_close(e, c["disp2"], 1e-9 if name == "C6H5I_CH3SH" else 1e-13, 2e-11 if name == "C6H5I_CH3SH" else 1e-16, name)


def _all_cases(runs):
    return [k for k in runs.cases if not k.startswith("stages")]


def test_runs_energy_and_grad(runs):
    """Energies and VJPs of every committed reference run."""
    for case in _all_cases(runs):
        c = runs.case(case)
        dtype, numbers, positions, g, param, cutoff = case_inputs(c)
        f32 = dtype == torch.float32
        pos = positions.clone().requires_grad_(True)
        par = {k: v.clone().requires_grad_(k in ("s6", "s8", "a1", "a2")) for k, v in param.items()}
        e = oracle_energy(numbers, pos, par, cutoff, dtype)
        rt, at = (2e-4, 1e-8) if f32 else (RTOL, ATOL)
        assert_close(e.detach(), c["energy"], rt, at, case + " energy")
        if not e.requires_grad:
            continue
        wrt = [pos] + [par[k] for k in ("s6", "s8", "a1", "a2") if k in par]
        grads = torch.autograd.grad((e * g).sum(), wrt, allow_unused=True)
        assert_close(grads[0], c["grad"], rt, 1e-7 if f32 else 1e-13, case + " grad")
        if "pgrad" in c:
            assert_close([float(x) for x in grads[1:]], c["pgrad"], rt, at, case + " pgrad")


@pytest.mark.parametrize("case", ["pbh4_hess_f64", "pbh4_hess_atm_f64", "two_atoms_f64", "heavy12_hess_atm_f64"])
def test_runs_hessian(runs, case):
    c = runs.case(case)
    dtype, numbers, positions, g, param, cutoff = case_inputs(c)

    def f(p):
        return oracle_energy(numbers, p, param, cutoff, dtype).sum(-1)

    h = torch.func.jacrev(torch.func.jacrev(f))(positions)
    assert_close(h, c["hessian"], 1e-9, 1e-13, case)


def test_stage_outputs(runs):
    c = runs.case("stages_c2")
    r = runs.case("c2_f64")
    numbers = torch.from_numpy(r["numbers"])
    pos = torch.from_numpy(r["positions"])
    refcn, refc6, rcov, r4r2, rvdw = tables(torch.double)
    cn = orc.cn_d3(numbers, pos, rcov[numbers])
    w = orc.weight_references(numbers, cn, refcn)
    c6 = orc.atomic_c6(numbers, w, refc6)
    assert_close(cn, c["cn"], RTOL, ATOL, "cn")
    assert_close(w, c["weights"], RTOL, ATOL, "weights")
    assert_close(c6, c["c6"], RTOL, ATOL, "c6")


def test_chunked_equals_full(runs):
    """`chunk_size` is a memory knob of the reference only (c6.py:220-273)."""
    a, b = runs.case("pbh4_hess_f64"), runs.case("pbh4_chunk2_f64")
    np.testing.assert_allclose(a["energy"], b["energy"], rtol=1e-14)
