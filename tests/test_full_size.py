"""
Parity at BASELINE.json's full sizes.

The CPU oracle cannot hold these inputs at once (the reference's C6 gather alone is
23 GB for config 3), so the full-size runs are checked through properties that do
not depend on the size -- translation / rotation / permutation invariance,
vanishing net force, energy-only vs energy+gradient kernels, row-partitioned vs
whole sweeps, fp32 vs fp64 -- and through the oracle on random samples of the
same inputs.
"""
import numpy as np
import pytest
import torch

from helpers import assert_close, oracle_energy, tables

pytestmark = pytest.mark.gpu
DEV = "cuda"
PARAM = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}


@pytest.fixture(scope="module")
def c3():
    from tad_dftd3_b200 import synthetic as syn

    numbers, positions = syn.drug_like_batch(4096, 50, 120, seed=0)
    return numbers.to(DEV), positions.to(DEV)


@pytest.fixture(scope="module")
def ref():
    import tad_dftd3_b200 as d3

    refcn, refc6, *_ = tables(torch.double, DEV)
    return d3.Reference(cn=refcn, c6=refc6)


def run(numbers, positions, ref, param=PARAM, **kw):
    import tad_dftd3_b200 as d3

    par = {k: torch.tensor(v, dtype=positions.dtype) for k, v in param.items()}
    x = positions.clone().requires_grad_(True)
    e = d3.dftd3(numbers, x, par, ref=ref.type(positions.dtype), **kw)
    (g,) = torch.autograd.grad(e.sum(), x)
    return e.detach(), g


def rotation(seed=0):
    q, _ = np.linalg.qr(np.random.RandomState(seed).normal(size=(3, 3)))
    return torch.from_numpy(q * np.sign(np.linalg.det(q))).to(DEV)


def test_c3_full_batch_properties(c3, ref):
    """4096 molecules, 50-120 atoms: invariances, zero net force, exact padding."""
    import tad_dftd3_b200 as d3

    numbers, positions = c3
    e, g = run(numbers, positions, ref)
    real = numbers != 0
    assert torch.isfinite(e).all() and torch.isfinite(g).all()
    assert (e[~real] == 0).all() and (g[~real] == 0).all() and (e[real] < 0).all()
    # translational invariance of every structure: net force vanishes
    net = g.sum(1).abs().max()
    assert float(net) < 1e-15 * numbers.shape[1] * 10, float(net)
    # rigid motion (rotation + shift applied to the real atoms only)
    R, shift = rotation(), torch.tensor([3.0, -2.0, 1.5], dtype=torch.double, device=DEV)
    moved = torch.where(real.unsqueeze(-1), positions @ R.T + shift, positions)
    e2, g2 = run(numbers, moved, ref)
    assert_close(e2.cpu(), e.cpu(), 1e-11, 1e-15, "rotated energies")
    assert_close(g2.cpu(), (g @ R.T).cpu(), 1e-9, 1e-14, "rotated gradients")
    # permutation of the atoms inside every structure (padding stays at the end)
    n = real.sum(1)
    keys = torch.rand(numbers.shape, device=DEV, generator=torch.Generator(DEV).manual_seed(1))
    keys = torch.where(real, keys, torch.full_like(keys, 2.0))
    perm = torch.argsort(keys, 1)
    ep, gp = run(torch.gather(numbers, 1, perm), torch.gather(positions, 1, perm.unsqueeze(-1).expand(-1, -1, 3)), ref)
    assert_close(ep.cpu(), torch.gather(e, 1, perm).cpu(), 1e-11, 1e-15, "permuted energies")
    assert_close(gp.cpu(), torch.gather(g, 1, perm.unsqueeze(-1).expand(-1, -1, 3)).cpu(), 1e-9, 1e-14, "permuted gradients")
    # energy-only kernel == energy of the fused energy+gradient kernel
    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    e_only = d3.dftd3(numbers, positions, par, ref=ref)
    assert_close(e_only.cpu(), e.cpu(), 1e-13, 1e-17, "energy-only kernel")
    # the batch is the concatenation of its parts (structures are independent)
    e_half, g_half = run(numbers[1000:1100], positions[1000:1100], ref)
    assert torch.equal(e_half, e[1000:1100]) and torch.equal(g_half, g[1000:1100])


def test_c3_samples_against_oracle(c3, ref):
    """48 random structures of the full batch against the CPU oracle (1e-10 / 1e-12)."""
    numbers, positions = c3
    e, g = run(numbers, positions, ref)
    pick = np.random.RandomState(3).choice(numbers.shape[0], 48, replace=False)
    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    for b in pick:
        nb, xb = numbers[b].cpu(), positions[b].cpu().requires_grad_(True)
        eo = oracle_energy(nb, xb, par, 50.0, torch.double)
        (go,) = torch.autograd.grad(eo.sum(), xb)
        assert_close(e[b].cpu(), eo.detach(), 1e-10, 1e-12, f"structure {b} energy")
        assert_close(g[b].cpu(), go, 1e-10, 1e-12, f"structure {b} gradient")


def test_c3_fp32_tracks_fp64(c3, ref):
    numbers, positions = c3
    e64, g64 = run(numbers[:512], positions[:512], ref)
    e32, g32 = run(numbers[:512], positions[:512].float(), ref)
    assert e32.dtype == torch.float32
    assert_close(e32.cpu(), e64.cpu(), 1e-5, 1e-9, "fp32 energies vs fp64")
    assert_close(g32.cpu(), g64.cpu(), 2e-4, 5e-8, "fp32 gradients vs fp64")


@pytest.fixture(scope="module")
def c4():
    from tad_dftd3_b200 import synthetic as syn

    numbers, positions = syn.water_cluster(16667, seed=0)
    return numbers.to(DEV), positions.to(DEV)


def test_c4_50k_atoms_properties(c4, ref):
    """50 001-atom cluster through the tiled sweeps: finite, net force zero, rigid-motion
    invariant, row-partitioned sweeps bitwise equal to whole sweeps, sample atoms vs a
    brute-force neighbourhood evaluation by the oracle."""
    from tad_dftd3_b200 import _lib
    from tad_dftd3_b200.data import COV_D3, R4R2
    from tad_dftd3_b200.system import SystemPlan, SystemRunner, row_partition

    numbers, positions = c4
    e, g = run(numbers, positions, ref)
    assert torch.isfinite(e).all() and torch.isfinite(g).all() and (e < 0).all()
    assert float(g.sum(0).abs().max()) < 1e-12
    R, shift = rotation(5), torch.tensor([-7.0, 2.0, 11.0], dtype=torch.double, device=DEV)
    e2, g2 = run(numbers, positions @ R.T + shift, ref)
    assert_close(e2.cpu(), e.cpu(), 1e-10, 1e-14, "rotated energies")
    assert_close(g2.cpu(), (g @ R.T).cpu(), 1e-8, 1e-13, "rotated gradients")

    # rows in three chunks (what three ranks would own) == all rows at once
    dt = torch.double
    refcn, refc6, _ = ref._prepared(torch.device(DEV, 0), dt)
    par = _lib.Params()
    par.s6, par.s8, par.a1, par.a2 = PARAM["s6"], PARAM["s8"], PARAM["a1"], PARAM["a2"]
    par.s9, par.rs9, par.alp, par.cutoff, par.cn_cutoff, par.kcn = 0.0, 4 / 3, 14.0, 50.0, 25.0, 16.0
    plan = SystemPlan(numbers, positions, COV_D3(dt, DEV)[numbers], R4R2(dt, DEV)[numbers], par.kcn)
    whole = SystemRunner(plan, refcn, refc6, par)
    whole.run_vjp()
    parts = SystemRunner(plan, refcn, refc6, par, None, 3)
    chunks = row_partition(plan.npad, 3)
    for rows in chunks:
        parts.sweep_cn(rows)
    parts.weights()
    for rows in chunks:
        parts.sweep_pair(True, rows)
    for rows in chunks:
        parts.sweep_cnbwd(rows)
    # jsplit differs between the two runners, so compare to rounding, not bitwise
    assert_close(parts.energy.cpu(), whole.energy.cpu(), 1e-12, 1e-16, "row-partitioned energies")
    assert_close(parts.grad.cpu(), whole.grad.cpu(), 1e-10, 1e-15, "row-partitioned gradients")
    assert_close(plan.scatter(whole.energy).cpu(), e.cpu(), 1e-12, 1e-16, "runner vs dftd3()")


def test_c4_sample_atoms_against_oracle(c4, ref):
    """Energies of atoms near the centre of the 50k cluster equal those of the same atoms
    in the sub-cluster of everything within 50 + 25 Bohr of them (the interaction range
    of an atom is cutoff + CN cutoff), evaluated by the CPU oracle."""
    numbers, positions = c4
    e, _ = run(numbers, positions, ref)
    centre = positions.mean(0)
    d0 = (positions - centre).norm(dim=1)
    probe = torch.argsort(d0)[:3]                       # three central atoms
    # shrink the cutoff so that the oracle's dense sub-cluster stays small
    cutoff = 14.0
    e_small, _ = run(numbers, positions, ref, cutoff=torch.tensor(cutoff))
    keep = (positions - centre).norm(dim=1) <= float(d0[probe].max()) + cutoff + 25.0 + 1e-6
    idx = torch.nonzero(keep).reshape(-1)
    assert idx.numel() < 6000
    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    eo = oracle_energy(numbers[idx].cpu(), positions[idx].cpu(), par, cutoff, torch.double)
    where = {int(a): k for k, a in enumerate(idx.tolist())}
    for a in probe.tolist():
        # CN of the probe's partners needs atoms up to 25 Bohr beyond the dispersion cutoff: kept above
        assert abs(float(e_small[a]) - float(eo[where[a]])) <= 1e-10 * abs(float(eo[where[a]])) + 1e-12


def test_c5a_atm_10k_atoms_properties(ref):
    """ATM on the 10 002-atom cluster (cutoff 25): finite, net force zero, rigid-motion
    invariant, and the general kernel equals the T-vector kernel."""
    import os

    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn

    numbers, positions = syn.water_cluster(3334, seed=0)
    numbers, positions = numbers.to(DEV), positions.to(DEV)
    d3.data.set_vdw_pairwise(syn.synthetic_rvdw_table(torch.double))
    try:
        par = dict(PARAM, s9=1.0)
        kw = dict(cutoff=torch.tensor(25.0))
        e, g = run(numbers, positions, ref, par, **kw)
        e2b, _ = run(numbers, positions, ref, PARAM, **kw)
        assert torch.isfinite(e).all() and torch.isfinite(g).all()
        assert float((e - e2b).sum()) > 0          # the three-body term is repulsive on balance here
        assert float(g.sum(0).abs().max()) < 1e-12
        R = rotation(9)
        e2, g2 = run(numbers, positions @ R.T, ref, par, **kw)
        assert_close(e2.cpu(), e.cpu(), 1e-10, 1e-14, "rotated ATM energies")
        assert_close(g2.cpu(), (g @ R.T).cpu(), 1e-8, 1e-13, "rotated ATM gradients")
        os.environ["TAD_DFTD3_B200_NO_TVEC"] = "1"
        try:
            e3, g3 = run(numbers[:1500], positions[:1500], ref, par, **kw)
        finally:
            del os.environ["TAD_DFTD3_B200_NO_TVEC"]
        e4, g4 = run(numbers[:1500], positions[:1500], ref, par, **kw)
        assert_close(e3.cpu(), e4.cpu(), 1e-11, 1e-15, "general vs T-vector ATM energies")
        assert_close(g3.cpu(), g4.cpu(), 1e-9, 1e-14, "general vs T-vector ATM gradients")
        # triples-once kernel (default) against the edge-owner T-vector kernel on the full cluster
        os.environ["TAD_DFTD3_B200_ATM2"] = "1"
        try:
            e5, g5 = run(numbers, positions, ref, par, **kw)
        finally:
            del os.environ["TAD_DFTD3_B200_ATM2"]
        assert_close(e5.cpu(), e.cpu(), 1e-10, 1e-14, "edge-owner vs triples-once ATM energies")
        assert_close(g5.cpu(), g.cpu(), 1e-8, 1e-13, "edge-owner vs triples-once ATM gradients")
    finally:
        d3.data.set_vdw_pairwise(None)


# ---------------------------------------------------------------------------
# Full-size comparisons with the blocked fp64 C oracle (oracle/d3_blocked.c, pinned
# against the unmodified reference by tests/test_blocked_oracle.py): EVERY atom,
# energies and gradients, at BASELINE.json's tolerances (1e-10 relative, 1e-12 absolute).
# ---------------------------------------------------------------------------
def _blocked(numbers, positions, param, cutoff, g=None):
    from d3_blocked import dftd3_blocked

    refcn, refc6, rcov, r4r2, rvdw = tables(torch.double)
    z = numbers.cpu()
    return dftd3_blocked(z, positions.cpu(), param, refcn=refcn, refc6=refc6, rcov=rcov[z], r4r2=r4r2[z],
                         rvdw_table=rvdw, g=None if g is None else g.cpu(), cutoff=cutoff)


def test_c4_50k_atoms_every_atom_against_blocked_oracle(c4, ref):
    """Config 4 at the default cutoffs (50 / 25 Bohr): energies AND gradients of all 50 001 atoms
    (core and surface alike), unit cotangent and a random cotangent."""
    import tad_dftd3_b200 as d3

    numbers, positions = c4
    e, g = run(numbers, positions, ref)
    out = _blocked(numbers, positions, PARAM, 50.0)
    assert_close(e.cpu(), out["energy"], 1e-10, 1e-12, "c4 energies, all atoms")
    assert_close(g.cpu(), out["grad"], 1e-10, 1e-12, "c4 gradients, all atoms")
    # general cotangent: takes the D3Vjp route of the tiled path
    gen = torch.Generator().manual_seed(17)
    w = (torch.rand(numbers.shape, dtype=torch.double, generator=gen) + 0.5)
    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    x = positions.clone().requires_grad_(True)
    ew = d3.dftd3(numbers, x, par, ref=ref)
    (gw,) = torch.autograd.grad((ew * w.to(DEV)).sum(), x)
    outw = _blocked(numbers, positions, PARAM, 50.0, g=w)
    assert_close(gw.cpu(), outw["grad"], 1e-10, 1e-12, "c4 gradients, random cotangent")


def test_c5a_atm_10k_atoms_every_atom_against_blocked_oracle(ref):
    """Config 5a (10 002 atoms, s9 = 1, cutoff 25 Bohr): energies and gradients of every atom."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn

    numbers, positions = syn.water_cluster(3334, seed=0)
    numbers, positions = numbers.to(DEV), positions.to(DEV)
    d3.data.set_vdw_pairwise(syn.synthetic_rvdw_table(torch.double))
    try:
        par = dict(PARAM, s9=1.0)
        e, g = run(numbers, positions, ref, par, cutoff=torch.tensor(25.0))
    finally:
        d3.data.set_vdw_pairwise(None)
    out = _blocked(numbers, positions, par, 25.0)
    assert_close(e.cpu(), out["energy"], 1e-10, 1e-12, "c5a energies, all atoms")
    assert_close(g.cpu(), out["grad"], 1e-10, 1e-12, "c5a gradients, all atoms")
