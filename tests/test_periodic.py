"""
Periodic images (`tad_dftd3_b200.periodic`).

The reference has no periodic boundary conditions, so there is no reference output to pin; what is pinned is the
DEFINITION: the energies of the atoms of the cell equal those of the same atoms inside an explicitly replicated
finite cluster evaluated with the reference's finite-structure rules (hard cutoffs, ordered ATM rule, CN cutoff 25).
CPU tests drive the front end with the blocked fp64 C oracle (`oracle/d3_blocked.c`, pinned to the unmodified
reference by `test_blocked_oracle.py`) as the evaluator; the GPU tests run the product path and compare with it.
"""
import itertools

import numpy as np
import pytest
import torch

from helpers import assert_close, tables

DT = torch.double
PARAM = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}


def water_cell(seed=0):
    """Two water molecules in a triclinic cell of about 10 x 11 x 12 Bohr."""
    g = torch.Generator().manual_seed(seed)
    numbers = torch.tensor([8, 1, 1, 8, 1, 1])
    mol = torch.tensor([[0.0, 0.0, 0.0], [1.81, 0.0, 0.0], [-0.45, 1.75, 0.0]], dtype=DT)
    positions = torch.cat([mol + torch.tensor([1.0, 1.5, 2.0], dtype=DT),
                           mol @ torch.tensor([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 0.0]], dtype=DT)
                           + torch.tensor([5.5, 6.0, 7.0], dtype=DT)])
    positions = positions + 0.05 * torch.randn(positions.shape, generator=g, dtype=DT)
    cell = torch.tensor([[10.0, 0.0, 0.0], [1.2, 11.0, 0.0], [0.7, -0.9, 12.0]], dtype=DT)
    return numbers, positions, cell


def blocked_evaluator(atm_table=True):
    """`_energy_fn` for dftd3_periodic: the blocked C oracle as a differentiable torch function of the positions."""
    from d3_blocked import dftd3_blocked

    refcn, refc6, rcov, r4r2, rvdw = tables(DT)

    def run(numbers, positions, param, cutoff, g, want_grad):
        par = {k: float(v) for k, v in param.items()}
        return dftd3_blocked(numbers, positions.detach(), par, refcn=refcn, refc6=refc6, rcov=rcov[numbers],
                             r4r2=r4r2[numbers], rvdw_table=rvdw if atm_table else None, g=g,
                             cutoff=50.0 if cutoff is None else float(cutoff), want_grad=want_grad)

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, positions, numbers, param, cutoff):
            ctx.save_for_backward(positions)
            ctx.args = (numbers, param, cutoff)
            return torch.from_numpy(run(numbers, positions, param, cutoff, None, False)["energy"])

        @staticmethod
        def backward(ctx, ge):
            (positions,) = ctx.saved_tensors
            numbers, param, cutoff = ctx.args
            out = run(numbers, positions, param, cutoff, ge.detach().contiguous(), True)
            return torch.from_numpy(out["grad"]), None, None, None

    def energy_fn(numbers, positions, param, *, cutoff=None, **kw):
        return Fn.apply(positions, numbers, param, cutoff)

    return energy_fn


def replica_cluster(numbers, positions, cell, reps):
    """Explicit block of (2 reps + 1)^3 cells, the cell itself first."""
    zs, xs = [numbers], [positions]
    for n in itertools.product(range(-reps[0], reps[0] + 1), range(-reps[1], reps[1] + 1), range(-reps[2], reps[2] + 1)):
        if n == (0, 0, 0):
            continue
        zs.append(numbers)
        xs.append(positions + torch.tensor(n, dtype=DT) @ cell)
    return torch.cat(zs), torch.cat(xs)


@pytest.mark.parametrize("pbc", [(True, True, True), (True, True, False), (False, True, False)])
def test_images_match_brute_force_enumeration(pbc):
    from tad_dftd3_b200.periodic import lattice_images

    numbers, positions, cell = water_cell()
    numbers = torch.cat([numbers, torch.tensor([0])])                      # one padding atom: no images
    positions = torch.cat([positions, torch.zeros(1, 3, dtype=DT)])
    radius = 21.0
    index, shifts = lattice_images(numbers, positions, cell, radius, pbc)
    n = numbers.shape[0]
    assert torch.equal(index[:n], torch.arange(n)) and not shifts[:n].any()
    got = {(int(i), *map(int, s)) for i, s in zip(index[n:], shifts[n:])}
    assert len(got) == index.shape[0] - n, "an image is listed twice"
    want = set()
    real = [i for i in range(n) if numbers[i] != 0]
    rng = [range(-5, 6) if p else range(0, 1) for p in pbc]
    for s in itertools.product(*rng):
        if s == (0, 0, 0):
            continue
        t = torch.tensor(s, dtype=DT) @ cell
        for i in real:
            if float((positions[i] + t - positions[real]).norm(dim=-1).min()) <= radius:
                want.add((i, *s))
    assert got == want


@pytest.mark.parametrize("atm", [False, True])
def test_periodic_energy_equals_explicit_replica_cluster(atm):
    """Definition check: cell atoms in the selected image shell == cell atoms in a much larger explicit block."""
    from tad_dftd3_b200.periodic import dftd3_periodic, image_radius

    numbers, positions, cell = water_cell()
    cutoff = 9.0 if atm else 11.0
    param = dict(PARAM, s9=1.0) if atm else PARAM
    fn = blocked_evaluator()
    e = dftd3_periodic(numbers, positions, cell, param, cutoff=cutoff, _energy_fn=fn)
    radius = image_radius(cutoff, atm)
    reps = [int(np.ceil(radius / 9.5)) + 1] * 3
    zc, xc = replica_cluster(numbers, positions, cell, reps)
    ec = fn(zc, xc, param, cutoff=cutoff)[: numbers.shape[0]]
    assert_close(e, ec, 1e-12, 1e-16, "cell atoms: image shell vs explicit block")
    # and the shell is needed: without images the energies differ
    e0 = fn(numbers, positions, param, cutoff=cutoff)
    assert float((e - e0).abs().max()) > 1e-3 * float(e.abs().max())


def test_supercell_and_lattice_translation_invariance():
    from tad_dftd3_b200.periodic import dftd3_periodic

    numbers, positions, cell = water_cell(1)
    fn = blocked_evaluator()
    # (the cutoff must not coincide with a lattice vector length, 10.0 here: an atom and its own image would sit
    # exactly on the step of the hard cutoff)
    e = dftd3_periodic(numbers, positions, cell, PARAM, cutoff=10.3, _energy_fn=fn)
    # 2 x 1 x 1 supercell: every atom twice, same energies
    z2 = torch.cat([numbers, numbers])
    x2 = torch.cat([positions, positions + cell[0]])
    c2 = cell.clone()
    c2[0] = 2 * cell[0]
    e2 = dftd3_periodic(z2, x2, c2, PARAM, cutoff=10.3, _energy_fn=fn)
    assert_close(e2, torch.cat([e, e]), 1e-12, 1e-16, "supercell")
    # atoms moved by lattice vectors (an unwrapped trajectory): same crystal
    moved = positions.clone()
    moved[1] += cell[2] - cell[0]
    moved[4] -= 2 * cell[1]
    e3 = dftd3_periodic(numbers, moved, cell, PARAM, cutoff=10.3, _energy_fn=fn)
    assert_close(e3, e, 1e-12, 1e-16, "lattice translation of single atoms")
    # one periodic direction switched off changes the result
    e4 = dftd3_periodic(numbers, positions, cell, PARAM, cutoff=10.3, pbc=(True, True, False), _energy_fn=fn)
    assert float((e4 - e).abs().max()) > 1e-7


@pytest.mark.parametrize("atm", [False, True])
def test_periodic_gradients_against_finite_differences(atm):
    """Gradient folded back from the images onto the cell atoms, and dE/dcell, against central differences of the
    periodic energy itself (same evaluator)."""
    from tad_dftd3_b200.periodic import dftd3_periodic

    numbers, positions, cell = water_cell(2)
    cutoff = 8.5 if atm else 10.3
    param = dict(PARAM, s9=1.0) if atm else PARAM
    fn = blocked_evaluator()
    w = torch.linspace(0.5, 1.5, numbers.shape[0], dtype=DT)                # cotangent of the per-atom energies

    def total(x, c):
        return (w * dftd3_periodic(numbers, x, c, param, cutoff=cutoff, _energy_fn=fn)).sum()

    x = positions.clone().requires_grad_(True)
    c = cell.clone().requires_grad_(True)
    gx, gc = torch.autograd.grad(total(x, c), (x, c))
    h = 1e-4
    fx = torch.zeros_like(positions)
    for i in range(positions.shape[0]):
        for k in range(3):
            d = torch.zeros_like(positions)
            d[i, k] = h
            fx[i, k] = (total(positions + d, cell) - total(positions - d, cell)) / (2 * h)
    fc = torch.zeros_like(cell)
    for i in range(3):
        for k in range(3):
            d = torch.zeros_like(cell)
            d[i, k] = h
            fc[i, k] = (total(positions, cell + d) - total(positions, cell - d)) / (2 * h)
    assert_close(gx, fx, 2e-6, 1e-11, "dE/dpositions")
    assert_close(gc, fc, 2e-6, 1e-11, "dE/dcell")
    # rigid translation of all atoms of the cell leaves the crystal unchanged
    assert float(gx.sum(0).abs().max()) < 1e-12


def test_argument_errors():
    from tad_dftd3_b200.periodic import dftd3_periodic, lattice_images

    numbers, positions, cell = water_cell()
    with pytest.raises(ValueError, match="not consistent"):
        dftd3_periodic(numbers, positions[:-1], cell, PARAM, _energy_fn=blocked_evaluator())
    with pytest.raises(ValueError, match=r"\[3, 3\]"):
        lattice_images(numbers, positions, cell[:2], 10.0)
    with pytest.raises(ValueError, match="singular"):
        lattice_images(numbers, positions, torch.zeros(3, 3, dtype=DT), 10.0)
    with pytest.raises(ValueError, match="cell first"):
        dftd3_periodic(numbers, positions, cell, PARAM, _energy_fn=blocked_evaluator(),
                       images=(torch.arange(3), torch.zeros(3, 3, dtype=torch.int64)))


# ---------------------------------------------------------------------------
# product path (CUDA) against the oracle evaluator
# ---------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cutoff, atm", [(8.0, False), (16.0, False), (8.5, True)])
def test_periodic_product_path_matches_oracle(cutoff, atm):
    """cutoff 8: ~700 atoms in the image shell (fused kernel); cutoff 16: ~1300 (tiled sweeps); ATM at cutoff 8.5
    (shell 2 x 8.5 + 25, ~1400 atoms) through the tiled three-body sweep with the registered element-pair radii."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200.periodic import dftd3_periodic

    dev = "cuda"
    numbers, positions, cell = water_cell(3)
    param = dict(PARAM, s9=1.0) if atm else PARAM
    w = torch.linspace(0.5, 1.5, numbers.shape[0], dtype=DT)
    # oracle
    x0 = positions.clone().requires_grad_(True)
    c0 = cell.clone().requires_grad_(True)
    e0 = dftd3_periodic(numbers, x0, c0, param, cutoff=cutoff, _energy_fn=blocked_evaluator())
    gx0, gc0 = torch.autograd.grad((w * e0).sum(), (x0, c0))
    # product
    refcn, refc6, rcov, r4r2, rvdw = tables(DT, dev)
    par = {k: torch.tensor(v, dtype=DT, device=dev) for k, v in param.items()}
    x1 = positions.to(dev).requires_grad_(True)
    c1 = cell.to(dev).requires_grad_(True)
    d3.data.set_vdw_pairwise(rvdw.cpu())
    try:
        e1 = dftd3_periodic(numbers.to(dev), x1, c1, par, ref=d3.Reference(cn=refcn, c6=refc6),
                            cutoff=torch.tensor(cutoff, dtype=DT, device=dev))
        gx1, gc1 = torch.autograd.grad((w.to(dev) * e1).sum(), (x1, c1))
    finally:
        d3.data.set_vdw_pairwise(None)
    assert e1.shape == numbers.shape and e1.device.type == "cuda"
    assert_close(e1.detach().cpu(), e0.detach(), 1e-10, 1e-12, "periodic energies")
    assert_close(gx1.cpu(), gx0, 1e-10, 1e-12, "periodic dE/dpositions")
    assert_close(gc1.cpu(), gc0, 1e-10, 1e-12, "periodic dE/dcell")


def test_image_list_reuse_and_per_atom_arguments():
    """An image list built with some slack on the radius serves displaced positions (MD / optimisation steps);
    per-atom `rcov` / `r4r2` and a dense `rvdw` of the cell are expanded to the images."""
    from tad_dftd3_b200.periodic import _gather, dftd3_periodic, image_radius, lattice_images

    numbers, positions, cell = water_cell(4)
    fn = blocked_evaluator()
    cutoff = 9.3
    images = lattice_images(numbers, positions, cell, image_radius(cutoff, False) + 0.5)
    g = torch.Generator().manual_seed(7)
    moved = positions + 0.1 * (torch.rand(positions.shape, generator=g, dtype=DT) - 0.5)      # |dx| < 0.09 Bohr
    e_reuse = dftd3_periodic(numbers, moved, cell, PARAM, cutoff=cutoff, images=images, _energy_fn=fn)
    e_fresh = dftd3_periodic(numbers, moved, cell, PARAM, cutoff=cutoff, _energy_fn=fn)
    assert_close(e_reuse, e_fresh, 1e-12, 1e-16, "reused image list")
    assert images[0].shape[0] > lattice_images(numbers, moved, cell, image_radius(cutoff, False))[0].shape[0]

    index = images[0]
    n = numbers.shape[0]
    per_atom = torch.arange(n, dtype=DT)
    assert torch.equal(_gather(per_atom, index, n), per_atom[index])
    dense = torch.arange(n * n, dtype=DT).reshape(n, n)
    got = _gather(dense, index, n)
    assert got.shape == (index.shape[0], index.shape[0]) and torch.equal(got[7, 11], dense[index[7], index[11]])
    table = torch.zeros(104, 104, dtype=DT)
    assert _gather(table, index, n) is table and _gather(None, index, n) is None
    seen = {}

    def spy(z, x, param, **kw):
        seen.update(kw)
        return torch.zeros(z.shape[0], dtype=DT)

    dftd3_periodic(numbers, positions, cell, PARAM, cutoff=cutoff, images=images, rcov=per_atom, r4r2=per_atom + 1,
                   rvdw=dense, _energy_fn=spy)
    assert torch.equal(seen["rcov"], per_atom[index]) and torch.equal(seen["r4r2"], per_atom[index] + 1)
    assert seen["rvdw"].shape == (index.shape[0],) * 2
