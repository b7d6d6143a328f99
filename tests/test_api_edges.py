"""
GPU tests of the call surface at its edges: empty and degenerate shapes, extra leading
batch dimensions, non-contiguous views, int32 numbers, parameters of another dtype or
device -- everything `tad_dftd3.dftd3` accepts (reference disp.py:79-165) has to be
accepted here with the same result.
"""
import numpy as np
import pytest
import torch

from helpers import assert_close, oracle_energy, tables

pytestmark = pytest.mark.gpu
DEV = "cuda"
PARAM = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}


def _ref(dtype=torch.double):
    import tad_dftd3_b200 as d3

    refcn, refc6, *_ = tables(dtype, DEV)
    return d3.Reference(cn=refcn, c6=refc6)


def _batch(nmol=6, seed=2):
    from tad_dftd3_b200 import synthetic as syn

    return syn.drug_like_batch(nmol, seed=seed, nmin=8, nmax=20)


def test_empty_shapes():
    import tad_dftd3_b200 as d3

    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    for shape in ((0,), (0, 7), (3, 0)):
        z = torch.zeros(shape, dtype=torch.int64, device=DEV)
        x = torch.zeros(*shape, 3, dtype=torch.double, device=DEV, requires_grad=True)
        e = d3.dftd3(z, x, par, ref=_ref())
        assert e.shape == shape and e.dtype == torch.double
        (g,) = torch.autograd.grad(e.sum(), x, allow_unused=True)
        assert g is None or g.shape == x.shape
    e, g = d3.dftd3_host(torch.zeros(0, 5, dtype=torch.int64), torch.zeros(0, 5, 3, dtype=torch.double), par, ref=_ref())
    assert e.shape == (0, 5) and g.shape == (0, 5, 3)


def test_all_padding_structure_and_ghost_only_rows():
    import tad_dftd3_b200 as d3

    numbers, positions = _batch()
    numbers = numbers.clone()
    numbers[2] = 0                                   # a structure that is padding only
    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    x = positions.to(DEV).requires_grad_(True)
    e = d3.dftd3(numbers.to(DEV), x, par, ref=_ref())
    (g,) = torch.autograd.grad(e.sum(), x)
    assert (e[2] == 0).all() and (g[2] == 0).all()
    eo = oracle_energy(numbers, positions, par, 50.0, torch.double)
    assert_close(e.detach().cpu(), eo, 1e-10, 1e-12, "energies with an empty row")


def test_extra_leading_dimensions_and_views():
    """[2, 3, N] batches, a transposed (non-contiguous) positions view, int32 numbers."""
    import tad_dftd3_b200 as d3

    numbers, positions = _batch(6)
    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    ref = _ref()
    flat = d3.dftd3(numbers.to(DEV), positions.to(DEV), par, ref=ref)
    n = numbers.shape[-1]
    e = d3.dftd3(numbers.reshape(2, 3, n).to(DEV), positions.reshape(2, 3, n, 3).to(DEV), par, ref=ref)
    assert e.shape == (2, 3, n) and torch.equal(e.reshape(6, n), flat)
    xt = positions.transpose(-1, -2).contiguous().to(DEV).transpose(-1, -2)      # strides (.., 1, N)
    assert not xt.is_contiguous()
    xt.requires_grad_(True)
    e2 = d3.dftd3(numbers.to(DEV), xt, par, ref=ref)
    # (the energy + gradient kernel is another instantiation than the energy-only one: last-bit differences)
    assert_close(e2.detach().cpu(), flat.cpu(), 1e-13, 1e-18, "energies from a non-contiguous view")
    (g2,) = torch.autograd.grad(e2.sum(), xt)
    x = positions.to(DEV).requires_grad_(True)
    e1 = d3.dftd3(numbers.to(DEV), x, par, ref=ref)
    (g1,) = torch.autograd.grad(e1.sum(), x)
    assert torch.equal(e1.detach(), e2.detach()) and torch.equal(g1, g2)
    e3 = d3.dftd3(numbers.to(torch.int32).to(DEV), positions.to(DEV), par, ref=ref)
    assert torch.equal(e3, flat)


def test_parameters_of_other_dtype_and_device():
    """float32 / CPU / python-float parameters with float64 CUDA positions: values are used as stored
    (reference README: `torch.tensor(0.49484001)` is float32)."""
    import tad_dftd3_b200 as d3

    numbers, positions = _batch(3)
    ref = _ref()
    z, x = numbers.to(DEV), positions.to(DEV)
    p32 = {k: torch.tensor(v) for k, v in PARAM.items()}                        # float32 on CPU
    p32_cuda = {k: v.to(DEV) for k, v in p32.items()}
    p_as64 = {k: torch.tensor(float(v), dtype=torch.double) for k, v in p32.items()}
    p_float = {k: float(v) for k, v in p32.items()}
    e = d3.dftd3(z, x, p_as64, ref=ref)
    for p in (p32, p32_cuda, p_float):
        assert torch.equal(d3.dftd3(z, x, p, ref=ref), e)
    eo = oracle_energy(numbers, positions, p_as64, 50.0, torch.double)
    assert_close(e.cpu(), eo, 1e-10, 1e-12, "float32-rounded parameters")
    exact = d3.dftd3(z, x, {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}, ref=ref)
    assert not torch.equal(exact, e)                                            # 1e-7 relative apart


def test_host_entry_pageable_and_pinned_inputs_agree():
    """Pageable inputs take the staged pipeline, pinned ones the zero-copy launch: same bits."""
    import tad_dftd3_b200 as d3

    numbers, positions = _batch(40, seed=5)
    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    ref = _ref()
    e1, g1 = d3.dftd3_host(numbers, positions, par, ref=ref)                     # pageable -> staged
    e2, g2 = d3.dftd3_host(numbers.pin_memory(), positions.pin_memory(), par, ref=ref)   # pinned -> zero-copy
    assert np.array_equal(e1.numpy(), e2.numpy()) and np.array_equal(g1.numpy(), g2.numpy())
    pad = (numbers == 0).numpy()
    assert (e2.numpy()[pad] == 0).all() and (g2.numpy()[pad] == 0).all()
    x = positions.to(DEV).requires_grad_(True)
    ed = d3.dftd3(numbers.to(DEV), x, par, ref=ref)
    (gd,) = torch.autograd.grad(ed.sum(), x)
    assert np.array_equal(e2.numpy(), ed.detach().cpu().numpy()) and np.array_equal(g2.numpy(), gd.cpu().numpy())


def test_batch_order_hint_does_not_change_results():
    """From the second call with the same numbers tensor on, CTAs take the structures largest first
    (`d3b200_model.batch_order`); energies and gradients must stay bitwise the same."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import ops
    from tad_dftd3_b200 import synthetic as syn

    numbers, positions = syn.drug_like_batch(1300, seed=21, nmin=4, nmax=14)
    assert numbers.shape[0] >= ops.ORDER_MIN_BATCH
    par = {k: torch.tensor(v, dtype=torch.double) for k, v in PARAM.items()}
    ref = _ref()
    z = numbers.to(DEV)
    outs = []
    for _ in range(3):
        x = positions.to(DEV).requires_grad_(True)
        e = d3.dftd3(z, x, par, ref=ref)
        (g,) = torch.autograd.grad(e.sum(), x)
        outs.append((e.detach().clone(), g.clone()))
    assert ops._order_cache[id(z)][3] is not None            # the hint was in use for calls 2 and 3
    order = ops._order_cache[id(z)][3].long()
    counts = (z != 0).sum(-1)
    assert torch.equal(torch.sort(order).values, torch.arange(z.shape[0], device=DEV))
    assert bool((counts[order][:-1] >= counts[order][1:]).all())
    for e, g in outs[1:]:
        assert torch.equal(e, outs[0][0]) and torch.equal(g, outs[0][1])
    nh, ph = numbers.pin_memory(), positions.pin_memory()
    e_h = torch.empty(numbers.shape, dtype=torch.double).pin_memory()
    g_h = torch.empty(*numbers.shape, 3, dtype=torch.double).pin_memory()
    for _ in range(3):
        e_h.zero_(); g_h.zero_()
        d3.dftd3_host(nh, ph, par, ref=ref, energy_out=e_h, grad_out=g_h)
        assert np.array_equal(e_h.numpy(), outs[0][0].cpu().numpy())
        assert np.array_equal(g_h.numpy(), outs[0][1].cpu().numpy())


def test_reference_tables_of_other_shapes_are_rejected_and_cache_follows_edits():
    """The kernels index the tables with fixed strides ([104,7] / [104,104,7,7]): any other mutually consistent
    shape (which the reference's holder accepts, reference.py:231-245) must be refused, not read out of bounds;
    editing a table in place invalidates the prepared device copy."""
    import tad_dftd3_b200 as d3

    dt = torch.double
    refcn, refc6, *_ = tables(dt, DEV)
    numbers = torch.tensor([8, 1, 1], device=DEV)
    pos = torch.tensor([[0.0, 0.0, -0.7], [1.4, 0.0, 0.4], [-1.4, 0.0, 0.4]], device=DEV, dtype=dt)
    par = {"s6": torch.tensor(1.0), "s8": torch.tensor(0.8)}
    small = d3.Reference(cn=refcn[:87, :5].contiguous(), c6=refc6[:87, :87, :5, :5].contiguous())   # consistent, but not 104 x 7
    with pytest.raises(ValueError, match="reference tables"):
        d3.dftd3(numbers, pos, par, ref=small)
    ref = d3.Reference(cn=refcn.clone(), c6=refc6.clone())
    e0 = d3.dftd3(numbers, pos, par, ref=ref)
    ref.c6.mul_(2.0)                                   # in-place edit: version counter changes
    e1 = d3.dftd3(numbers, pos, par, ref=ref)
    assert torch.allclose(e1, 2.0 * e0, rtol=1e-13, atol=0)


def test_arguments_without_derivatives_refuse_requires_grad():
    """alp / cutoff / rcov / r4r2 / rvdw are constants of the kernels: asking for their gradient raises instead of
    returning a silently detached result (the reference would differentiate them through autograd)."""
    import tad_dftd3_b200 as d3

    dt = torch.double
    refcn, refc6, rcov, r4r2, rvdw = tables(dt, DEV)
    ref = d3.Reference(cn=refcn, c6=refc6)
    numbers = torch.tensor([8, 1, 1], device=DEV)
    pos = torch.tensor([[0.0, 0.0, -0.7], [1.4, 0.0, 0.4], [-1.4, 0.0, 0.4]], device=DEV, dtype=dt)
    par = {"s6": torch.tensor(1.0), "s8": torch.tensor(0.8)}
    with pytest.raises(NotImplementedError, match="rcov"):
        d3.dftd3(numbers, pos, par, ref=ref, rcov=rcov[numbers].clone().requires_grad_(True))
    with pytest.raises(NotImplementedError, match="r4r2"):
        d3.dftd3(numbers, pos, par, ref=ref, r4r2=r4r2[numbers].clone().requires_grad_(True))
    with pytest.raises(NotImplementedError, match="alp"):
        d3.dftd3(numbers, pos, dict(par, s9=torch.tensor(1.0), alp=torch.tensor(14.0, requires_grad=True)), ref=ref,
                 rvdw=rvdw[numbers.unsqueeze(-1), numbers.unsqueeze(-2)])
    with pytest.raises(NotImplementedError, match="cutoff"):
        d3.dftd3(numbers, pos, par, ref=ref, cutoff=torch.tensor(30.0, dtype=dt, requires_grad=True))


def test_host_entry_non_contiguous_positions_are_re_read_every_call():
    """A non-contiguous `positions` needs a private contiguous copy: the prepared-call fast path must not keep
    returning the results of the first geometry when the caller updates its tensor in place."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn

    dt = torch.double
    refcn, refc6, *_ = tables(dt, DEV)
    ref = d3.Reference(cn=refcn, c6=refc6)
    numbers, positions = syn.drug_like_batch(6, seed=9, nmin=10, nmax=20)
    wide = torch.zeros(*positions.shape[:-1], 4, dtype=dt)
    wide[..., :3] = positions
    view = wide[..., :3]                                         # non-contiguous view
    assert not view.is_contiguous()
    par = {"s6": torch.tensor(1.0), "s8": torch.tensor(0.7898), "a1": torch.tensor(0.4948), "a2": torch.tensor(5.7308)}
    e_out = torch.empty(numbers.shape, dtype=dt).pin_memory()
    g_out = torch.empty(*numbers.shape, 3, dtype=dt).pin_memory()
    nh = numbers.pin_memory()
    for step in range(3):
        wide[..., :3] += 0.05 * torch.sin(wide[..., :3] + step)  # in-place update of the caller's tensor
        d3.dftd3_host(nh, view, par, ref=ref, energy_out=e_out, grad_out=g_out)
        x = view.contiguous().to(DEV).requires_grad_(True)
        e = d3.dftd3(numbers.to(DEV), x, par, ref=ref)
        (g,) = torch.autograd.grad(e.sum(), x)
        assert torch.equal(e_out.to(DEV), e.detach()) and torch.equal(g_out.to(DEV), g), step
