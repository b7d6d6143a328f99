"""
GPU tests of the host-buffer entry (`dftd3_host` -> d3b200_batch_vjp_host_*):
the chunked H2D / kernel / D2H pipeline must give exactly what `dftd3()` +
autograd give on device tensors, and the committed reference values.
"""
import numpy as np
import pytest
import torch

from helpers import assert_close, case_inputs, tables

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref(dtype):
    import tad_dftd3_b200 as d3

    refcn, refc6, *_ = tables(dtype, DEV)
    return d3.Reference(cn=refcn, c6=refc6)


@pytest.mark.parametrize("case", ["mols6_f64", "c2_f64"])
def test_host_entry_matches_reference_runs(runs, case):
    import tad_dftd3_b200 as d3

    c = runs.case(case)
    dtype, numbers, positions, g, param, cutoff = case_inputs(c)
    e, gr = d3.dftd3_host(numbers, positions, param, ref=_ref(dtype), cutoff=cutoff, g=g, chunk=1)
    assert e.device.type == "cpu" and gr.device.type == "cpu"
    assert_close(e.numpy(), c["energy"], 1e-10, 1e-12, case + " energy")
    assert_close(gr.numpy(), c["grad"], 1e-10, 1e-12, case + " grad")
    pad = (numbers == 0).numpy()
    assert (e.numpy()[pad] == 0).all() and (gr.numpy()[pad] == 0).all()


@pytest.mark.parametrize("dtype", [torch.double, torch.float32])
@pytest.mark.parametrize("chunk", [0, 5, 16, -16, -3])
def test_host_entry_bitwise_equals_device_path(dtype, chunk):
    """Chunking must not change a single bit: structures are independent and the
    kernel's summation order depends on the structure only.  chunk 0: zero-copy; > 0: staged;
    -16: hybrid (inputs staged, results written straight into host memory); -3: streamed (one
    launch whose CTAs wait for per-chunk flags set behind the copies, 2^3 structures per chunk)."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn

    numbers, positions = syn.drug_like_batch(37, seed=3, nmin=20, nmax=60)
    positions = positions.to(dtype)
    param = {"s6": torch.tensor(1.0), "s8": torch.tensor(0.7898), "a1": torch.tensor(0.4948), "a2": torch.tensor(5.7308)}
    ref = _ref(dtype)
    gen = torch.Generator().manual_seed(5)
    g = torch.rand(numbers.shape, generator=gen, dtype=dtype)
    for cot in (None, g):
        nh, ph = numbers.pin_memory(), positions.pin_memory()
        e, gr = d3.dftd3_host(nh, ph, param, ref=ref, g=cot, chunk=chunk)
        x = positions.to(DEV).requires_grad_(True)
        ed = d3.dftd3(numbers.to(DEV), x, param, ref=ref)
        w = torch.ones_like(ed) if cot is None else cot.to(DEV)
        (gd,) = torch.autograd.grad((ed * w).sum(), x)
        assert np.array_equal(e.numpy(), ed.detach().cpu().numpy())
        assert np.array_equal(gr.numpy(), gd.cpu().numpy())


def test_host_entry_atm_and_repeated_calls(runs):
    """ATM through the host entry (element-pair rvdw table) + buffer reuse over repeated asynchronous calls."""
    import tad_dftd3_b200 as d3

    c = runs.case("mols6_atm_f64")
    dtype, numbers, positions, g, param, cutoff = case_inputs(c)
    rvdw = tables(dtype)[4]
    d3.data.set_vdw_pairwise(rvdw)
    try:
        e_out = torch.empty(numbers.shape, dtype=dtype).pin_memory()
        g_out = torch.empty(*numbers.shape, 3, dtype=dtype).pin_memory()
        for _ in range(3):
            e_out.zero_(); g_out.zero_()
            d3.dftd3_host(numbers, positions, param, ref=_ref(dtype), cutoff=cutoff, g=g, chunk=2,
                          energy_out=e_out, grad_out=g_out, sync=False)
            torch.cuda.current_stream().synchronize()
            assert_close(e_out.numpy(), c["energy"], 1e-10, 1e-12, "atm energy")
            assert_close(g_out.numpy(), c["grad"], 1e-10, 1e-12, "atm grad")
    finally:
        d3.data.set_vdw_pairwise(None)


def test_host_entry_rejects_device_tensors():
    import tad_dftd3_b200 as d3

    z = torch.tensor([1, 1], device=DEV)
    x = torch.zeros(2, 3, device=DEV)
    with pytest.raises(RuntimeError, match="CPU tensors"):
        d3.dftd3_host(z, x, {})
    with pytest.raises(ValueError):
        d3.dftd3_host(torch.tensor([1]), torch.zeros(2, 3), {})


def test_host_entry_repeated_calls_follow_the_buffer_contents():
    """The prepared-call fast path (same buffers again) must see new positions, a changed
    parameter set must not reuse it, and in-place edits of `numbers` are re-validated."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn

    numbers, positions = syn.drug_like_batch(9, seed=8, nmin=10, nmax=30)
    nh, ph = numbers.pin_memory(), positions.pin_memory()
    e_h = torch.empty(numbers.shape, dtype=torch.double).pin_memory()
    g_h = torch.empty(*numbers.shape, 3, dtype=torch.double).pin_memory()
    ref = _ref(torch.double)
    par = {"s8": torch.tensor(0.7898, dtype=torch.double), "a1": torch.tensor(0.4948, dtype=torch.double),
           "a2": torch.tensor(5.7308, dtype=torch.double)}

    def device_result(p):
        x = ph.to(DEV).requires_grad_(True)
        e = d3.dftd3(nh.to(DEV), x, p, ref=ref)
        (g,) = torch.autograd.grad(e.sum(), x)
        return e.detach().cpu().numpy(), g.cpu().numpy()

    for step in range(3):
        if step:
            ph.mul_(1.01)                                   # same buffer, new geometry
        d3.dftd3_host(nh, ph, par, ref=ref, energy_out=e_h, grad_out=g_h)
        e_d, g_d = device_result(par)
        assert np.array_equal(e_h.numpy(), e_d) and np.array_equal(g_h.numpy(), g_d)
    par2 = dict(par, s8=torch.tensor(1.2576, dtype=torch.double))
    d3.dftd3_host(nh, ph, par2, ref=ref, energy_out=e_h, grad_out=g_h)
    e_d, g_d = device_result(par2)
    assert np.array_equal(e_h.numpy(), e_d) and np.array_equal(g_h.numpy(), g_d)
    nh[0, 0] = 110                                          # in-place edit bumps the version: re-validated
    with pytest.raises(ValueError):
        d3.dftd3_host(nh, ph, par2, ref=ref, energy_out=e_h, grad_out=g_h)


def test_host_entry_zero_copy_large_structures():
    """Structures of 500-600 atoms: the staged positions no longer fit the reduction tiles
    (other shared-memory path of the zero-copy form); pinned zero-copy == device path bitwise."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn

    numbers, positions = syn.drug_like_batch(3, nmin=500, nmax=600, seed=12)
    par = {"s8": torch.tensor(0.7898, dtype=torch.double), "a1": torch.tensor(0.4948, dtype=torch.double),
           "a2": torch.tensor(5.7308, dtype=torch.double)}
    ref = _ref(torch.double)
    pipe = d3.HostPipeline()
    e, g = pipe.run(numbers.pin_memory(), positions.pin_memory(), par, ref=ref)
    assert pipe.last_mode == "zero-copy"
    x = positions.to(DEV).requires_grad_(True)
    ed = d3.dftd3(numbers.to(DEV), x, par, ref=ref)
    (gd,) = torch.autograd.grad(ed.sum(), x)
    assert np.array_equal(e.numpy(), ed.detach().cpu().numpy())
    assert np.array_equal(g.numpy(), gd.cpu().numpy())
    pad = (numbers == 0).numpy()
    assert pad.any() and (e.numpy()[pad] == 0).all() and (g.numpy()[pad] == 0).all()
    pipe.close()
