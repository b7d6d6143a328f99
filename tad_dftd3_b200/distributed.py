"""
i-row sharded evaluation of ONE large structure over several GPUs
(BASELINE config 4).

One process per GPU (torchrun), NCCL through `torch.distributed`.  Every rank
holds the whole structure (x, Z: 1.2 MB for 50 k atoms), builds the same
`SystemPlan`, and owns a contiguous, tile-aligned range of rows of the sorted
atom list.  The sweeps are "full-row", so a rank produces complete CN_i, E_i,
dL/dCN_i and dL/dx_i for its rows and nothing has to be reduced; the only
exchange steps are the ones the data dependencies force:

    sweep_cn(rows)      -> all-gather CN            (N reals)
    weights()              per-atom, redundantly on every rank (cheaper than gathering w)
    sweep_pair(rows)    -> all-gather dL/dCN        (N reals)
    sweep_atm(rows)        (s9 != 0) three-body term of the centres this rank owns, ADDED to
                           E, dL/dx, dL/dCN of all three atoms -- summed by the same exchanges
    sweep_cnbwd(rows)   -> all-gather E, dL/dx      (4 N reals)

The gathers are written as all-reduce(SUM) over arrays that are zero outside
the rank's rows: row ranges differ by a tile between ranks, the messages are
<= 1.6 MB, i.e. latency-bound, and all-reduce is the one collective that every
backend (NCCL here, gloo in the CPU tests) supports for ragged ownership.

`run_sharded` is independent of where the sweeps run: anything with the
`SystemRunner` interface works, which is how the world_size-2 gloo tests drive
it on CPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
from torch import Tensor

from . import defaults
from .system import TILE, SystemPlan, SystemRunner, row_partition

__all__ = ["run_sharded", "dftd3_sharded"]


LOCAL = False   # `group=LOCAL`: this process alone, even inside an initialised process group


def _world(group):
    if group is LOCAL or not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _equal_rows(npad: int, world: int) -> bool:
    return world >= 1 and (npad // TILE) % world == 0


def _exchange(t: Tensor, rows, group, mode: str) -> None:
    """Complete a per-atom array on every rank.

    mode "gather": every rank computed complete values for ITS rows (full-row sweeps, equal
    contiguous row ranges): one in-place all-gather -- each rank sends N / world values.
    mode "sum": contributions to any row may come from any rank (pairs-once sweeps, the
    three-body term, ragged row ranges): all-reduce over arrays that are zero where a rank
    did not contribute."""
    if _world(group)[1] <= 1:
        return
    if mode == "gather":
        flat = t.reshape(-1)
        per = flat.numel() // t.shape[0]
        dist.all_gather_into_tensor(flat, flat[rows[0] * per: rows[1] * per], group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def run_sharded(runner, want_grad: bool, group=None, zero: bool = True, mark=None):
    """Drive the sweeps of `runner` on this rank's rows and exchange the
    per-atom arrays.  Returns the rank's row range; afterwards `runner.energy`
    (and `runner.grad`) are complete on every rank.

    Exchange steps, exactly where the data dependencies force them: CN after the CN sweep,
    dL/dCN after the pair (and three-body) sweep, E and dL/dx at the end.  Full-row sweeps on
    equal row ranges use in-place all-gathers; everything else sums (see `_exchange`)."""
    rank, world = _world(group)
    part = getattr(runner, "partition", None)
    rows = part(rank, world) if part is not None else row_partition(runner.plan.npad, world)[rank]
    npad = runner.plan.npad
    contiguous = world > 1 and _equal_rows(npad, world) and rows == row_partition(npad, world)[rank] \
        and not getattr(runner, "symmetric", False)
    atm_on = getattr(runner, "sweep_atm", None) is not None and getattr(runner.par, "s9", 0.0) != 0.0
    mode_cn = "gather" if contiguous else "sum"
    mode_out = "gather" if (contiguous and not atm_on) else "sum"
    mark = mark or (lambda name: None)          # phase boundaries (SystemEngine.profile_step)
    if zero:
        for t in (runner.cn, runner.decn, runner.energy, runner.grad):
            t.zero_()
    mark("zero")
    runner.sweep_cn(rows)
    mark("sweep_cn")
    _exchange(runner.cn, rows, group, mode_cn)
    mark("exchange_cn")
    runner.weights()
    mark("weights")
    runner.sweep_pair(want_grad, rows)
    mark("sweep_pair")
    atm = getattr(runner, "sweep_atm", None)
    if atm is not None:
        apart = getattr(runner, "atm_partition", None)
        # adds to energy / grad / dL/dCN of every atom of a triple (no-op when s9 == 0)
        atm(want_grad, apart(rank, world) if apart is not None else rows)
        mark("sweep_atm")
    if want_grad:
        _exchange(runner.decn, rows, group, mode_out)
        mark("exchange_decn")
        runner.sweep_cnbwd(rows)
        mark("sweep_cnbwd")
    # E and dL/dx: one message when they sit next to each other in the runner's work buffer and are summed
    both = getattr(runner, "energy_grad", None) if (want_grad and mode_out == "sum") else None
    if both is not None:
        _exchange(both, rows, group, "sum")
    else:
        if want_grad:
            _exchange(runner.grad, rows, group, mode_out)
        _exchange(runner.energy, rows, group, mode_out)
    mark("exchange_out")
    return rows


class _ShardedD3(torch.autograd.Function):
    """Energies of one structure with the rows sharded over `group`, through the cached
    `SystemEngine` of the structure (plan, work arrays and the CUDA graph of the sweep
    sequence are reused while the same `numbers` tensor keeps coming back)."""

    @staticmethod
    def forward(ctx, numbers, positions, rcov, r4r2, refcn, refc6, par, group, tables=(False, False), rvdw=None):
        from .engine import engine_for

        # When the positions require grad the sweeps run with the gradient right away
        # (unit cotangent): `energy.sum().backward()` is then one multiply, as in
        # ops.D3Energy; any other cotangent repeats the sweeps in backward.
        spec = bool(positions.requires_grad)
        eng = engine_for(numbers, positions, rcov, r4r2, refcn, refc6, par, tables=tables, rvdw=rvdw, group=group,
                         want_grad=spec)
        e, gr = eng.step(positions)
        ctx.save_for_backward(numbers, positions, rcov, r4r2, refcn, refc6)
        ctx.par, ctx.group, ctx.tables, ctx.rvdw = par, group, tables, rvdw
        ctx.grad1 = gr.clone() if spec else None
        return e.clone()

    @staticmethod
    def backward(ctx, gE):
        from .engine import engine_for

        numbers, positions, rcov, r4r2, refcn, refc6 = ctx.saved_tensors
        uniform = ctx.grad1 is not None and gE.numel() > 1 and all(s == 0 for s in gE.stride())
        if _world(ctx.group)[1] > 1:
            # every rank must take the same branch (the general path issues collectives): agree on it
            flag = torch.tensor([1 if uniform else 0], dtype=torch.int32, device=positions.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=ctx.group)
            uniform = bool(int(flag))
        if uniform:
            return (None, ctx.grad1 * gE.reshape(-1)[0]) + (None,) * 8
        eng = engine_for(numbers, positions, rcov, r4r2, refcn, refc6, ctx.par, tables=ctx.tables, rvdw=ctx.rvdw,
                         group=ctx.group, want_grad=True, general_g=True)
        eng.set_cotangent(gE)
        _, gr = eng.step(positions)
        return None, gr.clone(), None, None, None, None, None, None, None, None


def dftd3_sharded(numbers: Tensor, positions: Tensor, param: dict, *, ref=None, rcov=None,
                  r4r2=None, cutoff=None, group=None) -> Tensor:
    """
    `dftd3()` for one structure `(N,)`, rows sharded over the ranks of `group`
    (two-body + CN, plus the ATM term when `param["s9"] != 0`; every rank passes
    the same inputs and receives the full atom-resolved energy; `.backward()`
    gives the full gradient on every rank).
    """
    from . import _lib
    from .disp import _check_numbers, _default_reference, _scalar

    if numbers.dim() != 1 or positions.shape != (numbers.shape[0], 3):
        raise ValueError("dftd3_sharded expects one structure: numbers (N,), positions (N, 3)")
    atm = "s9" in param and bool(param["s9"] != 0.0)   # reference disp.py:234
    _check_numbers(numbers)
    dt, dev = positions.dtype, positions.device
    if ref is None:
        ref = _default_reference(dt, dev)
    refcn, refc6, _ = ref._prepared(dev, dt)
    from .disp import _element_table

    tables = (rcov is None, r4r2 is None)
    if rcov is None:
        rcov = _element_table("COV_D3", dt, dev)
    if r4r2 is None:
        r4r2 = _element_table("R4R2", dt, dev)
    par = _lib.Params()
    par.s6 = _scalar(param.get("s6", defaults.S6))
    par.s8 = _scalar(param.get("s8", defaults.S8))
    par.a1 = _scalar(param.get("a1", defaults.A1))
    par.a2 = _scalar(param.get("a2", defaults.A2))
    par.s9 = _scalar(param["s9"]) if atm else 0.0
    par.rs9 = float(torch.tensor(defaults.RS9))       # float32(4/3), reference disp.py:312
    par.alp = _scalar(param.get("alp", defaults.ALP))
    rvdw = _element_table("VDW_PAIRWISE", dt, dev) if atm else None
    par.cutoff = defaults.D3_DISP_CUTOFF if cutoff is None else _scalar(cutoff)
    par.cn_cutoff, par.kcn = defaults.D3_CN_CUTOFF, defaults.D3_KCN
    return _ShardedD3.apply(numbers, positions, rcov, r4r2, refcn, refc6, par, group, tables, rvdw)
