"""
Persistent evaluator of ONE large structure (tiled path), optionally with its rows
sharded over the ranks of a process group (BASELINE config 4 / 5a).

`dftd3()` on a structure beyond the one-CTA limit used to rebuild the plan (Morton
sort, element bookkeeping, a dozen allocations), the work arrays and the C structs
on every call, and to issue ~12 launches plus four blocking collectives from Python:
about 1 ms of fixed cost around 0.5 ms of sweeps per rank at 8 GPUs.  MD and geometry
optimisation call the path thousands of times with the SAME numbers, so all of that
is hoisted into an object that lives as long as the structure:

    eng = SystemEngine(numbers, positions, rcov, r4r2, refcn, refc6, par, group=...)
    energy, grad = eng.step(positions)          # tensors owned by the engine

* the plan (sorted order, padding to `world * 128` rows so that every rank owns the
  same number of row tiles, element tables), every work array, the `d3b200_system`
  struct and the row partition are built once;
* one step = pack kernel (positions -> sorted AoS + sub-tile boxes), one memset, the
  sweeps with the exchanges in between (`distributed.run_sharded`), one scatter
  back to the caller's atom order.  The whole sequence, collectives included, is
  captured in a CUDA graph after two eager steps and replayed afterwards: one
  `cudaGraphLaunch` per step instead of ~20 launches and their Python;
* full-row sweeps on equal row ranges exchange by in-place all-gather (each rank sends
  N / world values); pairs-once sweeps and the three-body term sum (all-reduce).

Results are identical to the un-cached path: the same kernels run on the same arrays
in the same order (tests/test_cuda_parity.py::test_engine_*).
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from typing import Optional

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib
from .distributed import _world, run_sharded
from .system import SystemPlan, SystemRunner

__all__ = ["SystemEngine", "engine_for", "shutdown"]


class SystemEngine:
    """See the module docstring.  `rcov` / `r4r2` per atom `[N]`, or per-element tables with
    `tables=(True, True)`.  `par`: `_lib.Params`.  `rvdw`: `[104, 104]` table (s9 != 0 only)."""

    def __init__(self, numbers: Tensor, positions: Tensor, rcov: Tensor, r4r2: Tensor, refcn: Tensor,
                 refc6: Tensor, par: "_lib.Params", *, tables=(False, False), rvdw: Optional[Tensor] = None,
                 group=None, graph: Optional[bool] = None, want_grad: bool = True, general_g: bool = False):
        if positions.device.type != "cuda":
            raise RuntimeError("SystemEngine runs on CUDA devices only (tad_dftd3_b200 has no CPU path)")
        self.group = group
        self.rank, self.world = _world(group)
        self.device, self.dtype = positions.device, positions.dtype
        self.want_grad = want_grad
        self.natoms = int(numbers.shape[0])
        self.par = par
        dev, dt = self.device, self.dtype
        self.plan = SystemPlan(numbers, positions, rcov, r4r2, par.kcn, *tables, pad_tiles=self.world)
        # writable cotangent column: the runner's aux array is private to this engine
        ones = torch.ones(self.natoms, dtype=dt, device=dev)
        self.runner = SystemRunner(self.plan, refcn, refc6, par, ones, self.world, rvdw)
        self._keep = (numbers, rcov, r4r2, refcn, refc6, rvdw)
        # graph inputs / outputs at fixed addresses
        self.x_in = torch.empty(self.natoms, 3, dtype=dt, device=dev)
        self.g_in = ones
        self.energy = torch.zeros(self.natoms, dtype=dt, device=dev)   # ghosts stay exactly zero
        self.grad = torch.zeros(self.natoms, 3, dtype=dt, device=dev)
        self._unit_g = not general_g
        self._numbers_host = None
        if graph is None:
            graph = os.environ.get("TAD_DFTD3_B200_NO_GRAPH") != "1"
        self.use_graph = bool(graph)
        self._graph = None
        self._eager_steps = 0
        self.steps = 0

    # ------------------------------------------------------------------
    def _sequence(self, mark=None):
        plan, run = self.plan, self.runner
        plan.repack(self.x_in)
        if not self._unit_g:
            run.auxt[: plan.n, 1] = self.g_in[plan.perm]
        run.buf[: 6 * plan.npad].zero_()            # cn, decn, energy, grad
        if mark:
            mark("pack")
        run_sharded(run, self.want_grad, self.group, zero=False, mark=mark)
        self.energy.index_copy_(0, plan.perm, run.energy[: plan.n])
        if self.want_grad:
            self.grad.index_copy_(0, plan.perm, run.grad[: plan.n])
        if mark:
            mark("scatter")

    def profile_step(self, positions: Tensor) -> dict:
        """One EAGER step with CUDA events at the phase boundaries: {phase: ms} (diagnostics for bench.py;
        collectives included, so every rank of the group must call it)."""
        self.x_in.copy_(positions.detach())
        stream = torch.cuda.current_stream(self.device)
        names, events = [], [torch.cuda.Event(enable_timing=True)]
        events[0].record(stream)

        def mark(name):
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            names.append(name)
            events.append(ev)

        with torch.cuda.device(self.device):
            self._sequence(mark)
        torch.cuda.synchronize(self.device)
        out = {n: events[k].elapsed_time(events[k + 1]) for k, n in enumerate(names)}
        out["total"] = events[0].elapsed_time(events[-1])
        return out

    def _capture(self):
        stream = torch.cuda.current_stream(self.device)
        side = torch.cuda.Stream(self.device)
        side.wait_stream(stream)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side, capture_error_mode="thread_local"):
                self._sequence()
        stream.wait_stream(side)
        self._graph = g

    def set_cotangent(self, g: Optional[Tensor]) -> None:
        """Cotangent of the per-atom energies (L = sum g_i E_i) of an engine built with
        `general_g=True`; None = ones."""
        if self._unit_g:
            if g is not None:
                raise ValueError("engine was built for the unit cotangent (general_g=False)")
            return
        if g is None:
            self.g_in.fill_(1.0)
        else:
            self.g_in.copy_(g.detach().to(self.dtype).reshape(-1))

    def step(self, positions: Optional[Tensor] = None):
        """One evaluation at `positions` (same structure; None: `x_in` was filled by the caller).
        Returns `(energy [N], grad [N, 3])`, overwritten by the next step."""
        if positions is not None:
            self.x_in.copy_(positions.detach(), non_blocking=True)
        with torch.cuda.device(self.device):
            if self._graph is not None:
                self._graph.replay()
                _lib.launch_count += 1
            else:
                self._sequence()
                self._eager_steps += 1
                if self.use_graph and self._eager_steps >= 2:
                    try:
                        self._capture()
                    except RuntimeError:
                        # capture is an optimisation: fall back to eager launches of the same kernels
                        self.use_graph, self._graph = False, None
        self.steps += 1
        return self.energy, (self.grad if self.want_grad else None)


    def step_host(self, numbers: Tensor, positions: Tensor, energy_out: Tensor, grad_out: Optional[Tensor] = None,
                  sync: bool = True):
        """One step for callers that keep the structure in HOST memory (page-locked tensors):
        positions are copied to the device, the results back, all on the current stream.
        `numbers` must be the structure the engine was built for (checked on the host, 50 kB compare)."""
        if self._numbers_host is None:
            self._numbers_host = self._keep[0].detach().cpu()
        if numbers.shape != self._numbers_host.shape or not torch.equal(numbers, self._numbers_host):
            raise ValueError("step_host: `numbers` differs from the structure this engine was built for")
        self.x_in.copy_(positions, non_blocking=True)
        e, g = self.step(None)
        energy_out.copy_(e, non_blocking=True)
        if grad_out is not None and g is not None:
            grad_out.copy_(g, non_blocking=True)
        if sync:
            torch.cuda.current_stream(self.device).synchronize()
        return energy_out, grad_out


    def release_graph(self) -> None:
        """Drop the captured CUDA graph (the next steps run eagerly and re-capture).  A graph that
        holds captured NCCL work must be destroyed BEFORE its communicator:
        `dist.destroy_process_group()` otherwise waits forever (observed with NCCL 2.28)."""
        if self._graph is not None:
            torch.cuda.synchronize(self.device)
            self._graph = None
        self._eager_steps = 0


_engines: dict = {}


def shutdown() -> None:
    """Release every cached engine (graphs first) and the cached ATM workspaces (neighbour scratch + pair table:
    32 N^2 bytes per device and stream).  Call before `dist.destroy_process_group()`."""
    import gc

    from . import system as _system

    for _, eng in list(_engines.values()):
        eng.release_graph()
    _engines.clear()
    _system._atm_work_cache.clear()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def engine_for(numbers: Tensor, positions: Tensor, rcov: Tensor, r4r2: Tensor, refcn: Tensor, refc6: Tensor,
               par: "_lib.Params", *, tables=(False, False), rvdw=None, group=None, want_grad: bool = True,
               general_g: bool = False) -> SystemEngine:
    """Engine cache for call sites that see the same `numbers` tensor again and again
    (`dftd3()` on a large structure, `dftd3_sharded()`): keyed on the identity and version of
    the structure-defining tensors and on the scalar parameters."""
    pkey = tuple(getattr(par, k) for k, _ in par._fields_)
    key = (id(numbers), numbers._version, positions.dtype, str(positions.device), id(rcov), id(r4r2), id(refcn),
           id(refc6), id(rvdw), tables, pkey, id(group), want_grad, general_g)
    hit = _engines.get(key)
    if hit is not None and hit[0]() is numbers:
        return hit[1]
    eng = SystemEngine(numbers, positions, rcov, r4r2, refcn, refc6, par, tables=tables, rvdw=rvdw, group=group,
                       want_grad=want_grad, general_g=general_g)
    if len(_engines) >= 8:
        _engines.pop(next(iter(_engines)))
    try:
        _engines[key] = (weakref.ref(numbers), eng)
    except TypeError:  # pragma: no cover
        pass
    return eng
