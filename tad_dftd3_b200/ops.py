"""
Host side of the drop-in boundary: `torch.autograd.Function`s over the C ABI.

Three levels, each a new-style Function (`forward` without ctx +
`setup_context`) with a manual `vmap` rule, mirroring the reference's only
custom op (`AtomicC6_V2`, model/c6.py:401-480):

    D3Energy : (numbers, x, p, tables...)          -> E[..., N]      d3b200_batch_energy_*
    D3Vjp    : (numbers, x, p, g, tables...)       -> dL/dx, dL/dp   d3b200_batch_vjp_*
    D3Hvp    : (numbers, x, p, g, vx, vp, tables...) -> tangents     d3b200_batch_hvp_*

`D3Energy.backward` calls `D3Vjp`, `D3Vjp.backward` calls `D3Hvp`, so
`.backward()`, `autograd.grad`, `gradcheck`, `gradgradcheck`,
`torch.func.jacrev`, nested `jacrev` (examples/hessian.py:84-94) and
`vmap(jacrev(jacrev))` all run on the CUDA kernels: torch autograd never
differentiates through the pair sums.  Third derivatives are not provided.

`p` is the 5-vector (s6, s8, a1, a2, s9).  Every tensor is passed as a
Function argument (functorch unwraps arguments level by level; tensors hidden
inside Python objects would stay wrapped).  All per-structure tensors may carry
arbitrary leading batch dimensions; raw pointers are only taken inside
`forward`, the `vmap` rules only move/expand dimensions and re-dispatch.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import _lib

__all__ = ["Settings", "D3Energy", "D3Vjp", "D3Hvp", "d3_energy"]

NPAR = 5


@dataclass(frozen=True)
class Settings:
    """Plain-Python (non-tensor) settings of one `dftd3` call; static under vmap."""

    rcov_per_element: bool
    r4r2_per_element: bool
    rvdw_mode: int = 0  # 0 none, 1 element-pair table, 2 dense [..., N, N]
    cutoff: float = 50.0
    cn_cutoff: float = 25.0
    kcn: float = 16.0
    rs9: float = 1.3333333730697632  # float32(4/3), reference disp.py:312
    alp: float = 14.0
    counting: int = 0   # d3b200_params.counting: 0 exp_count, 1 erf_count
    damping: int = 0    # d3b200_params.damping: 0 rational (p[2], p[3] = a1, a2), 1 zero (p[2], p[3] = rs6, rs8)


def _sfx(dtype: torch.dtype) -> str:
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise TypeError(f"tad_dftd3_b200 supports float32 and float64 positions, got {dtype}")


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _params(vals, st: Settings) -> _lib.Params:
    p = _lib.Params()
    p.s6, p.s8, p.a1, p.a2, p.s9 = (float(v) for v in vals)
    p.rs9, p.alp = st.rs9, st.alp
    p.cutoff, p.cn_cutoff, p.kcn = st.cutoff, st.cn_cutoff, st.kcn
    p.counting, p.damping = int(st.counting), int(st.damping)
    return p


def _pvals(p: Tensor):
    return p.detach().reshape(-1).tolist()


def _as(t: Tensor, dtype, device) -> Tensor:
    """detach + move + make contiguous, skipping every step that is a no-op
    (each skipped torch call saves a few microseconds of host time per launch)."""
    if t.requires_grad:
        t = t.detach()
    if t.dtype != dtype or t.device != device:
        t = t.to(device=device, dtype=dtype)
    return t if t.is_contiguous() else t.contiguous()


_order_cache: dict = {}   # id(numbers) -> [weakref, version, calls seen, order tensor or None]
ORDER_MIN_BATCH = 1184    # two waves of CTAs on 148 SMs: below that the schedule does not matter


def batch_order(z: Tensor) -> Tensor:
    """Scheduling hint of the batch kernels (`d3b200_model.batch_order`): structures sorted by
    their number of real atoms, largest first, so that the last wave of CTAs consists of the
    cheapest structures.  `z` is the flattened `[nflat, N]` numbers tensor (any device)."""
    counts = (z != 0).sum(-1)
    return torch.argsort(counts, descending=True, stable=True).to(torch.int32)


def _cached_order(numbers: Tensor, z: Tensor):
    """Order hint for a numbers tensor that keeps coming back (MD / optimisation loops, repeated
    evaluations): computed on the second call with the same tensor, reused afterwards; a one-off
    call never pays for it."""
    import weakref

    if z.shape[0] < ORDER_MIN_BATCH or _is_functorch(numbers) or os.environ.get("D3B200_NO_ORDER_HINT") == "1":
        return None
    key = id(numbers)
    hit = _order_cache.get(key)
    if hit is None or hit[0]() is not numbers or hit[1] != numbers._version:
        if len(_order_cache) > 64:
            _order_cache.clear()
        try:
            _order_cache[key] = [weakref.ref(numbers), numbers._version, 1, None]
        except TypeError:  # pragma: no cover
            pass
        return None
    hit[2] += 1
    if hit[3] is None:
        hit[3] = batch_order(z)
    return hit[3] if hit[3].shape[0] == z.shape[0] else None


class _Call:
    """Flattened, contiguous device views of one call + the C structs."""

    def __init__(self, numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st: Settings, order: int,
                 pvals=None):
        if positions.device.type != "cuda":
            raise RuntimeError(
                "tad_dftd3_b200 runs on CUDA devices only (there is no CPU path); "
                f"got positions on {positions.device}"
            )
        self.lead = lead = positions.shape[:-2]
        self.n = n = positions.shape[-2]
        self.dtype, self.device = dt, dev = positions.dtype, positions.device
        nflat = 1
        for d in lead:
            nflat *= int(d)
        x = _as(positions, dt, dev)
        self.x = x if x.dim() == 3 else x.reshape(nflat, n, 3)     # (explicit sizes: n or nflat may be 0)
        z = numbers
        if z.shape != positions.shape[:-1]:
            z = z.expand(*lead, n)
        z = _as(z, torch.int64, dev)
        self.z = z if z.dim() == 2 else z.reshape(nflat, n)
        self.nflat = nflat
        self.st = st

        def atomwise(t, trailing=()):
            if t.shape != (*lead, n, *trailing):
                t = t.expand(*lead, n, *trailing)
            return _as(t, dt, dev).reshape(nflat, n, *trailing)

        self.rcov = _as(rcov, dt, dev) if st.rcov_per_element else atomwise(rcov)
        self.r4r2 = _as(r4r2, dt, dev) if st.r4r2_per_element else atomwise(r4r2)
        self.rvdw = None
        if st.rvdw_mode == 1:
            self.rvdw = _as(rvdw, dt, dev)
        elif st.rvdw_mode == 2:
            self.rvdw = atomwise(rvdw, (n,))
        self.refcn = _as(refcn, dt, dev)
        self.refc6 = _as(refc6, dt, dev)
        self.pvals = _pvals(p) if pvals is None else pvals
        self.par = _params(self.pvals, st)
        m = _lib.Model()
        m.rcov, m.r4r2 = self.rcov.data_ptr(), self.r4r2.data_ptr()
        m.rcov_per_element, m.r4r2_per_element = int(st.rcov_per_element), int(st.r4r2_per_element)
        m.refcn, m.refc6 = self.refcn.data_ptr(), self.refc6.data_ptr()
        m.rvdw = self.rvdw.data_ptr() if self.rvdw is not None else None
        m.rvdw_mode = int(st.rvdw_mode)
        self.order = _cached_order(numbers, self.z) if order == 0 else None
        m.batch_order = self.order.data_ptr() if self.order is not None else None
        self.model = m
        self.sfx = _sfx(dt)
        # ATM scratch (sqrt|C6| matrix per structure)
        self.work = None
        self.work_bytes = 0
        if self.par.s9 != 0.0 and nflat and n and n <= _fused_limit(self.sfx, order):
            fn = getattr(_lib.lib(), f"d3b200_batch_workspace_bytes_{self.sfx}")
            self.work_bytes = int(fn(nflat, n, order, 1))
            self.work = torch.empty(self.work_bytes, dtype=torch.uint8, device=dev)

    def per_atom(self, t: Tensor, trailing=()):
        return t.detach().to(device=self.device, dtype=self.dtype).expand(*self.lead, self.n, *trailing) \
            .reshape(self.nflat, self.n, *trailing).contiguous()

    def stream(self):
        _lib.launch_count += 1  # every entry point issues exactly one kernel launch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def new(self, *shape, zero=False):
        f = torch.zeros if zero else torch.empty
        return f(*shape, dtype=self.dtype, device=self.device)


_limits: dict = {}


def _fused_limit(sfx: str, order: int) -> int:
    """Largest structure the one-CTA-per-structure kernels take."""
    key = (sfx, order)
    limit = _limits.get(key)
    if limit is None:
        limit = _limits[key] = getattr(_lib.lib(), f"d3b200_batch_max_atoms_{sfx}")(order)
    return limit


def _tiled(c: "_Call", order: int) -> bool:
    """Structures beyond the one-CTA limit go to the tiled `d3b200_system_*` sweeps."""
    if not (c.nflat and c.n):
        return False
    default_model = c.st.counting == 0 and c.st.damping == 0
    if os.environ.get("TAD_DFTD3_B200_TILED") == "1" and default_model:
        return True
    if c.n > _fused_limit(c.sfx, order) and not default_model:
        raise NotImplementedError(
            "erf_count / zero_damping run in the general fused kernel, which takes structures of up to "
            f"{_fused_limit(c.sfx, 1)} atoms; the tiled large-structure sweeps implement the default "
            "D3(BJ) model functions (exp_count, rational_damping) only")
    return c.n > _fused_limit(c.sfx, order)


def _tiled_run(c: "_Call", g, want_grad: bool, numbers=None, want_p: bool = False):
    if c.par.s9 != 0.0 and c.st.rvdw_mode != 1:
        raise NotImplementedError(
            "the tiled path evaluates the ATM term from the [104,104] element-pair "
            "rvdw table only (pass rvdw=None); a dense per-pair rvdw is limited to "
            "structures that fit the fused kernel"
        )
    from .distributed import LOCAL
    from .engine import engine_for

    es, gs, ps = [], [], []
    for b in range(c.nflat):
        z = numbers if (c.nflat == 1 and numbers.dim() == 1 and numbers.dtype == torch.int64) else c.z[b]
        rcov = c.rcov if c.st.rcov_per_element else c.rcov[b]
        r4r2 = c.r4r2 if c.st.r4r2_per_element else c.r4r2[b]
        # plan, work arrays and the CUDA graph of the sweep sequence live in a cached engine while
        # the same numbers tensor keeps coming back (MD / optimisation loops)
        eng = engine_for(z, c.x[b], rcov, r4r2, c.refcn, c.refc6, c.par,
                         tables=(c.st.rcov_per_element, c.st.r4r2_per_element), rvdw=c.rvdw,
                         group=LOCAL, want_grad=want_grad, general_g=g is not None)
        if g is not None:
            eng.set_cotangent(g[b])
        e, gr = eng.step(c.x[b])
        es.append(e.clone())
        if want_grad:
            gs.append(gr.clone())
        if want_p:
            ps.append(eng.runner.param_gradient())
    if want_p:
        return torch.stack(es), (torch.stack(gs) if want_grad else None), torch.stack(ps)
    return torch.stack(es), (torch.stack(gs) if want_grad else None)


def _run_energy(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st):
    c = _Call(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st, 0)
    if _tiled(c, 0):
        return _tiled_run(c, None, False, numbers)[0].reshape(*c.lead, c.n)
    out = c.new(c.nflat, c.n)
    if c.nflat and c.n:
        fn = getattr(_lib.lib(), f"d3b200_batch_energy_{c.sfx}")
        with torch.cuda.device(c.device):
            _lib.check(fn(c.nflat, c.n, _ptr(c.z), _ptr(c.x), C.byref(c.model), C.byref(c.par),
                          _ptr(out), _ptr(c.work), C.c_size_t(c.work_bytes), c.stream()))
    return out.reshape(*c.lead, c.n)


def _run_vjp(numbers, positions, p, g, rcov, r4r2, rvdw, refcn, refc6, st, want_p: bool):
    c = _Call(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st, 0)
    if _tiled(c, 0):
        if want_p:
            _, gr, pg = _tiled_run(c, c.per_atom(g), True, numbers, True)
            return gr.reshape(*c.lead, c.n, 3), pg.reshape(*c.lead, NPAR)
        _, gr = _tiled_run(c, c.per_atom(g), True, numbers)
        return gr.reshape(*c.lead, c.n, 3), c.new(c.nflat, NPAR, zero=True).reshape(*c.lead, NPAR)
    grad = c.new(c.nflat, c.n, 3)
    pgrad = c.new(c.nflat, NPAR, zero=True)
    if c.nflat and c.n:
        gg = c.per_atom(g)
        fn = getattr(_lib.lib(), f"d3b200_batch_vjp_{c.sfx}")
        with torch.cuda.device(c.device):
            _lib.check(fn(c.nflat, c.n, _ptr(c.z), _ptr(c.x), C.byref(c.model), C.byref(c.par),
                          _ptr(gg), None, _ptr(grad), _ptr(pgrad) if want_p else None,
                          _ptr(c.work), C.c_size_t(c.work_bytes), c.stream()))
    return grad.reshape(*c.lead, c.n, 3), pgrad.reshape(*c.lead, NPAR)


def _run_hvp(numbers, positions, p, g, vx, vp, rcov, r4r2, rvdw, refcn, refc6, st):
    c = _Call(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st, 1)
    if c.nflat and c.n > _fused_limit(c.sfx, 1):
        raise NotImplementedError(
            f"second derivatives (gradgradcheck, nested jacrev, hessian()) are provided by the fused kernel on dual "
            f"numbers, which takes structures of up to {_fused_limit(c.sfx, 1)} atoms ({c.n} given); the tiled "
            "large-structure sweeps provide energies, gradients and parameter gradients")
    de = c.new(c.nflat, c.n, zero=True)
    dgrad = c.new(c.nflat, c.n, 3, zero=True)
    dpgrad = c.new(c.nflat, NPAR, zero=True)
    if c.nflat and c.n:
        gg = c.per_atom(g)
        dx = c.per_atom(vx, (3,))
        dpar = _params(_pvals(vp), st)
        fn = getattr(_lib.lib(), f"d3b200_batch_hvp_{c.sfx}")
        with torch.cuda.device(c.device):
            _lib.check(fn(c.nflat, c.n, _ptr(c.z), _ptr(c.x), _ptr(dx), C.byref(c.model),
                          C.byref(c.par), C.byref(dpar), _ptr(gg), None, None, _ptr(de),
                          _ptr(dgrad), _ptr(dpgrad), _ptr(c.work), C.c_size_t(c.work_bytes),
                          c.stream()))
    return de.reshape(*c.lead, c.n), dgrad.reshape(*c.lead, c.n, 3), dpgrad.reshape(*c.lead, NPAR)


# --------------------------------------------------------------------------
# vmap helpers
# --------------------------------------------------------------------------
def _roles(names, st: Settings):
    """Role of every positional argument: 's' per-structure tensor, 'p' parameter
    vector, 'c' constant table (must be unbatched), '-' non-tensor."""
    role = {}
    for k, name in enumerate(names):
        if name in ("numbers", "positions", "g", "vx"):
            role[k] = "s"
        elif name in ("p", "vp"):
            role[k] = "p"
        elif name == "rcov":
            role[k] = "c" if st.rcov_per_element else "s"
        elif name == "r4r2":
            role[k] = "c" if st.r4r2_per_element else "s"
        elif name == "rvdw":
            role[k] = "s" if st.rvdw_mode == 2 else "c"
        elif name in ("refcn", "refc6"):
            role[k] = "c"
        else:
            role[k] = "-"
    return role


def _vmap(fn, names, info, in_dims, args):
    st = args[names.index("st")]
    role = _roles(names, st)
    batched_p = any(role[k] == "p" and isinstance(a, Tensor) and d is not None
                    for k, (a, d) in enumerate(zip(args, in_dims)))
    if batched_p:
        # one parameter set per launch: loop over the batch (parameter Hessians
        # under jacrev only; positions-only transforms never come here)
        outs = []
        for c in range(info.batch_size):
            sl = [a.select(d, c) if (isinstance(a, Tensor) and d is not None) else a
                  for a, d in zip(args, in_dims)]
            outs.append(fn(*sl))
        if isinstance(outs[0], tuple):
            return tuple(torch.stack(o) for o in zip(*outs)), tuple(0 for _ in outs[0])
        return torch.stack(outs), 0
    new = []
    for k, (a, d) in enumerate(zip(args, in_dims)):
        if not isinstance(a, Tensor):
            if d is not None:
                raise ValueError("non-tensor arguments must be static under vmap")
            new.append(a)
        elif role[k] == "s":
            new.append(a.unsqueeze(0).expand(info.batch_size, *a.shape) if d is None else a.movedim(d, 0))
        else:
            if d is not None:
                raise ValueError(f"`{names[k]}` must not be batched under vmap")
            new.append(a)
    out = fn(*new)
    if isinstance(out, tuple):
        return out, tuple(0 for _ in out)
    return out, 0


_E_NAMES = ("numbers", "positions", "p", "rcov", "r4r2", "rvdw", "refcn", "refc6", "st")
_V_NAMES = ("numbers", "positions", "p", "g", "rcov", "r4r2", "rvdw", "refcn", "refc6", "st", "want_p")
_H_NAMES = ("numbers", "positions", "p", "g", "vx", "vp", "rcov", "r4r2", "rvdw", "refcn", "refc6", "st")


_SPECULATE = os.environ.get("TAD_DFTD3_B200_SPECULATE", "1") == "1"


def _is_functorch(x) -> bool:
    f = torch._C._functorch
    return isinstance(x, Tensor) and bool(f.is_batchedtensor(x) or f.is_gradtrackingtensor(x))


class D3Energy(torch.autograd.Function):
    """Atom-resolved D3 energies (body of the reference's dftd3(), disp.py:150-165).

    When the positions require grad, forward already runs the fused
    energy + gradient kernel with a unit cotangent and keeps dE_total/dx; a
    backward pass whose cotangent is a broadcast scalar (`energy.sum().backward()`,
    every stride 0) then costs one elementwise multiply instead of a second
    sweep over the pairs.  Any other cotangent, `create_graph=True` and
    functorch transforms take the general D3Vjp path.
    """

    generate_vmap_rule = False

    @staticmethod
    def forward(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st):
        if (positions.requires_grad and not p.requires_grad and _SPECULATE
                and not _is_functorch(positions)):
            pvals = _pvals(p)
            c = _Call(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st, 0, pvals)
            if c.nflat and c.n and _tiled(c, 0):
                # large structures: the sweeps with gradient give the energies as well
                e, gr = _tiled_run(c, None, True, numbers)
                return e.reshape(*c.lead, c.n), gr.reshape(*c.lead, c.n, 3)
            if pvals[4] == 0.0:
                if c.nflat and c.n:
                    energy, grad = c.new(c.nflat, c.n), c.new(c.nflat, c.n, 3)
                    fn = getattr(_lib.lib(), f"d3b200_batch_vjp_{c.sfx}")
                    with torch.cuda.device(c.device):
                        _lib.check(fn(c.nflat, c.n, _ptr(c.z), _ptr(c.x), C.byref(c.model), C.byref(c.par),
                                      None, _ptr(energy), _ptr(grad), None, None, C.c_size_t(0), c.stream()))
                    return energy.reshape(*c.lead, c.n), grad.reshape(*c.lead, c.n, 3)
        e = _run_energy(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st)
        return e, e.new_empty(0)

    @staticmethod
    def setup_context(ctx, inputs, output):
        numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st = inputs
        ctx.save_for_backward(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6)
        ctx.st = st
        ctx.grad1 = output[1] if output[1].numel() else None
        ctx.mark_non_differentiable(output[1])

    @staticmethod
    def backward(ctx, gE, _unused=None):
        numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6 = ctx.saved_tensors
        need_x, need_p = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        out = [None] * 9
        if not (need_x or need_p):
            return tuple(out)
        if (ctx.grad1 is not None and need_x and not need_p and not torch.is_grad_enabled()
                and gE.numel() > 1 and all(s == 0 for s in gE.stride()) and not _is_functorch(gE)):
            # uniform cotangent c: dL/dx = c * d(sum E)/dx, already computed in forward
            out[1] = ctx.grad1 * gE.reshape(-1)[0]
            return tuple(out)
        gx, gp = D3Vjp.apply(numbers, positions, p, gE, rcov, r4r2, rvdw, refcn, refc6, ctx.st, bool(need_p))
        if need_x:
            out[1] = gx
        if need_p:
            out[2] = gp.reshape(-1, NPAR).sum(0).to(device=p.device, dtype=p.dtype)
        return tuple(out)

    @staticmethod
    def vmap(info, in_dims, *args):
        (e, _), _dims = _vmap(lambda *a: D3Energy.apply(*a), _E_NAMES, info, in_dims, args)
        return (e, e.new_empty(info.batch_size, 0)), (0, 0)


class D3Vjp(torch.autograd.Function):
    """(x, p, g) -> (dL/dx [..., N, 3], dL/dp per structure [..., 5]), L = sum g E."""

    generate_vmap_rule = False

    @staticmethod
    def forward(numbers, positions, p, g, rcov, r4r2, rvdw, refcn, refc6, st, want_p=True):
        return _run_vjp(numbers, positions, p, g, rcov, r4r2, rvdw, refcn, refc6, st, want_p)

    @staticmethod
    def setup_context(ctx, inputs, output):
        numbers, positions, p, g, rcov, r4r2, rvdw, refcn, refc6, st, _ = inputs
        ctx.save_for_backward(numbers, positions, p, g, rcov, r4r2, rvdw, refcn, refc6)
        ctx.st = st

    @staticmethod
    def backward(ctx, vx, vp):
        numbers, positions, p, g, rcov, r4r2, rvdw, refcn, refc6 = ctx.saved_tensors
        if vx is None:
            vx = torch.zeros_like(positions)
        # The cotangent on the per-structure parameter gradients comes from the
        # sum over structures in D3Energy.backward, so it is the same 5-vector
        # for every structure: it acts as one tangent on the shared parameters.
        if vp is None:
            vp_shared = torch.zeros(NPAR, dtype=p.dtype, device=p.device)
        else:
            vp_shared = vp.reshape(-1, NPAR)[0].to(device=p.device, dtype=p.dtype)
        de, dgx, dgp = D3Hvp.apply(numbers, positions, p, g, vx, vp_shared, rcov, r4r2, rvdw,
                                   refcn, refc6, ctx.st)
        out = [None] * 11
        out[1] = dgx
        out[2] = dgp.reshape(-1, NPAR).sum(0).to(device=p.device, dtype=p.dtype)
        out[3] = de
        return tuple(out)

    @staticmethod
    def vmap(info, in_dims, *args):
        return _vmap(D3Vjp.apply, _V_NAMES, info, in_dims, args)


class D3Hvp(torch.autograd.Function):
    """Tangents of (E, dL/dx, dL/dp) along (vx, vp): second-order path."""

    generate_vmap_rule = False

    @staticmethod
    def forward(numbers, positions, p, g, vx, vp, rcov, r4r2, rvdw, refcn, refc6, st):
        return _run_hvp(numbers, positions, p, g, vx, vp, rcov, r4r2, rvdw, refcn, refc6, st)

    @staticmethod
    def setup_context(ctx, inputs, output):
        pass

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(
            "tad_dftd3_b200 provides analytic first and second derivatives; "
            "third-order derivatives of dftd3() are not implemented"
        )

    @staticmethod
    def vmap(info, in_dims, *args):
        return _vmap(D3Hvp.apply, _H_NAMES, info, in_dims, args)


def d3_energy(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st: Settings) -> Tensor:
    return D3Energy.apply(numbers, positions, p, rcov, r4r2, rvdw, refcn, refc6, st)[0]
