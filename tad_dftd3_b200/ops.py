"""
Host side of the drop-in boundary: `torch.autograd.Function`s over the C ABI.

Three levels, each a new-style Function (`forward` without ctx +
`setup_context`) with a manual `vmap` rule, mirroring the reference's only
custom op (`AtomicC6_V2`, model/c6.py:401-480):

    D3Energy : (numbers, x, p)        -> E[..., N]          d3b200_batch_energy_*
    D3Vjp    : (numbers, x, p, g)     -> dL/dx, dL/dp       d3b200_batch_vjp_*
    D3Hvp    : (numbers, x, p, g, vx, vp) -> tangents       d3b200_batch_hvp_*

`D3Energy.backward` calls `D3Vjp`, `D3Vjp.backward` calls `D3Hvp`, so
`.backward()`, `autograd.grad`, `gradcheck`, `gradgradcheck`,
`torch.func.jacrev`, nested `jacrev` (examples/hessian.py:84-94) and
`vmap(jacrev(jacrev))` all run on the CUDA kernels: torch autograd never
differentiates through the pair sums.  Third derivatives are not provided.

`p` is the 5-vector (s6, s8, a1, a2, s9).  All tensor arguments may carry
arbitrary leading batch dimensions; raw pointers are only taken inside
`forward`, the `vmap` rules only move/expand dimensions and re-dispatch.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import torch
from torch import Tensor

from . import _lib

__all__ = ["Settings", "D3Energy", "D3Vjp", "D3Hvp", "d3_energy"]

NPAR = 5


@dataclass
class Settings:
    """Non-differentiable inputs of one `dftd3` call (static under vmap)."""

    rcov: Tensor
    r4r2: Tensor
    rcov_per_element: bool
    r4r2_per_element: bool
    refcn: Tensor
    refc6: Tensor
    rvdw: Optional[Tensor] = None
    rvdw_mode: int = 0
    cutoff: float = 50.0
    cn_cutoff: float = 25.0
    kcn: float = 16.0
    rs9: float = 1.3333333730697632  # float32(4/3), reference disp.py:312
    alp: float = 14.0
    keep: list = field(default_factory=list)


def _sfx(dtype: torch.dtype) -> str:
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise TypeError(f"tad_dftd3_b200 supports float32 and float64 positions, got {dtype}")


def _ptr(t: Optional[Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _model(st: Settings, nflat: int, natoms: int, dtype, device) -> _lib.Model:
    m = _lib.Model()
    m.rcov = st.rcov.data_ptr()
    m.r4r2 = st.r4r2.data_ptr()
    m.rcov_per_element = int(st.rcov_per_element)
    m.r4r2_per_element = int(st.r4r2_per_element)
    m.refcn = st.refcn.data_ptr()
    m.refc6 = st.refc6.data_ptr()
    m.rvdw = st.rvdw.data_ptr() if st.rvdw is not None else None
    m.rvdw_mode = int(st.rvdw_mode)
    return m


def _params(vals, st: Settings) -> _lib.Params:
    p = _lib.Params()
    p.s6, p.s8, p.a1, p.a2, p.s9 = (float(v) for v in vals)
    p.rs9, p.alp = st.rs9, st.alp
    p.cutoff, p.cn_cutoff, p.kcn = st.cutoff, st.cn_cutoff, st.kcn
    return p


def _flatten(numbers: Tensor, positions: Tensor):
    if positions.device.type != "cuda":
        raise RuntimeError(
            "tad_dftd3_b200 runs on CUDA devices only (there is no CPU path); "
            f"got positions on {positions.device}"
        )
    lead = positions.shape[:-2]
    n = positions.shape[-2]
    x = positions.detach().reshape(-1, n, 3).contiguous()
    z = numbers.detach().expand(*lead, n).reshape(-1, n).contiguous()
    if z.dtype != torch.int64:
        z = z.to(torch.int64)
    return lead, n, z, x


def _per_atom(t: Tensor, st_flag: bool, lead, n, nflat):
    """Per-atom model arrays follow the batch shape of `numbers`."""
    if st_flag:
        return t
    return t.detach().expand(*lead, n).reshape(nflat, n).contiguous()


def _prep_settings(st: Settings, lead, n, nflat, dtype, device) -> Settings:
    rcov = _per_atom(st.rcov, st.rcov_per_element, lead, n, nflat)
    r4r2 = _per_atom(st.r4r2, st.r4r2_per_element, lead, n, nflat)
    rvdw = st.rvdw
    if st.rvdw_mode == 2 and rvdw is not None:
        rvdw = rvdw.detach().expand(*lead, n, n).reshape(nflat, n, n).contiguous()
    if rcov is st.rcov and r4r2 is st.r4r2 and rvdw is st.rvdw:
        return st
    return Settings(rcov, r4r2, st.rcov_per_element, st.r4r2_per_element, st.refcn, st.refc6,
                    rvdw, st.rvdw_mode, st.cutoff, st.cn_cutoff, st.kcn, st.rs9, st.alp)


def _pvals(p: Tensor):
    return p.detach().reshape(-1).tolist()


def _run_energy(numbers, positions, p, st):
    lead, n, z, x = _flatten(numbers, positions)
    nflat = z.shape[0]
    out = torch.empty(nflat, n, dtype=x.dtype, device=x.device)
    if nflat and n:
        st2 = _prep_settings(st, lead, n, nflat, x.dtype, x.device)
        model = _model(st2, nflat, n, x.dtype, x.device)
        par = _params(_pvals(p), st)
        fn = getattr(_lib.lib(), f"d3b200_batch_energy_{_sfx(x.dtype)}")
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(fn(nflat, n, _ptr(z), _ptr(x), C.byref(model), C.byref(par), _ptr(out), stream))
    return out.reshape(*lead, n)


def _run_vjp(numbers, positions, p, g, st, want_p: bool):
    lead, n, z, x = _flatten(numbers, positions)
    nflat = z.shape[0]
    grad = torch.empty(nflat, n, 3, dtype=x.dtype, device=x.device)
    pgrad = torch.zeros(nflat, NPAR, dtype=x.dtype, device=x.device)
    if nflat and n:
        gg = g.detach().to(x.dtype).expand(*lead, n).reshape(nflat, n).contiguous()
        st2 = _prep_settings(st, lead, n, nflat, x.dtype, x.device)
        model = _model(st2, nflat, n, x.dtype, x.device)
        par = _params(_pvals(p), st)
        fn = getattr(_lib.lib(), f"d3b200_batch_vjp_{_sfx(x.dtype)}")
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(fn(nflat, n, _ptr(z), _ptr(x), C.byref(model), C.byref(par), _ptr(gg), None,
                          _ptr(grad), _ptr(pgrad) if want_p else None, stream))
    return grad.reshape(*lead, n, 3), pgrad.reshape(*lead, NPAR)


def _run_hvp(numbers, positions, p, g, vx, vp, st):
    lead, n, z, x = _flatten(numbers, positions)
    nflat = z.shape[0]
    de = torch.zeros(nflat, n, dtype=x.dtype, device=x.device)
    dgrad = torch.zeros(nflat, n, 3, dtype=x.dtype, device=x.device)
    dpgrad = torch.zeros(nflat, NPAR, dtype=x.dtype, device=x.device)
    if nflat and n:
        gg = g.detach().to(x.dtype).expand(*lead, n).reshape(nflat, n).contiguous()
        dx = vx.detach().to(x.dtype).expand(*lead, n, 3).reshape(nflat, n, 3).contiguous()
        st2 = _prep_settings(st, lead, n, nflat, x.dtype, x.device)
        model = _model(st2, nflat, n, x.dtype, x.device)
        par = _params(_pvals(p), st)
        dvals = _pvals(vp)
        dpar = _params(dvals, st)
        fn = getattr(_lib.lib(), f"d3b200_batch_hvp_{_sfx(x.dtype)}")
        with torch.cuda.device(x.device):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(fn(nflat, n, _ptr(z), _ptr(x), _ptr(dx), C.byref(model), C.byref(par),
                          C.byref(dpar), _ptr(gg), None, None, _ptr(de), _ptr(dgrad), _ptr(dpgrad),
                          stream))
    return de.reshape(*lead, n), dgrad.reshape(*lead, n, 3), dpgrad.reshape(*lead, NPAR)


# --------------------------------------------------------------------------
# vmap helper: move batch dims to the front / expand unbatched arguments
# --------------------------------------------------------------------------
def _batched_params(in_dims, args, per_structure) -> bool:
    return any(isinstance(a, Tensor) and d is not None and k not in per_structure
               for k, (a, d) in enumerate(zip(args, in_dims)))


def _loop(fn, info, in_dims, args):
    """Fallback when a parameter vector is batched: one launch per element."""
    outs = []
    for c in range(info.batch_size):
        sl = [a.select(d, c) if (isinstance(a, Tensor) and d is not None) else a
              for a, d in zip(args, in_dims)]
        outs.append(fn(*sl))
    if isinstance(outs[0], tuple):
        return tuple(torch.stack(o) for o in zip(*outs)), tuple(0 for _ in outs[0])
    return torch.stack(outs), 0


def _front(info, in_dims, args, per_structure):
    """Returns args with a leading dim of size info.batch_size on every
    per-structure tensor (index in `per_structure`); other tensors must be
    unbatched."""
    out = []
    for k, (a, d) in enumerate(zip(args, in_dims)):
        if not isinstance(a, Tensor):
            if d is not None:
                raise ValueError("non-tensor arguments must be static under vmap")
            out.append(a)
            continue
        if k in per_structure:
            if d is None:
                a = a.unsqueeze(0).expand(info.batch_size, *a.shape)
            else:
                a = a.movedim(d, 0)
            out.append(a)
        else:
            if d is not None:
                raise ValueError(
                    "damping parameters must not be batched under vmap "
                    "(one parameter set per call)"
                )
            out.append(a)
    return out


class D3Energy(torch.autograd.Function):
    """Atom-resolved D3 energies (body of the reference's dftd3(), disp.py:150-165)."""

    generate_vmap_rule = False

    @staticmethod
    def forward(numbers: Tensor, positions: Tensor, p: Tensor, st: Settings) -> Tensor:
        return _run_energy(numbers, positions, p, st)

    @staticmethod
    def setup_context(ctx, inputs, output):
        numbers, positions, p, st = inputs
        ctx.save_for_backward(numbers, positions, p)
        ctx.st = st

    @staticmethod
    def backward(ctx, gE):
        numbers, positions, p = ctx.saved_tensors
        need_x, need_p = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        if not (need_x or need_p):
            return None, None, None, None
        gx, gp = D3Vjp.apply(numbers, positions, p, gE, ctx.st, bool(need_p))
        gp_tot = None
        if need_p:
            gp_tot = gp.reshape(-1, NPAR).sum(0).to(device=p.device, dtype=p.dtype)
        return None, (gx if need_x else None), gp_tot, None

    @staticmethod
    def vmap(info, in_dims, numbers, positions, p, st):
        args = (numbers, positions, p, st)
        if _batched_params(in_dims, args, {0, 1}):
            return _loop(D3Energy.apply, info, in_dims, args)
        numbers, positions, p, st = _front(info, in_dims, args, {0, 1})
        return D3Energy.apply(numbers, positions, p, st), 0


class D3Vjp(torch.autograd.Function):
    """(x, p, g) -> (dL/dx [..., N, 3], dL/dp per structure [..., 5]), L = sum g E."""

    generate_vmap_rule = False

    @staticmethod
    def forward(numbers, positions, p, g, st, want_p=True):
        return _run_vjp(numbers, positions, p, g, st, want_p)

    @staticmethod
    def setup_context(ctx, inputs, output):
        numbers, positions, p, g, st, _ = inputs
        ctx.save_for_backward(numbers, positions, p, g)
        ctx.st = st

    @staticmethod
    def backward(ctx, vx, vp):
        numbers, positions, p, g = ctx.saved_tensors
        lead = positions.shape[:-2]
        n = positions.shape[-2]
        if vx is None:
            vx = torch.zeros_like(positions)
        # a cotangent on the per-structure parameter gradients acts like one
        # tangent on the (shared) parameters only if it is the same for every
        # structure; D3Energy.backward sums over structures, so it is.
        if vp is None:
            vp_shared = torch.zeros(NPAR, dtype=p.dtype, device=p.device)
        else:
            flat = vp.reshape(-1, NPAR)
            vp_shared = flat[0].to(device=p.device, dtype=p.dtype)
        de, dgx, dgp = D3Hvp.apply(numbers, positions, p, g, vx, vp_shared, ctx.st)
        gp_tot = dgp.reshape(-1, NPAR).sum(0).to(device=p.device, dtype=p.dtype)
        return None, dgx, gp_tot, de, None, None

    @staticmethod
    def vmap(info, in_dims, numbers, positions, p, g, st, want_p):
        args = (numbers, positions, p, g, st, want_p)
        if _batched_params(in_dims, args, {0, 1, 3}):
            return _loop(D3Vjp.apply, info, in_dims, args)
        numbers, positions, p, g, st, want_p = _front(info, in_dims, args, {0, 1, 3})
        gx, gp = D3Vjp.apply(numbers, positions, p, g, st, want_p)
        return (gx, gp), (0, 0)


class D3Hvp(torch.autograd.Function):
    """Tangents of (E, dL/dx, dL/dp) along (vx, vp): second-order path."""

    generate_vmap_rule = False

    @staticmethod
    def forward(numbers, positions, p, g, vx, vp, st):
        return _run_hvp(numbers, positions, p, g, vx, vp, st)

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
    def vmap(info, in_dims, numbers, positions, p, g, vx, vp, st):
        args = (numbers, positions, p, g, vx, vp, st)
        if _batched_params(in_dims, args, {0, 1, 3, 4}):
            return _loop(D3Hvp.apply, info, in_dims, args)
        numbers, positions, p, g, vx, vp, st = _front(info, in_dims, args, {0, 1, 3, 4})
        out = D3Hvp.apply(numbers, positions, p, g, vx, vp, st)
        return out, (0, 0, 0)


def d3_energy(numbers: Tensor, positions: Tensor, p: Tensor, st: Settings) -> Tensor:
    return D3Energy.apply(numbers, positions, p, st)
