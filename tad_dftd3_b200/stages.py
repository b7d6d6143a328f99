"""
Stage-level drop-ins: the reference's operator API
(`ncoord.cn_d3`, `model.weight_references`, `model.atomic_c6`,
`disp.dispersion/dispersion2/dispersion3`, `damping.dispersion_atm`), each on
its own CUDA kernel (`d3b200_stage_*`), with the reference's shapes, defaults
and `ValueError`s (README.rst:238-261, disp.py:168-347, damping/atm.py:43-155).

These keep the dense intermediates (`c6 [..., N, N]`) that the fused `dftd3()`
path never materialises.  `atomic_c6` is differentiable with respect to the
weights (the reference's custom AtomicC6 op, model/c6.py:294-480); the other
stages are FORWARD ONLY: differentiable use goes through `dftd3()`, whose
analytic derivatives cover the whole chain.  Inputs that require grad are
rejected there instead of being silently detached.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import Tensor

from . import _lib, data, defaults

__all__ = ["cn_d3", "weight_references", "atomic_c6", "dispersion", "dispersion2", "dispersion3",
           "dispersion_atm"]


def _sfx(dtype):
    if dtype == torch.float64:
        return "f64"
    if dtype == torch.float32:
        return "f32"
    raise TypeError(f"float32 or float64 expected, got {dtype}")


def _cuda(t: Tensor, what: str):
    if t.device.type != "cuda":
        raise RuntimeError(f"tad_dftd3_b200 runs on CUDA devices only (no CPU fallback); {what} is on {t.device}")


def _no_grad(*tensors):
    for t in tensors:
        if isinstance(t, Tensor) and t.requires_grad and torch.is_grad_enabled():
            raise NotImplementedError(
                "this argument of the stage-level function is treated as a constant (derivatives exist for "
                "positions, cn, weights, c6 and s6, s8, a1, a2); use dftd3() or wrap the call in torch.no_grad()"
            )


def _stream(dev):
    _lib.launch_count += 1
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr())


def _z(numbers: Tensor) -> Tensor:
    return (numbers if numbers.dtype == torch.int64 else numbers.to(torch.int64)).contiguous()


def _scalar(v, default=None) -> float:
    if v is None:
        return float(default)
    return float(v.detach()) if isinstance(v, Tensor) else float(v)


def _check_z(numbers: Tensor):
    from .disp import _check_numbers

    _check_numbers(numbers)


def _wants_grad(*tensors) -> bool:
    f = torch._C._functorch
    return any(isinstance(t, Tensor) and ((t.requires_grad and torch.is_grad_enabled()) or f.is_gradtrackingtensor(t))
               for t in tensors)


_SECOND_ORDER = ("the stage-level functions provide first derivatives (vector-Jacobian products on the CUDA "
                 "kernels); for second derivatives use dftd3(), whose backward is differentiable again")


class _CnD3(torch.autograd.Function):
    """cn_d3 with its VJP with respect to the positions (`d3b200_stage_cn_bwd_*`)."""

    @staticmethod
    def forward(ctx, numbers, positions, rcov, cutoff, kcn, counting):
        ctx.save_for_backward(numbers, positions, rcov)
        ctx.cutoff, ctx.kcn, ctx.counting = cutoff, kcn, counting
        return _cn_forward(numbers, positions, rcov, cutoff, kcn, counting)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gcn):
        numbers, positions, rcov = ctx.saved_tensors
        dt, dev = positions.dtype, positions.device
        n = numbers.shape[-1]
        z = _z(numbers).reshape(-1, n)
        x = positions.detach().reshape(-1, n, 3).contiguous()
        rc = rcov.detach().to(dt).reshape(-1, n).contiguous()
        g = gcn.detach().to(dt).expand(numbers.shape).reshape(-1, n).contiguous()
        gx = torch.empty_like(x)
        fn = getattr(_lib.lib(), f"d3b200_stage_cn_model_bwd_{_sfx(dt)}")
        with torch.cuda.device(dev):
            _lib.check(fn(z.shape[0], n, _p(z), _p(x), _p(rc), ctx.cutoff, ctx.kcn, ctx.counting, _p(g), _p(gx),
                          _stream(dev)))
        return None, gx.reshape(positions.shape), None, None, None, None


def _cn_forward(numbers, positions, rcov, cutoff, kcn, counting=0):
    dt, dev = positions.dtype, positions.device
    n = numbers.shape[-1]
    z = _z(numbers).reshape(-1, n)
    x = positions.detach().reshape(-1, n, 3).contiguous()
    rc = rcov.detach().to(dt).reshape(-1, n).contiguous()
    out = torch.empty(z.shape, dtype=dt, device=dev)
    fn = getattr(_lib.lib(), f"d3b200_stage_cn_model_{_sfx(dt)}")
    with torch.cuda.device(dev):
        _lib.check(fn(z.shape[0], n, _p(z), _p(x), _p(rc), cutoff, kcn, counting, _p(out), _stream(dev)))
    return out.reshape(numbers.shape)


def cn_d3(numbers, positions, *, counting_function=None, rcov=None, cutoff=None, **kwargs):
    """D3 coordination number (tad-mctc `cn_d3`, call site disp.py:150-152) with `exp_count` (default) or
    `erf_count`; differentiable with respect to the positions (first order) on the CUDA kernels."""
    from .ncoord import COUNTING_IDS, exp_count

    if counting_function is None:
        counting_function = exp_count
    if counting_function not in COUNTING_IDS:
        raise NotImplementedError("cn_d3 implements the counting functions exp_count and erf_count of "
                                  "tad_dftd3_b200.ncoord as CUDA kernels; other callables are not supported")
    counting, kcn_default = COUNTING_IDS[counting_function]
    _cuda(positions, "positions")
    _no_grad(rcov)
    dt, dev = positions.dtype, positions.device
    if rcov is None:
        rcov = data.COV_D3(dt, dev)[numbers]
    if numbers.shape != rcov.shape:
        raise ValueError(f"Shape of covalent radii {tuple(rcov.shape)} is not consistent with ({tuple(numbers.shape)}).")
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(
            f"Shape of positions ({tuple(positions.shape[:-1])}) is not consistent with atomic numbers ({tuple(numbers.shape)})."
        )
    cut, kcn = _scalar(cutoff, defaults.D3_CN_CUTOFF), float(kwargs.get("kcn", kcn_default))
    if _wants_grad(positions):
        return _CnD3.apply(numbers, positions, rcov, cut, kcn, counting)
    return _cn_forward(numbers, positions, rcov, cut, kcn, counting)


class _Weights(torch.autograd.Function):
    """weight_references with its VJP with respect to the coordination numbers (`d3b200_stage_weights_bwd_*`)."""

    @staticmethod
    def forward(ctx, numbers, cn, refcn):
        ctx.save_for_backward(numbers, cn, refcn)
        return _weights_forward(numbers, cn, refcn)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gw):
        numbers, cn, refcn = ctx.saved_tensors
        dt, dev = cn.dtype, cn.device
        z = _z(numbers).reshape(-1)
        c = cn.detach().reshape(-1).contiguous()
        g = gw.detach().to(dt).expand(*numbers.shape, 7).reshape(-1, 7).contiguous()
        out = torch.empty_like(c)
        fn = getattr(_lib.lib(), f"d3b200_stage_weights_bwd_{_sfx(dt)}")
        with torch.cuda.device(dev):
            _lib.check(fn(C.c_size_t(z.shape[0]), _p(z), _p(c), _p(refcn), _p(g), _p(out), _stream(dev)))
        return None, out.reshape(cn.shape), None


def _weights_forward(numbers, cn, refcn):
    dt, dev = cn.dtype, cn.device
    z = _z(numbers).reshape(-1)
    c = cn.detach().reshape(-1).contiguous()
    out = torch.empty(z.shape[0], 7, dtype=dt, device=dev)
    fn = getattr(_lib.lib(), f"d3b200_stage_weights_{_sfx(dt)}")
    with torch.cuda.device(dev):
        _lib.check(fn(C.c_size_t(z.shape[0]), _p(z), _p(c), _p(refcn), _p(out), _stream(dev)))
    return out.reshape(*numbers.shape, 7)


def weight_references(numbers, cn, reference, weighting_function=None, **kwargs):
    """Gaussian reference weights (model/weights.py:78-176); differentiable with respect to `cn` (first order)."""
    from .model import gaussian_weight

    if weighting_function is not None and weighting_function is not gaussian_weight or kwargs:
        raise NotImplementedError("only the default gaussian_weight (factor 4) is implemented")
    _cuda(cn, "cn")
    refcn, _, _ = reference._prepared(cn.device, cn.dtype)
    if _wants_grad(cn):
        return _Weights.apply(numbers, cn, refcn)
    return _weights_forward(numbers, cn, refcn)


def _c6_launch(name, numbers, a, b, refc6, out_shape):
    """Common launcher of the two C6 primitives (any leading batch dimensions)."""
    dt, dev = b.dtype, b.device
    n = numbers.shape[-1]
    z = _z(numbers).expand(*b.shape[:-2], n).reshape(-1, n).contiguous()
    a2 = a.detach().to(dt).reshape(z.shape[0], n, -1).contiguous()
    b2 = b.detach().reshape(z.shape[0], n, 7).contiguous()
    out = torch.empty(z.shape[0], *out_shape(n), dtype=dt, device=dev)
    fn = getattr(_lib.lib(), f"d3b200_stage_c6_{name}_{_sfx(dt)}")
    with torch.cuda.device(dev):
        _lib.check(fn(z.shape[0], n, _p(z), _p(a2), _p(b2), _p(refc6), _p(out), _stream(dev)))
    return out.reshape(*b.shape[:-2], *out_shape(n))


def _move_batch(info, in_dims, args, per_structure):
    """vmap helper: batch dimension to the front (or expand) for the per-structure
    tensors; like the reference's AtomicC6_V2.vmap (model/c6.py:436-480) the rule only
    moves dimensions and re-dispatches through `apply`."""
    out = []
    for k, (a, d) in enumerate(zip(args, in_dims)):
        if k in per_structure:
            out.append(a.unsqueeze(0).expand(info.batch_size, *a.shape) if d is None else a.movedim(d, 0))
        else:
            if d is not None:
                raise ValueError("Batch size mismatch: the reference C6 table must not be batched under vmap.")
            out.append(a)
    return out


class _C6Bilinear(torch.autograd.Function):
    """out_ij = sum_ab u_ia K[Zi,Zj,a,b] w_jb (K symmetric: K[zi,zj,a,b] = K[zj,zi,b,a])."""

    generate_vmap_rule = False

    @staticmethod
    def forward(numbers, u, w, refc6):
        return _c6_launch("bilinear", numbers, u, w, refc6, lambda n: (n, n))

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(*inputs)

    @staticmethod
    def backward(ctx, g):
        numbers, u, w, refc6 = ctx.saved_tensors
        gu = _C6Contract.apply(numbers, g, w, refc6) if ctx.needs_input_grad[1] else None
        gw = _C6Contract.apply(numbers, g.transpose(-1, -2), u, refc6) if ctx.needs_input_grad[2] else None
        return None, gu, gw, None

    @staticmethod
    def vmap(info, in_dims, *args):
        return _C6Bilinear.apply(*_move_batch(info, in_dims, args, (0, 1, 2))), 0


class _C6Contract(torch.autograd.Function):
    """out_ia = sum_j G_ij sum_b K[Zi,Zj,a,b] w_jb -- the reference's AtomicC6 backward
    (model/c6.py:317-335) as one kernel."""

    generate_vmap_rule = False

    @staticmethod
    def forward(numbers, G, w, refc6):
        return _c6_launch("contract", numbers, G, w, refc6, lambda n: (n, 7))

    @staticmethod
    def setup_context(ctx, inputs, output):
        ctx.save_for_backward(*inputs)

    @staticmethod
    def backward(ctx, v):
        numbers, G, w, refc6 = ctx.saved_tensors
        gG = _C6Bilinear.apply(numbers, v, w, refc6) if ctx.needs_input_grad[1] else None
        gw = _C6Contract.apply(numbers, G.transpose(-1, -2), v, refc6) if ctx.needs_input_grad[2] else None
        return None, gG, gw, None

    @staticmethod
    def vmap(info, in_dims, *args):
        return _C6Contract.apply(*_move_batch(info, in_dims, args, (0, 1, 2))), 0


def atomic_c6(numbers, weights, reference, chunk_size=None):
    """Dense `[..., N, N]` C6 matrix (reference model/c6.py:42-96), differentiable with
    respect to `weights` to any order on the CUDA kernels (the reference's
    hand-written backward, model/c6.py:294-374).  `chunk_size` is accepted and
    ignored: the kernels never build the reference's `N x N x 7 x 7` gather."""
    _cuda(weights, "weights")
    dt, dev = weights.dtype, weights.device
    _, refc6, symmetric = reference._prepared(dev, dt)
    needs_grad = weights.requires_grad and torch.is_grad_enabled()
    f = torch._C._functorch
    if not needs_grad and not (f.is_batchedtensor(weights) or f.is_gradtrackingtensor(weights)):
        n = numbers.shape[-1]
        z = _z(numbers).reshape(-1, n)
        w = weights.detach().reshape(-1, n, 7).contiguous()
        out = torch.empty(z.shape[0], n, n, dtype=dt, device=dev)
        fn = getattr(_lib.lib(), f"d3b200_stage_c6_{_sfx(dt)}")
        with torch.cuda.device(dev):
            _lib.check(fn(z.shape[0], n, _p(z), _p(w), _p(refc6), _p(out), _stream(dev)))
        return out.reshape(*numbers.shape, n)
    if not symmetric:
        raise ValueError("reference C6 table must be symmetric: c6[z1,z2,a,b] == c6[z2,z1,b,a]")
    return _C6Bilinear.apply(numbers, weights, weights, refc6)


def _params(param: dict, cutoff, dt) -> _lib.Params:
    p = _lib.Params()
    p.s6 = _scalar(param.get("s6"), defaults.S6)
    p.s8 = _scalar(param.get("s8"), defaults.S8)
    p.a1 = _scalar(param.get("a1"), defaults.A1)
    p.a2 = _scalar(param.get("a2"), defaults.A2)
    p.s9 = _scalar(param.get("s9"), defaults.S9)
    p.alp = _scalar(param.get("alp"), defaults.ALP)
    p.rs9 = float(torch.tensor(defaults.RS9))  # float32(4/3), reference disp.py:312
    p.cutoff = _scalar(cutoff, defaults.D3_DISP_CUTOFF)
    p.cn_cutoff, p.kcn = defaults.D3_CN_CUTOFF, defaults.D3_KCN
    return p


def _prep(numbers, positions, c6):
    _cuda(positions, "positions")
    dt, dev = positions.dtype, positions.device
    n = numbers.shape[-1]
    z = _z(numbers).reshape(-1, n)
    x = positions.detach().reshape(-1, n, 3).contiguous()
    c = c6.detach().to(dt).reshape(-1, n, n).contiguous()
    return dt, dev, n, z, x, c


_PKEYS2 = (("s6", defaults.S6), ("s8", defaults.S8), ("a1", defaults.A1), ("a2", defaults.A2))
# zero damping reads (s6, s8, rs6, rs8) in the same four slots (disp._param_vector)
_PKEYS2_ZERO = (("s6", defaults.S6), ("s8", defaults.S8), ("rs6", 1.0), ("rs8", 1.0))


def _params2(pvals, cutoff, model) -> _lib.Params:
    par = _lib.Params()
    par.s6, par.s8, par.a1, par.a2 = pvals
    par.cutoff = cutoff
    par.damping, par.alp = model
    return par


class _Disp2(torch.autograd.Function):
    """dispersion2 with its VJP with respect to positions, the dense c6 and (s6, s8, a1, a2)
    (`d3b200_stage_disp2_bwd_*`)."""

    @staticmethod
    def forward(ctx, numbers, positions, c6, r4r2, pvec, cutoff, model):
        ctx.save_for_backward(numbers, positions, c6, r4r2, pvec)
        ctx.cutoff, ctx.model = cutoff, model
        return _disp2_forward(numbers, positions, c6, r4r2, pvec.detach().tolist(), cutoff, model)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, ge):
        numbers, positions, c6, r4r2, pvec = ctx.saved_tensors
        dt, dev, n, z, x, c = _prep(numbers, positions, c6)
        q = r4r2.detach().to(dt).reshape(-1, n).contiguous()
        par = _params2(pvec.detach().tolist(), ctx.cutoff, ctx.model)
        g = ge.detach().to(dt).expand(numbers.shape).reshape(-1, n).contiguous()
        gx, gc6 = torch.empty_like(x), torch.empty_like(c)
        gpar = torch.empty(z.shape[0], n, 4, dtype=dt, device=dev)
        fn = getattr(_lib.lib(), f"d3b200_stage_disp2_bwd_{_sfx(dt)}")
        with torch.cuda.device(dev):
            _lib.check(fn(z.shape[0], n, _p(z), _p(x), _p(c), _p(q), C.byref(par), _p(g), _p(gx), _p(gc6), _p(gpar),
                          _stream(dev)))
        gp = gpar.sum((0, 1)).to(device=pvec.device, dtype=pvec.dtype) if ctx.needs_input_grad[4] else None
        return None, gx.reshape(positions.shape), gc6.reshape(c6.shape).to(c6.dtype), None, gp, None, None


def _disp2_forward(numbers, positions, c6, r4r2, pvals, cutoff, model=(0, defaults.ALP)):
    dt, dev, n, z, x, c = _prep(numbers, positions, c6)
    q = r4r2.detach().to(dt).reshape(-1, n).contiguous()
    par = _params2(pvals, cutoff, model)
    out = torch.empty(z.shape, dtype=dt, device=dev)
    fn = getattr(_lib.lib(), f"d3b200_stage_disp2_{_sfx(dt)}")
    with torch.cuda.device(dev):
        _lib.check(fn(z.shape[0], n, _p(z), _p(x), _p(c), _p(q), C.byref(par), _p(out), _stream(dev)))
    return out.reshape(numbers.shape)


def dispersion2(numbers, positions, param, c6, r4r2, damping_function=None, cutoff=None, **kwargs):
    """Two-body energy from a dense c6 (disp.py:245-302) with `rational_damping` (Becke-Johnson, default) or
    `zero_damping`; differentiable (first order) with respect to positions, c6 and the parameters s6, s8 and
    a1, a2 (rs6, rs8 for zero damping)."""
    from .damping import DAMPING_IDS, rational_damping

    if damping_function is None:
        damping_function = rational_damping
    if damping_function not in DAMPING_IDS or kwargs:
        raise NotImplementedError("dispersion2 implements the damping functions rational_damping and zero_damping "
                                  "of tad_dftd3_b200.damping as CUDA kernels; other callables / keyword arguments "
                                  "are not supported")
    damping = DAMPING_IDS[damping_function]
    _no_grad(r4r2, param.get("alp"))
    _cuda(positions, "positions")
    cut = _scalar(cutoff, defaults.D3_DISP_CUTOFF)
    model = (damping, _scalar(param.get("alp"), defaults.ALP))
    vals = [param.get(k, d) for k, d in (_PKEYS2 if damping == 0 else _PKEYS2_ZERO)]
    if _wants_grad(positions, c6, *vals):
        dt = positions.dtype
        live = [v for v in vals if isinstance(v, Tensor) and v.requires_grad]
        pdev = live[0].device if live else torch.device("cpu")
        pvec = torch.stack([(v.to(device=pdev, dtype=dt) if isinstance(v, Tensor)
                             else torch.tensor(float(v), dtype=dt, device=pdev)).reshape(()) for v in vals])
        return _Disp2.apply(numbers, positions, c6, r4r2, pvec, cut, model)
    return _disp2_forward(numbers, positions, c6, r4r2, [_scalar(v) for v in vals], cut, model)


class _Atm(torch.autograd.Function):
    """dispersion_atm with its VJP with respect to positions and the dense c6 (`d3b200_stage_atm_bwd_*`)."""

    @staticmethod
    def forward(ctx, numbers, positions, c6, rvdw, scalars):
        ctx.save_for_backward(numbers, positions, c6, rvdw)
        ctx.scalars = scalars
        return _atm_forward(numbers, positions, c6, rvdw, scalars)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, ge):
        numbers, positions, c6, rvdw = ctx.saved_tensors
        dt, dev, n, z, x, c = _prep(numbers, positions, c6)
        rv = rvdw.detach().to(dt).reshape(-1, n, n).contiguous()
        par = _atm_params(ctx.scalars)
        g = ge.detach().to(dt).expand(numbers.shape).reshape(-1, n).contiguous()
        gx, gc6 = torch.zeros_like(x), torch.zeros_like(c)
        fn = getattr(_lib.lib(), f"d3b200_stage_atm_bwd_{_sfx(dt)}")
        with torch.cuda.device(dev):
            _lib.check(fn(z.shape[0], n, _p(z), _p(x), _p(c), _p(rv), C.byref(par), _p(g), _p(gx), _p(gc6), _stream(dev)))
        return None, gx.reshape(positions.shape), gc6.reshape(c6.shape).to(c6.dtype), None, None


def _atm_params(scalars) -> _lib.Params:
    par = _lib.Params()
    par.s9, par.alp, par.rs9, par.cutoff = scalars
    return par


def _atm_forward(numbers, positions, c6, rvdw, scalars):
    dt, dev, n, z, x, c = _prep(numbers, positions, c6)
    rv = rvdw.detach().to(dt).reshape(-1, n, n).contiguous()
    par = _atm_params(scalars)
    out = torch.empty(z.shape, dtype=dt, device=dev)
    fn = getattr(_lib.lib(), f"d3b200_stage_atm_{_sfx(dt)}")
    with torch.cuda.device(dev):
        _lib.check(fn(z.shape[0], n, _p(z), _p(x), _p(c), _p(rv), C.byref(par), _p(out), _stream(dev)))
    return out.reshape(numbers.shape)


def dispersion_atm(numbers, positions, c6, rvdw, cutoff, s9=None, rs9=None, alp=None):
    """Axilrod-Teller-Muto energy from dense c6 / rvdw (damping/atm.py:43-155); differentiable (first order) with
    respect to positions and c6."""
    _no_grad(rvdw, s9, rs9, alp)
    _cuda(positions, "positions")
    dt = positions.dtype
    # the reference's default rs9 tensor is float32(4/3) (damping/atm.py:50)
    scalars = (_scalar(s9, defaults.S9), _scalar(alp, defaults.ALP),
               float(torch.tensor(defaults.RS9)) if rs9 is None else float(rs9.detach().to(dt)),
               _scalar(cutoff, defaults.D3_DISP_CUTOFF))
    if _wants_grad(positions, c6):
        return _Atm.apply(numbers, positions, c6, rvdw, scalars)
    return _atm_forward(numbers, positions, c6, rvdw, scalars)


def dispersion3(numbers, positions, param, c6, rvdw, cutoff, rs9=None):
    return dispersion_atm(numbers, positions, c6, rvdw, cutoff,
                          param.get("s9", defaults.S9), rs9, param.get("alp", defaults.ALP))


def dispersion(numbers, positions, param, c6, rvdw=None, r4r2=None, damping_function=None, cutoff=None,
               **kwargs):
    """Two-body (+ ATM if `param["s9"] != 0`) energy from a caller-provided dense c6
    (reference disp.py:168-242)."""
    dt, dev = positions.dtype, positions.device
    if r4r2 is None:
        r4r2 = data.R4R2(dt, dev)[numbers]
    if numbers.shape != positions.shape[:-1]:
        raise ValueError("Shape of positions is not consistent with atomic numbers.")
    if numbers.shape != r4r2.shape:
        raise ValueError("Shape of expectation values is not consistent with atomic numbers.")
    _check_z(numbers)
    energy = dispersion2(numbers, positions, param, c6, r4r2, damping_function, cutoff, **kwargs)
    if "s9" in param and bool(param["s9"] != 0.0):
        if rvdw is None:
            rvdw = data.VDW_PAIRWISE(dt, dev)[numbers.unsqueeze(-1), numbers.unsqueeze(-2)]
        energy = energy + dispersion3(numbers, positions, param, c6, rvdw, cutoff)
    return energy
