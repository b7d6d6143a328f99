"""
Dispersion energy: the drop-in for the reference's `disp.py`.

`dftd3()` keeps the reference's signature, defaults and errors
(disp.py:79-165) and dispatches the whole body -- coordination number,
reference weights, C6 interpolation, Becke-Johnson two-body sum, optional
Axilrod-Teller-Muto term, and their analytic first/second derivatives -- to
the sm_100a kernels behind `libd3b200.so` (`ops.D3Energy`).

There is no CPU path and no torch-eager fallback: non-CUDA inputs and
non-default counting / weighting / damping callables raise.
"""
from __future__ import annotations

import weakref
from typing import Any

import torch
from torch import Tensor

from . import data, defaults
from .damping import rational_damping
from .model import gaussian_weight
from .ncoord import exp_count
from .ops import Settings, d3_energy
from .reference import Reference

__all__ = ["dftd3", "hessian", "dispersion", "dispersion2", "dispersion3"]

_Z2S_104 = "Rf"  # element 104, named in the reference's error message (disp.py:134)


def _is_functorch(x: Tensor) -> bool:
    f = torch._C._functorch
    return bool(f.is_batchedtensor(x) or f.is_gradtrackingtensor(x))


def _is_functorch_grad(x: Tensor) -> bool:
    return bool(torch._C._functorch.is_gradtrackingtensor(x))


def _plain(t: Tensor) -> Tensor:
    """Strip functorch grad-tracking wrappers from a CONSTANT tensor before it
    is cached (tensors created inside jacrev/vmap are wrapped at the live
    levels; a cached wrapper would outlive its level)."""
    f = torch._C._functorch
    while f.is_gradtrackingtensor(t):
        t = f.get_unwrapped(t)
    if f.is_batchedtensor(t):
        raise ValueError("model tables must not be batched under vmap")
    return t


_checked: dict = {}  # id(tensor) -> (weakref, version)


def _check_numbers(numbers: Tensor) -> None:
    """ValueError for Z >= 104 (disp.py:130-135); one device sync per distinct tensor."""
    if _is_functorch(numbers) or numbers.numel() == 0:
        return
    key = id(numbers)
    hit = _checked.get(key)
    if hit is not None and hit[0]() is numbers and hit[1] == numbers._version:
        return
    lo, hi = torch.aminmax(numbers)
    if int(hi) >= defaults.MAX_ELEMENT:
        raise ValueError(
            f"No D3 parameters available for Z > {defaults.MAX_ELEMENT - 1} ({_Z2S_104})."
        )
    if int(lo) < 0:
        raise ValueError("Atomic numbers must be non-negative.")
    if len(_checked) > 256:
        for k in [k for k, v in _checked.items() if v[0]() is None]:
            del _checked[k]
    _checked[key] = (weakref.ref(numbers), numbers._version)


_table_cache: dict = {}


def _element_table(name: str, dtype, device) -> Tensor:
    key = (name, dtype, str(device))
    t = _table_cache.get(key)
    if t is None:
        t = _plain(getattr(data, name)(dtype=dtype, device=device)).contiguous()
        _table_cache[key] = t
    return t


def _scalar(v: Any) -> float:
    return float(v.detach()) if isinstance(v, Tensor) else float(v)


def _param_vector(param: dict, dtype, damping_id: int = 0) -> Tensor:
    """(s6, s8, a1, a2, s9) with the reference's defaults -- (s6, s8, rs6, rs8, s9) for zero damping --;
    keeps the autograd graph."""
    keys = ("s6", "s8", "a1", "a2", "s9") if damping_id == 0 else ("s6", "s8", "rs6", "rs8", "s9")
    dflt = (defaults.S6, defaults.S8, defaults.A1, defaults.A2, 0.0) if damping_id == 0 else (defaults.S6, defaults.S8, 1.0, 1.0, 0.0)
    vals = [param.get(k, d) for k, d in zip(keys, dflt)]
    live = [v for v in vals if isinstance(v, Tensor) and (v.requires_grad or _is_functorch(v))]
    if not live:
        key = tuple(_scalar(v) for v in vals) + (dtype, damping_id)
        t = _pvec_cache.get(key)
        if t is None:
            if len(_pvec_cache) > 64:
                _pvec_cache.clear()
            # (created inside jacrev/vmap the tensor is wrapped at the live levels;
            # only the plain tensor may outlive the transform)
            t = _pvec_cache[key] = _plain(torch.tensor(key[:5], dtype=dtype))
        return t
    dev = live[0].device
    return torch.stack([
        (v.to(device=dev, dtype=dtype) if isinstance(v, Tensor) else torch.tensor(v, dtype=dtype, device=dev)).reshape(())
        for v in vals
    ])


_pvec_cache: dict = {}


def _atm_requested(param: dict) -> bool:
    # reference disp.py:234 -- a data-dependent Python branch there as well
    return "s9" in param and bool(param["s9"] != 0.0)


def dftd3(
    numbers: Tensor,
    positions: Tensor,
    param: dict,
    *,
    ref: Reference | None = None,
    rcov: Tensor | None = None,
    rvdw: Tensor | None = None,
    r4r2: Tensor | None = None,
    cutoff: Tensor | None = None,
    counting_function=exp_count,
    weighting_function=gaussian_weight,
    damping_function=rational_damping,
    chunk_size: int | None = None,
    _raw: bool = False,
) -> Tensor:
    """
    Atom-resolved DFT-D3 dispersion energy of a structure `(N,)` or a
    zero-padded batch `(B, N)` -- same contract as the reference's
    `tad_dftd3.dftd3` (disp.py:79-165).

    `chunk_size` is accepted and ignored: no `N x N x 7 x 7` tensor exists.
    """
    from .damping import DAMPING_IDS
    from .ncoord import COUNTING_IDS

    if counting_function not in COUNTING_IDS or weighting_function is not gaussian_weight \
            or damping_function not in DAMPING_IDS:
        raise NotImplementedError(
            "tad_dftd3_b200 implements these model functions as CUDA kernels: counting_function exp_count "
            "(default) or erf_count, weighting_function gaussian_weight, damping_function rational_damping "
            "(default) or zero_damping (the objects exported by tad_dftd3_b200.ncoord / .model / .damping); "
            "arbitrary Python callables cannot run inside the kernels and are not supported."
        )
    counting_id, kcn = COUNTING_IDS[counting_function]
    damping_id = DAMPING_IDS[damping_function]
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(
            f"Shape of positions ({tuple(positions.shape[:-1])}) is not consistent "
            f"with atomic numbers ({tuple(numbers.shape)})."
        )
    if positions.shape[-1] != 3:
        raise ValueError("positions must have shape (..., nat, 3)")
    _check_numbers(numbers)

    dtype, device = positions.dtype, positions.device
    if device.type != "cuda":
        raise RuntimeError(
            "tad_dftd3_b200.dftd3 runs on CUDA devices only (no CPU fallback); "
            f"positions are on {device}."
        )

    if ref is None:
        ref = _default_reference(dtype, device)
    refcn, refc6, symmetric = ref._prepared(device, dtype)
    if not symmetric:
        raise ValueError(
            "reference C6 table must be symmetric: c6[z1,z2,a,b] == c6[z2,z1,b,a]"
        )

    if rcov is None:
        rcov_t, rcov_pe = _element_table("COV_D3", dtype, device), True
    else:
        if rcov.shape != numbers.shape:
            raise ValueError(
                f"Shape of covalent radii {tuple(rcov.shape)} is not consistent with ({tuple(numbers.shape)})."
            )
        rcov_t, rcov_pe = rcov, False
    if r4r2 is None:
        r4r2_t, r4r2_pe = _element_table("R4R2", dtype, device), True
    else:
        if r4r2.shape != numbers.shape:
            raise ValueError("Shape of expectation values is not consistent with atomic numbers.")
        r4r2_t, r4r2_pe = r4r2, False

    # The reference propagates gradients into every tensor argument through autograd; the kernels
    # provide them for positions and s6, s8, a1, a2, s9 only.  Anything else that asks for a gradient
    # is refused instead of being detached silently (same rule as the stage functions).
    for name, t in (("alp", param.get("alp")), ("cutoff", cutoff), ("rcov", rcov), ("r4r2", r4r2), ("rvdw", rvdw)):
        if isinstance(t, Tensor) and (t.requires_grad or _is_functorch_grad(t)):
            raise NotImplementedError(
                f"tad_dftd3_b200 does not provide derivatives with respect to `{name}` "
                "(analytic derivatives exist for positions and s6, s8, a1, a2, s9)."
            )

    atm = _atm_requested(param)
    rvdw_t, rvdw_mode = None, 0
    if atm:
        if rvdw is None:
            rvdw_t, rvdw_mode = _element_table("VDW_PAIRWISE", dtype, device), 1
        else:
            if rvdw.shape != (*numbers.shape, numbers.shape[-1]):
                raise ValueError("Shape of van-der-Waals radii is not consistent with atomic numbers.")
            rvdw_t, rvdw_mode = rvdw, 2

    st = Settings(
        rcov_per_element=rcov_pe, r4r2_per_element=r4r2_pe, rvdw_mode=rvdw_mode,
        cutoff=defaults.D3_DISP_CUTOFF if cutoff is None else _scalar(cutoff),
        cn_cutoff=defaults.D3_CN_CUTOFF, kcn=kcn,
        alp=_scalar(param.get("alp", defaults.ALP)), counting=counting_id, damping=damping_id,
    )
    p = _param_vector(param if atm else {k: v for k, v in param.items() if k != "s9"}, dtype, damping_id)
    if _raw:
        return p, rcov_t, r4r2_t, rvdw_t, refcn, refc6, st
    return d3_energy(numbers, positions, p, rcov_t, r4r2_t, rvdw_t, refcn, refc6, st)


def hessian(numbers: Tensor, positions: Tensor, param: dict, **kwargs) -> Tensor:
    """
    Nuclear Hessian of the total dispersion energy of ONE structure,
    `H[i, a, j, b] = d2 (sum_k E_k) / dx_ia dx_jb`, shape `(N, 3, N, 3)` -- what
    `jacrev(jacrev(lambda x: dftd3(numbers, x, param).sum(-1)))(positions)` returns
    (reference examples/hessian.py:84-94, test/test_grad/test_hessian.py), from ONE
    launch of the second-order kernel (`d3b200_batch_hvp_*` with the 3N unit
    tangents as a batch) instead of two nested functorch transforms, whose host
    overhead (about 3 ms) exceeds the kernel time for a few hundred atoms.
    Keyword arguments as for `dftd3()`.  Not differentiable further.
    """
    from .ops import _run_hvp

    if numbers.dim() != 1 or positions.shape != (numbers.shape[0], 3):
        raise ValueError("hessian() expects one structure: numbers (N,), positions (N, 3)")
    n = numbers.shape[0]
    p, rcov_t, r4r2_t, rvdw_t, refcn, refc6, st = dftd3(numbers, positions.detach(), param, _raw=True, **kwargs)
    dtype, device = positions.dtype, positions.device
    if n == 0:
        return torch.zeros(0, 3, 0, 3, dtype=dtype, device=device)
    ntan = 3 * n
    # With the three-body term every tangent of a launch needs its own dual-number sqrt|C6| matrix (16 N^2 bytes
    # in fp64): the 3N unit tangents go in chunks that keep that workspace below 1 GiB.
    per_tangent = 2 * positions.element_size() * n * n if float(p[4]) != 0.0 else 0
    chunk = ntan if per_tangent == 0 else max(1, min(ntan, (1 << 30) // per_tangent))
    eye = torch.eye(ntan, dtype=dtype, device=device).reshape(ntan, n, 3)
    cols = []
    for lo in range(0, ntan, chunk):
        nb = min(chunk, ntan - lo)
        g = torch.ones(nb, n, dtype=dtype, device=device)

        def per_batch(t, per_element):
            return t if (t is None or per_element) else t.unsqueeze(0).expand(nb, *t.shape)

        _, dgrad, _ = _run_hvp(
            numbers.unsqueeze(0).expand(nb, n), positions.detach().unsqueeze(0).expand(nb, n, 3), p.detach(), g,
            eye[lo:lo + nb], torch.zeros(5, dtype=dtype), per_batch(rcov_t, st.rcov_per_element),
            per_batch(r4r2_t, st.r4r2_per_element), per_batch(rvdw_t, st.rvdw_mode != 2), refcn, refc6, st)
        cols.append(dgrad)
    dgrad = cols[0] if len(cols) == 1 else torch.cat(cols)
    # row k of dgrad is H e_k, i.e. column k of the (symmetric) Hessian
    return dgrad.reshape(n, 3, n, 3).permute(2, 3, 0, 1).contiguous()


_default_refs: dict = {}


def _default_reference(dtype, device) -> Reference:
    key = (dtype, str(device))
    r = _default_refs.get(key)
    if r is None:
        r = Reference(dtype=dtype, device=device)
        _default_refs[key] = r
    return r


def dispersion(numbers, positions, param, c6, rvdw=None, r4r2=None,
               damping_function=rational_damping, cutoff=None, **kwargs):
    from .stages import dispersion as _d

    return _d(numbers, positions, param, c6, rvdw, r4r2, damping_function, cutoff, **kwargs)


def dispersion2(numbers, positions, param, c6, r4r2, damping_function, cutoff, **kwargs):
    from .stages import dispersion2 as _d

    return _d(numbers, positions, param, c6, r4r2, damping_function, cutoff, **kwargs)


def dispersion3(numbers, positions, param, c6, rvdw, cutoff, rs9=None):
    from .stages import dispersion3 as _d

    return _d(numbers, positions, param, c6, rvdw, cutoff, rs9)
