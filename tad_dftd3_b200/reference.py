"""
Reference systems of the D3 model (mirror of the reference's `Reference`
holder, `reference.py:190-341`): `cn [104, 7]`, `c6 [104, 104, 7, 7]`,
same-device / same-dtype / shape checks, `.to()` and `.type()` returning new
objects.

The C6 table (`reference-c6.pt`, 4.2 MB) is not bundled with this repository
(it is a missing large blob of the reference checkout as well).  It is loaded
from `$TAD_DFTD3_B200_REFERENCE_C6`, from `reference-c6.pt` next to this file,
or from an installed `tad_dftd3`; otherwise pass `c6=` / `ref=` explicitly.
"""
from __future__ import annotations

import os
from typing import Any, NoReturn, Optional

import torch
from torch import Tensor

from .data import REFCN

__all__ = ["Reference"]


def _load_cn(dtype: torch.dtype = torch.double, device: Optional[torch.device] = None) -> Tensor:
    return REFCN(dtype, device)


def _c6_path() -> str | None:
    cands = [os.environ.get("TAD_DFTD3_B200_REFERENCE_C6", ""),
             os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference-c6.pt")]
    try:
        import importlib.util

        spec = importlib.util.find_spec("tad_dftd3")
        if spec is not None and spec.submodule_search_locations:
            cands.append(os.path.join(list(spec.submodule_search_locations)[0], "reference-c6.pt"))
    except (ImportError, ValueError):
        pass
    for c in cands:
        if c and os.path.isfile(c) and os.path.getsize(c) > 1_000_000:
            return c
    return None


def _load_c6(dtype: torch.dtype = torch.double, device: Optional[torch.device] = None) -> Tensor:
    path = _c6_path()
    if path is None:
        raise FileNotFoundError(
            "reference-c6.pt (the [104,104,7,7] D3 reference C6 table) was not "
            "found. Set $TAD_DFTD3_B200_REFERENCE_C6, copy the file next to "
            "tad_dftd3_b200/reference.py, install tad-dftd3, or pass "
            "`ref=Reference(c6=...)`."
        )
    return torch.load(path, map_location=device, weights_only=True).type(dtype)


class Reference:
    """Reference systems for the D3 dispersion model."""

    __slots__ = ["c6", "cn", "__dtype", "__device", "_cache"]

    def __init__(
        self,
        cn: Optional[Tensor] = None,
        c6: Optional[Tensor] = None,
        device: Optional[torch.device] = None,
        dtype: Optional[torch.dtype] = None,
    ):
        if cn is None:
            cn = _load_cn(dtype if dtype is not None else torch.get_default_dtype(), device)
        if c6 is None:
            c6 = _load_c6(dtype if dtype is not None else torch.get_default_dtype(), device)
        self.cn = cn
        self.c6 = c6
        self.__dtype = c6.dtype
        self.__device = c6.device
        self._cache = {}

        if any(t.device != self.device for t in (self.cn, self.c6)):
            raise RuntimeError("All tensors must be on the same device!")
        if any(t.dtype != self.dtype for t in (self.cn, self.c6)):
            raise RuntimeError("All tensors must have the same dtype!")
        s6, sn = self.c6.shape, self.cn.shape
        if s6[-2] != s6[-1] or s6[-1] != sn[-1] or s6[-4] != s6[-3] or s6[-3] != sn[-2]:
            raise RuntimeError("`c6` & `cn` size mismatch found")

    @property
    def device(self) -> torch.device:
        return self.__device

    @device.setter
    def device(self, *_: Any) -> NoReturn:
        raise AttributeError("Move object to device using the `.to` method")

    @property
    def dtype(self) -> torch.dtype:
        return self.__dtype

    def to(self, device: Optional[torch.device] = None, dtype: Optional[torch.dtype] = None) -> "Reference":
        if self.__device == device or device is None:
            if dtype is not None:
                return self.type(dtype)
            return self
        return self.__class__(self.cn.to(device=device, dtype=dtype), self.c6.to(device=device, dtype=dtype))

    def type(self, dtype: torch.dtype) -> "Reference":
        if self.__dtype == dtype:
            return self
        return self.__class__(self.cn.type(dtype), self.c6.type(dtype))

    def __str__(self) -> str:
        return (
            f"{self.__class__.__name__}(n_element={self.cn.shape[-2]}, "
            f"n_reference={self.cn.shape[-1]}, dtype={self.__dtype}, device={self.__device})"
        )

    __repr__ = __str__

    # ---- helpers for the CUDA path (not part of the reference API) ----------
    def _prepared(self, device: torch.device, dtype: torch.dtype):
        """(cn, c6, symmetric) contiguous on `device`/`dtype`, cached per object.

        The kernels index the tables as `[104, 7]` / `[104, 104, 7, 7]` (`refc6 + (Zi*104+Zj)*49`),
        so any other shape is rejected here instead of being read out of bounds.  The cache entry is
        keyed on the identity and the version counters of `cn` / `c6`: reassigning or editing either
        in place invalidates it."""
        from .defaults import MAX_ELEMENT

        nz, nref = MAX_ELEMENT, 7
        if tuple(self.cn.shape) != (nz, nref) or tuple(self.c6.shape) != (nz, nz, nref, nref):
            raise ValueError(
                f"tad_dftd3_b200 kernels need reference tables of shape cn ({nz}, {nref}) and "
                f"c6 ({nz}, {nz}, {nref}, {nref}); got cn {tuple(self.cn.shape)}, c6 {tuple(self.c6.shape)}"
            )
        key = (str(device), dtype)
        stamp = (id(self.cn), getattr(self.cn, "_version", 0), id(self.c6), getattr(self.c6, "_version", 0))
        hit = self._cache.get(key)
        if hit is None or hit[3] != stamp:
            from .disp import _plain

            # (inside jacrev / vmap even `detach()` of a plain tensor comes back wrapped at the live
            # levels; only plain tensors may be cached, so unwrap the results as well)
            cn = _plain(_plain(self.cn).detach().to(device=device, dtype=dtype).contiguous())
            c6 = _plain(_plain(self.c6).detach().to(device=device, dtype=dtype).contiguous())
            sym = bool(torch.allclose(c6, c6.permute(1, 0, 3, 2), rtol=1e-6 if dtype == torch.float32 else 1e-12, atol=0))
            hit = (cn, c6, sym, stamp)
            self._cache[key] = hit
        return hit[:3]
