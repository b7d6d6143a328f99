"""
ctypes loader of the blocked fp64 C oracle (`oracle/d3_blocked.c`).

TEST INFRASTRUCTURE ONLY (see the header of the C file): `tests/`,
`__graft_entry__` and `bench.py`'s CPU-baseline legs may import this module;
nothing under `tad_dftd3_b200/` does.

    from d3_blocked import dftd3_blocked
    out = dftd3_blocked(numbers, positions, param, refcn=..., refc6=..., rcov=..., r4r2=..., rvdw_table=..., g=...)
    out["energy"], out["grad"], out["cn"], out["decn"]

`build()` compiles the library with gcc (OpenMP) into `oracle/_build/` when the
source is newer than the binary; the binary is git-ignored and travels to the
GPU box with the working tree.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "d3_blocked.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libd3_blocked.so")


class Params(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("s6", "s8", "a1", "a2", "s9", "rs9", "alp", "cutoff", "cn_cutoff", "kcn")]


def build(force: bool = False) -> str:
    fresh = os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC)
    if fresh and not force:
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        if os.path.isfile(LIB):
            return LIB
        raise RuntimeError("gcc not found and oracle/_build/libd3_blocked.so is missing")
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [gcc, "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c11", "-o", LIB, SRC, "-lm"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("gcc failed building the blocked oracle:\n" + res.stderr)
    return LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        h = C.CDLL(build())
        h.d3o_dftd3.restype = C.c_int
        h.d3o_dftd3.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.POINTER(Params), C.c_void_p, C.c_int] + [C.c_void_p] * 4
        _lib = h
    return _lib


def _np(t, dtype):
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(t), dtype=dtype)


def dftd3_blocked(numbers, positions, param: dict, *, refcn, refc6, rcov, r4r2, rvdw_table=None, g=None,
                  cutoff: float = 50.0, cn_cutoff: float = 25.0, kcn: float = 16.0,
                  rs9: float = 1.3333333730697632, want_grad: bool = True) -> dict:
    """One structure `(N,)`.  `rcov`, `r4r2` are per-ATOM arrays (as the reference's keyword
    arguments), `rvdw_table` the `[104, 104]` element-pair table (the reference's `rvdw[N, N]` is
    `table[Zi, Zj]`).  `param` may hold floats or 0-d tensors; `s9` absent or 0 skips the ATM term
    (reference disp.py:234).  Returns numpy arrays: energy [N], grad [N, 3] = d(sum g E)/dx, cn [N],
    decn [N] = dL/dCN."""
    z = _np(numbers, np.int32)
    x = _np(positions, np.float64)
    n = z.shape[0]
    assert x.shape == (n, 3)
    rc, q = _np(rcov, np.float64), _np(r4r2, np.float64)
    assert rc.shape == (n,) and q.shape == (n,)
    cn_t, c6_t = _np(refcn, np.float64), _np(refc6, np.float64)
    assert cn_t.shape == (104, 7) and c6_t.shape == (104, 104, 7, 7)
    p = Params()
    p.s6, p.s8 = float(param.get("s6", 1.0)), float(param.get("s8", 1.0))
    p.a1, p.a2 = float(param.get("a1", 0.4)), float(param.get("a2", 5.0))
    p.s9 = float(param["s9"]) if "s9" in param else 0.0
    p.rs9, p.alp = float(rs9), float(param.get("alp", 14.0))
    p.cutoff, p.cn_cutoff, p.kcn = float(cutoff), float(cn_cutoff), float(kcn)
    rv = None
    if p.s9 != 0.0:
        assert rvdw_table is not None, "the ATM term needs the [104,104] rvdw table"
        rv = _np(rvdw_table, np.float64)
        assert rv.shape == (104, 104)
    gg = None if g is None else _np(g, np.float64)
    out = {k: np.zeros(s) for k, s in (("cn", n), ("energy", n), ("grad", (n, 3)), ("decn", n))}

    def ptr(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)

    rcode = lib().d3o_dftd3(n, ptr(z), ptr(x), ptr(rc), ptr(q), ptr(cn_t), ptr(c6_t), ptr(rv), C.byref(p), ptr(gg),
                            int(want_grad), ptr(out["cn"]), ptr(out["energy"]), ptr(out["grad"]), ptr(out["decn"]))
    if rcode != 0:
        raise MemoryError("d3o_dftd3: allocation failed")
    return out


if __name__ == "__main__":
    print(build(force=True))
