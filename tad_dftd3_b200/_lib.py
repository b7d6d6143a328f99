"""
ctypes binding of `libd3b200.so` (the C ABI declared in `include/d3b200.h`).

There is no CPU fallback: if the shared library is missing this module raises
at first use and tells the caller how to build it.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# TAD_DFTD3_B200_LIB points at another build of the same library (kernel experiments)
LIB_PATH = os.environ.get("TAD_DFTD3_B200_LIB") or os.path.join(HERE, "libd3b200.so")

ABI_VERSION = 2


class Params(C.Structure):
    """`d3b200_params` (host memory)."""

    _fields_ = [(k, C.c_double) for k in
                ("s6", "s8", "a1", "a2", "s9", "rs9", "alp", "cutoff", "cn_cutoff", "kcn")] + [
        ("counting", C.c_int32), ("damping", C.c_int32)]   # model functions, 0 = defaults (see d3b200.h)


class Model(C.Structure):
    """`d3b200_model_f64` / `_f32` (same layout: pointers + ints)."""

    _fields_ = [
        ("rcov", C.c_void_p),
        ("r4r2", C.c_void_p),
        ("rcov_per_element", C.c_int),
        ("r4r2_per_element", C.c_int),
        ("refcn", C.c_void_p),
        ("refc6", C.c_void_p),
        ("rvdw", C.c_void_p),
        ("rvdw_mode", C.c_int),
        ("batch_order", C.c_void_p),
    ]


class System(C.Structure):
    """`d3b200_system_f64` / `_f32`."""

    _fields_ = [("natoms", C.c_int), ("row_begin", C.c_int), ("row_end", C.c_int)] + [
        (k, C.c_void_p) for k in
        ("atoms", "aux", "z", "box", "refcn", "refc6", "cn", "w", "dw", "decn", "energy", "grad")] + [
        ("jsplit", C.c_int), ("scratch", C.c_void_p),
        ("nelem", C.c_int), ("zlist", C.c_void_p), ("eidx", C.c_void_p), ("tvec", C.c_void_p),
        ("symmetric", C.c_int), ("tile_stride", C.c_int),
        ("atm_work", C.c_void_p), ("atm_work_bytes", C.c_size_t)]


_P = C.c_void_p
_I = C.c_int
_SIGNATURES = {
    "d3b200_abi_version": ([], C.c_int),
    "d3b200_last_error": ([], C.c_char_p),
    "d3b200_batch_max_atoms_f64": ([_I], C.c_int),
    "d3b200_batch_max_atoms_f32": ([_I], C.c_int),
    "d3b200_pipe_create": ([C.POINTER(C.c_void_p)], C.c_int),
    "d3b200_pipe_destroy": ([_P], C.c_int),
    "d3b200_batch_host_chunks": ([_I, _I, C.POINTER(C.c_int)], C.c_int),
}
for _sfx in ("f64", "f32"):
    _SIGNATURES[f"d3b200_batch_energy_{_sfx}"] = (
        [_I, _I, _P, _P, C.POINTER(Model), C.POINTER(Params), _P, _P, C.c_size_t, _P], C.c_int)
    _SIGNATURES[f"d3b200_batch_vjp_{_sfx}"] = (
        [_I, _I, _P, _P, C.POINTER(Model), C.POINTER(Params), _P, _P, _P, _P, _P, C.c_size_t, _P], C.c_int)
    _SIGNATURES[f"d3b200_batch_hvp_{_sfx}"] = (
        [_I, _I, _P, _P, _P, C.POINTER(Model), C.POINTER(Params), C.POINTER(Params),
         _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P], C.c_int)
    _SIGNATURES[f"d3b200_batch_workspace_bytes_{_sfx}"] = ([_I, _I, _I, _I], C.c_size_t)
    _SIGNATURES[f"d3b200_batch_host_workspace_bytes_{_sfx}"] = ([_I, _I, _I, _I], C.c_size_t)
    _SIGNATURES[f"d3b200_batch_vjp_host_{_sfx}"] = (
        [_P, _I, _I, _P, _P, C.POINTER(Model), C.POINTER(Params), _P, _P, _P, _P, C.c_size_t, _I, _P], C.c_int)

    _SIGNATURES[f"d3b200_probe_fma_{_sfx}"] = ([_I, _I, _P, _P], C.c_int)
    for _k in ("cn", "weights", "cnbwd", "pgrad"):
        _SIGNATURES[f"d3b200_system_{_k}_{_sfx}"] = ([C.POINTER(System), C.POINTER(Params), _P], C.c_int)
    _SIGNATURES[f"d3b200_system_pack_{_sfx}"] = ([_I, _I, _P, _P, _P, _P, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_system_pair_{_sfx}"] = ([C.POINTER(System), C.POINTER(Params), _I, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_cn_{_sfx}"] = ([_I, _I, _P, _P, _P, C.c_double, C.c_double, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_weights_{_sfx}"] = ([C.c_size_t, _P, _P, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_c6_{_sfx}"] = ([_I, _I, _P, _P, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_c6_bilinear_{_sfx}"] = ([_I, _I, _P, _P, _P, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_c6_contract_{_sfx}"] = ([_I, _I, _P, _P, _P, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_disp2_{_sfx}"] = ([_I, _I, _P, _P, _P, _P, C.POINTER(Params), _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_atm_{_sfx}"] = ([_I, _I, _P, _P, _P, _P, C.POINTER(Params), _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_cn_model_{_sfx}"] = ([_I, _I, _P, _P, _P, C.c_double, C.c_double, _I, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_cn_model_bwd_{_sfx}"] = ([_I, _I, _P, _P, _P, C.c_double, C.c_double, _I, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_cn_bwd_{_sfx}"] = ([_I, _I, _P, _P, _P, C.c_double, C.c_double, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_weights_bwd_{_sfx}"] = ([C.c_size_t, _P, _P, _P, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_disp2_bwd_{_sfx}"] = ([_I, _I, _P, _P, _P, _P, C.POINTER(Params), _P, _P, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_stage_atm_bwd_{_sfx}"] = ([_I, _I, _P, _P, _P, _P, C.POINTER(Params), _P, _P, _P, _P], C.c_int)
    _SIGNATURES[f"d3b200_system_atm_{_sfx}"] = ([C.POINTER(System), C.POINTER(Params), _P, _I, _P], C.c_int)
    _SIGNATURES[f"d3b200_system_atm_workspace_bytes_{_sfx}"] = ([_I], C.c_size_t)
    _SIGNATURES[f"d3b200_system_tvec_{_sfx}"] = ([C.POINTER(System), C.POINTER(Params), _P], C.c_int)

_lib = None

# number of kernel launches issued through this binding (bench.py reports it)
launch_count = 0


class D3LibraryError(RuntimeError):
    pass


def exported_symbols():
    """Names that `include/d3b200.h` declares (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise D3LibraryError(
                f"{LIB_PATH} is missing. tad_dftd3_b200 has no CPU or eager "
                "fallback; build the CUDA library with "
                "`python -m tad_dftd3_b200.build` (needs nvcc, sm_100a)."
            )
        handle = C.CDLL(LIB_PATH)
        for name, (argtypes, restype) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.argtypes = argtypes
            fn.restype = restype
        if handle.d3b200_abi_version() != ABI_VERSION:
            raise D3LibraryError("libd3b200.so ABI version mismatch; rebuild it")
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().d3b200_last_error().decode()
        if rc == 2:
            raise ValueError(msg)
        raise D3LibraryError(f"libd3b200 error {rc}: {msg}")
