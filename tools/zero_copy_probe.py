"""Experiment: the fused kernel reading its inputs / writing its outputs directly in pinned host memory
(UVA zero-copy) instead of staged copies."""
import ctypes as C
import time

import torch

import tad_dftd3_b200 as d3
from tad_dftd3_b200 import _lib, ops
from tad_dftd3_b200 import synthetic as syn
from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

dev = torch.device("cuda:0")
dt = torch.double
lib = _lib.lib()
numbers_h, positions_h = syn.drug_like_batch(4096, seed=0)
numbers_h, positions_h = numbers_h.pin_memory(), positions_h.to(dt).pin_memory()
numbers, positions = numbers_h.to(dev), positions_h.to(dev)
ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
refcn, refc6, _ = ref._prepared(dev, dt)
p = torch.tensor([1.0, 0.78981345, 0.49484001, 5.73083694, 0.0], dtype=dt)
st = ops.Settings(rcov_per_element=True, r4r2_per_element=True)
call = ops._Call(numbers, positions, p, COV_D3(dt, dev), R4R2(dt, dev), None, refcn, refc6, st, 0)
e_d = torch.empty(numbers.shape, dtype=dt, device=dev)
g_d = torch.empty(*numbers.shape, 3, dtype=dt, device=dev)
e_h = torch.empty(numbers.shape, dtype=dt).pin_memory()
g_h = torch.empty(*numbers.shape, 3, dtype=dt).pin_memory()
fn = lib.d3b200_batch_vjp_f64
stream = torch.cuda.current_stream(dev)


def launch(z, x, e, g, lo=0, hi=4096, s=None):
    n = 120
    off = lo * n
    _lib.check(fn(hi - lo, n, C.c_void_p(z.data_ptr() + off * 8), C.c_void_p(x.data_ptr() + off * 24),
                  C.byref(call.model), C.byref(call.par), None, C.c_void_p(e.data_ptr() + off * 8),
                  C.c_void_p(g.data_ptr() + off * 24), None, None, 0, C.c_void_p((s or stream).cuda_stream)))


def timeit(f, reps=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        f()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


print("resident                     %.3f ms" % timeit(lambda: launch(numbers, positions, e_d, g_d)))
print("inputs zero-copy             %.3f ms" % timeit(lambda: launch(numbers_h, positions_h, e_d, g_d)))
print("outputs zero-copy            %.3f ms" % timeit(lambda: launch(numbers, positions, e_h, g_h)))
print("all zero-copy                %.3f ms" % timeit(lambda: launch(numbers_h, positions_h, e_h, g_h)))
ref_e, ref_g = e_d.cpu(), g_d.cpu()
launch(numbers_h, positions_h, e_h, g_h)
torch.cuda.synchronize()
print("zero-copy equals resident:", torch.equal(ref_e, e_h), torch.equal(ref_g, g_h))

s2 = torch.cuda.Stream(dev)


def in_zero_out_dma(nchunk=8):
    ev = []
    for k in range(nchunk):
        lo, hi = 4096 * k // nchunk, 4096 * (k + 1) // nchunk
        launch(numbers_h, positions_h, e_d, g_d, lo, hi)
        e = torch.cuda.Event()
        e.record(stream)
        with torch.cuda.stream(s2):
            s2.wait_event(e)
            e_h[lo:hi].copy_(e_d[lo:hi], non_blocking=True)
            g_h[lo:hi].copy_(g_d[lo:hi], non_blocking=True)


for nc in (4, 8, 16):
    print("inputs zero-copy, outputs DMA in %2d chunks  %.3f ms" % (nc, timeit(lambda: in_zero_out_dma(nc))))
