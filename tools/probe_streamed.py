"""Timing probe of the host-buffer forms of d3b200_batch_vjp_host_f64 (zero-copy / staged / hybrid / streamed)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tad_dftd3_b200 as d3
from tad_dftd3_b200 import synthetic as syn
from tad_dftd3_b200.data import REFCN

dt = torch.double; dev = torch.device("cuda", 0)
numbers, positions = syn.drug_like_batch(4096, 50, 120, seed=0)
nh, ph = numbers.pin_memory(), positions.pin_memory()
eh = torch.empty(numbers.shape, dtype=dt).pin_memory(); gh = torch.empty(*numbers.shape, 3, dtype=dt).pin_memory()
ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
par = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}
par = {k: torch.tensor(v, dtype=dt) for k, v in par.items()}
# pure copies
xd = torch.empty_like(positions, device=dev); zd = torch.empty_like(numbers, device=dev)
def t_copy():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    zd.copy_(nh, non_blocking=True); xd.copy_(ph, non_blocking=True); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3
print("H2D 15.7 MB ms", np.median([t_copy() for _ in range(20)]))
ed = torch.empty(numbers.shape, dtype=dt, device=dev); gd = torch.empty(*numbers.shape, 3, dtype=dt, device=dev)
def t_d2h():
    torch.cuda.synchronize(); t0 = time.perf_counter()
    eh.copy_(ed, non_blocking=True); gh.copy_(gd, non_blocking=True); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3
print("D2H 15.7 MB ms", np.median([t_d2h() for _ in range(20)]))
for chunk in (0, -11, -10, -9, -8, 592):
    ts = []
    for it in range(25):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        d3.dftd3_host(nh, ph, par, ref=ref, energy_out=eh, grad_out=gh, chunk=chunk, device=dev, sync=True)
        ts.append((time.perf_counter() - t0) * 1e3)
    print("chunk", chunk, "median ms %.4f min %.4f" % (np.median(ts[5:]), np.min(ts[5:])), flush=True)
