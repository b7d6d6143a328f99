"""Where the end-to-end time of dftd3_host goes: host-side enqueue time vs total, per chunk size."""
import time

import torch

import tad_dftd3_b200 as d3
from tad_dftd3_b200 import synthetic as syn
from tad_dftd3_b200.data import REFCN

dev = torch.device("cuda:0")
dt = torch.double
numbers, positions = syn.drug_like_batch(4096, seed=0)
numbers, positions = numbers.pin_memory(), positions.to(dt).pin_memory()
ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
par = {"s6": torch.tensor(1.0), "s8": torch.tensor(0.78981345), "a1": torch.tensor(0.49484001), "a2": torch.tensor(5.73083694)}
e_h = torch.empty(numbers.shape, dtype=dt).pin_memory()
g_h = torch.empty(*numbers.shape, 3, dtype=dt).pin_memory()
for chunk in (0, 296, 444, 888, 1184, 2048, 4096):
    for _ in range(5):
        d3.dftd3_host(numbers, positions, par, ref=ref, energy_out=e_h, grad_out=g_h, chunk=chunk)
    torch.cuda.synchronize()
    n = 30
    t_enq = t_tot = 0.0
    for _ in range(n):
        t0 = time.perf_counter()
        d3.dftd3_host(numbers, positions, par, ref=ref, energy_out=e_h, grad_out=g_h, chunk=chunk, sync=False)
        t1 = time.perf_counter()
        torch.cuda.current_stream().synchronize()
        t2 = time.perf_counter()
        t_enq += t1 - t0
        t_tot += t2 - t0
    print(f"chunk {chunk:5d}: enqueue {1e3 * t_enq / n:.3f} ms, total {1e3 * t_tot / n:.3f} ms, chunks {d3.HostPipeline.chunks(4096, chunk)}")
