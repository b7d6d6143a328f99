"""Do the H2D copy of a batch and the fused kernel overlap? (explains the end-to-end floor of dftd3_host)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tad_dftd3_b200 as d3
from tad_dftd3_b200 import synthetic as syn
from tad_dftd3_b200.data import REFCN

dt = torch.double; dev = torch.device("cuda", 0)
numbers, positions = syn.drug_like_batch(4096, 50, 120, seed=0)
nh, ph = numbers.pin_memory(), positions.pin_memory()
ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
par = {k: torch.tensor(v, dtype=dt) for k, v in {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}.items()}
zd, xd = numbers.to(dev), positions.to(dev)
x = xd.clone().requires_grad_(True)
z2, x2 = torch.empty_like(zd), torch.empty_like(xd)
side = torch.cuda.Stream(dev)
main = torch.cuda.current_stream(dev)

def kernel():
    e = d3.dftd3(zd, x, par, ref=ref)          # forward runs the fused energy+gradient kernel
    return e

def copies(nchunk):
    per = 4096 // nchunk
    for k in range(nchunk):
        sl = slice(k * per, (k + 1) * per)
        z2[sl].copy_(nh[sl], non_blocking=True)
        x2[sl].copy_(ph[sl], non_blocking=True)

def timed(fn, reps=20):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts[3:]))

for _ in range(3): kernel()
print("kernel alone (via dftd3 forward) ms", timed(kernel))
for nc in (1, 4, 8):
    print(f"copies alone, {nc} chunk(s) ms", timed(lambda: copies(nc)))
def both(nc):
    with torch.cuda.stream(side):
        copies(nc)
    kernel()
for nc in (1, 4):
    print(f"kernel || copies ({nc} chunks) ms", timed(lambda: both(nc)))
