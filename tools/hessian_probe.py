import time, torch, sys
sys.path.insert(0, '/root/repo')
import tad_dftd3_b200 as d3
from tad_dftd3_b200 import synthetic as syn
from tad_dftd3_b200.data import REFCN
dev = torch.device('cuda:0'); dt = torch.double
ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
param = {"s6": torch.tensor(1.0, dtype=dt), "s8": torch.tensor(0.78981345, dtype=dt), "a1": torch.tensor(0.49484001, dtype=dt), "a2": torch.tensor(5.73083694, dtype=dt)}
for nat in (10, 50, 100, 200):
    numbers, positions = syn.drug_like_batch(1, nat, nat, seed=1)
    numbers, positions = numbers[0].to(dev), positions[0].to(dev)
    def energy(pos):
        return d3.dftd3(numbers, pos, param, ref=ref).sum(-1)
    hess = torch.func.jacrev(torch.func.jacrev(energy))
    for _ in range(3): hess(positions)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10): h = hess(positions)
    t1 = time.perf_counter()          # host time to enqueue
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(nat, "host enqueue %.3f ms, total %.3f ms per Hessian" % ((t1 - t0) * 100, (t2 - t0) * 100))

import cProfile, pstats
for nat in (100, 200):
    numbers, positions = syn.drug_like_batch(1, nat, nat, seed=1)
    numbers, positions = numbers[0].to(dev), positions[0].to(dev)
    def energy(pos):
        return d3.dftd3(numbers, pos, param, ref=ref).sum(-1)
    hess = torch.func.jacrev(torch.func.jacrev(energy))
    for _ in range(3): hess(positions)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): hess(positions)
    torch.cuda.synchronize()
    pr.disable()
    print("==== nat", nat)
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
