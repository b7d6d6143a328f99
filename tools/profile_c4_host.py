"""Host-side breakdown of one tiled step on the 50k-atom cluster (plan, runner, sweeps, scatter)."""
import sys, time, torch
sys.path.insert(0, "/root/repo")
import tad_dftd3_b200 as d3
from tad_dftd3_b200 import _lib, synthetic as syn
from tad_dftd3_b200.data import COV_D3, R4R2, REFCN
from tad_dftd3_b200.system import SystemPlan, SystemRunner, row_partition
dt, dev = torch.double, torch.device("cuda", 0)
numbers, positions = syn.water_cluster(16667, seed=0); numbers, positions = numbers.to(dev), positions.to(dev)
ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev)); refcn, refc6, _ = ref._prepared(dev, dt)
rcov, r4r2 = COV_D3(dt, dev)[numbers], R4R2(dt, dev)[numbers]
par = _lib.Params(); par.s6, par.s8, par.a1, par.a2 = 1.0, 0.78981345, 0.49484001, 5.73083694
par.s9, par.rs9, par.alp, par.cutoff, par.cn_cutoff, par.kcn = 0.0, 4/3, 14.0, 50.0, 25.0, 16.0
g = torch.ones(numbers.shape[0], dtype=dt, device=dev)
def sync(): torch.cuda.synchronize()
for world in (1, 8):
    rows = row_partition(50048, world)[0]
    acc = {}
    for it in range(6):
        t = {}
        sync(); t0 = time.perf_counter()
        plan = SystemPlan(numbers, positions, rcov, r4r2, par.kcn); sync(); t["plan"] = time.perf_counter() - t0; t0 = time.perf_counter()
        run = SystemRunner(plan, refcn, refc6, par, g, world); sync(); t["runner"] = time.perf_counter() - t0; t0 = time.perf_counter()
        run.sweep_cn(rows); sync(); t["cn"] = time.perf_counter() - t0; t0 = time.perf_counter()
        run.weights(); sync(); t["weights+tvec"] = time.perf_counter() - t0; t0 = time.perf_counter()
        run.sweep_pair(True, rows); sync(); t["pair"] = time.perf_counter() - t0; t0 = time.perf_counter()
        run.sweep_cnbwd(rows); sync(); t["cnbwd"] = time.perf_counter() - t0; t0 = time.perf_counter()
        e, gr = plan.scatter(run.energy), plan.scatter(run.grad); sync(); t["scatter"] = time.perf_counter() - t0
        if it >= 2:
            for k, v in t.items(): acc[k] = acc.get(k, 0) + v / 4
    print("world", world, "jsplit", run.jsplit, {k: round(v * 1e3, 3) for k, v in acc.items()}, "total ms", round(sum(acc.values()) * 1e3, 3))
