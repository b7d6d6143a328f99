import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np
import tad_dftd3_b200 as d3
from tad_dftd3_b200 import synthetic as syn
from tad_dftd3_b200.data import REFCN
dt=torch.double; dev="cuda"
ref=d3.Reference(cn=REFCN(dt,dev), c6=syn.synthetic_c6_table(dt,dev))
d3.data.set_vdw_pairwise(syn.synthetic_rvdw_table(dt))
par={"s8":torch.tensor(0.7898),"a1":torch.tensor(0.4948),"a2":torch.tensor(5.7308)}
paratm=dict(par, s9=torch.tensor(1.0))
def run(numbers, positions, p, **kw):
    x=positions.to(dev).requires_grad_(True)
    e=d3.dftd3(numbers.to(dev), x, p, ref=ref, **kw)
    (g,)=torch.autograd.grad((e*torch.linspace(0.5,1.5,e.numel(),device=dev,dtype=dt).reshape(e.shape)).sum(), x)
    return float(e.sum()), float(g.abs().sum())
n,x=syn.drug_like_batch(37, 3, 77, seed=5)
print("v2 batch", run(n,x,par))
print("v1 atm batch", run(n,x,paratm))
n1,x1=syn.water_cluster(130, seed=3)      # 390 atoms: v2 multi-block
print("v2 390", run(n1,x1,par))
n2,x2=syn.water_cluster(250, seed=3)      # 750 atoms: v1 fused
print("v1 750", run(n2,x2,par))
os.environ["TAD_DFTD3_B200_TILED"]="1"
print("tiled 390", run(n1,x1,par, cutoff=torch.tensor(15.0)))
print("tiled atm 390", run(n1,x1,paratm, cutoff=torch.tensor(12.0)))
os.environ["TAD_DFTD3_B200_NO_TVEC"]="1"
print("tiled atm general 390", run(n1,x1,paratm, cutoff=torch.tensor(12.0)))
os.environ["TAD_DFTD3_B200_TILED"]="0"
h=torch.func.jacrev(torch.func.jacrev(lambda p: d3.dftd3(n[0].to(dev), p, par, ref=ref).sum()))(x[0].to(dev))
print("hess", float(h.abs().sum()))
with torch.no_grad():
    cn=d3.ncoord.cn_d3(n.to(dev), x.to(dev)); w=d3.model.weight_references(n.to(dev), cn, ref); c6=d3.model.atomic_c6(n.to(dev), w, ref)
    print("stages", float(d3.disp.dispersion(n.to(dev), x.to(dev), paratm, c6).sum()))
torch.cuda.synchronize(); print("done")
