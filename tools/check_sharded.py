#!/usr/bin/env python
"""
Multi-GPU parity check of the row-sharded large-structure path (run under torchrun):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/check_sharded.py [--nwater 700] [--atm] [--cutoff 25] [--oracle]

Every rank evaluates the same structure (a) alone on its own GPU through `dftd3()` and (b) sharded
over all ranks through `dftd3_sharded()` (NCCL exchanges, cached engine, CUDA graph), for the unit
cotangent and for a random cotangent, over several steps with moving positions.  Rank 0 prints ONE
JSON line with the maximum absolute / relative deviations between (a) and (b) and, with --oracle,
between (b) and the blocked fp64 C oracle (`oracle/d3_blocked.c`).  Used by
`tests/test_multi_gpu.py` and quoted by `bench.py` (`checks.max_rel_err_vs_single_gpu`).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

PARAM = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}


def rel(a, b, atol=1e-12):
    """max |a-b| / (atol/rtol-style scale): the allclose residual max(|a-b| - atol, 0) / |b|."""
    d = (a - b).abs()
    return float(d.max()), float(((d - atol).clamp_min(0) / b.abs().clamp_min(1e-300)).max())


def main():
    import faulthandler

    faulthandler.dump_traceback_later(int(os.environ.get("D3B200_HANG_DUMP_S", "120")), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--nwater", type=int, default=700)
    ap.add_argument("--atm", action="store_true")
    ap.add_argument("--cutoff", type=float, default=None)
    ap.add_argument("--oracle", action="store_true")
    ap.add_argument("--steps", type=int, default=4)
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/d3b200_nccl_%h_%p.log")
    dist.init_process_group("nccl", device_id=dev)

    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import REFCN
    from tad_dftd3_b200.distributed import dftd3_sharded

    dt = torch.double
    numbers, positions = syn.water_cluster(args.nwater, seed=0)
    numbers, positions = numbers.to(dev), positions.to(dev)
    ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
    rvdw = syn.synthetic_rvdw_table(dt)
    d3.data.set_vdw_pairwise(rvdw)
    param = dict(PARAM, s9=1.0) if args.atm else dict(PARAM)
    par = {k: torch.tensor(v, dtype=dt) for k, v in param.items()}
    cutoff = None if args.cutoff is None else torch.tensor(args.cutoff, dtype=dt)
    gen = torch.Generator().manual_seed(3)
    w = (torch.rand(numbers.shape, dtype=dt, generator=gen) + 0.5).to(dev)
    out = {"world": world, "natoms": int(numbers.shape[0]), "atm": bool(args.atm), "steps": args.steps}
    worst = {"e_abs": 0.0, "e_rel": 0.0, "g_abs": 0.0, "g_rel": 0.0, "gw_abs": 0.0, "gw_rel": 0.0}
    for step in range(args.steps):
        x0 = positions + 0.01 * step * torch.sin(positions + step)     # same on every rank
        # (a) this GPU alone
        xa = x0.clone().requires_grad_(True)
        ea = d3.dftd3(numbers, xa, par, ref=ref, cutoff=cutoff)
        (ga,) = torch.autograd.grad(ea.sum(), xa)
        xa2 = x0.clone().requires_grad_(True)
        (gwa,) = torch.autograd.grad((d3.dftd3(numbers, xa2, par, ref=ref, cutoff=cutoff) * w).sum(), xa2)
        # (b) rows sharded over the ranks
        xb = x0.clone().requires_grad_(True)
        eb = dftd3_sharded(numbers, xb, par, ref=ref, cutoff=cutoff)
        (gb,) = torch.autograd.grad(eb.sum(), xb)
        xb2 = x0.clone().requires_grad_(True)
        (gwb,) = torch.autograd.grad((dftd3_sharded(numbers, xb2, par, ref=ref, cutoff=cutoff) * w).sum(), xb2)
        for key, (p, q) in (("e", (eb.detach(), ea.detach())), ("g", (gb, ga)), ("gw", (gwb, gwa))):
            a_, r_ = rel(p, q)
            worst[key + "_abs"] = max(worst[key + "_abs"], a_)
            worst[key + "_rel"] = max(worst[key + "_rel"], r_)
        last = (x0, eb.detach(), gb)
    t = torch.tensor(list(worst.values()), dtype=dt, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = dict(zip(worst.keys(), t.tolist()))
    out["sharded_vs_single_gpu"] = worst
    out["max_rel_err_vs_single_gpu"] = max(worst["e_rel"], worst["g_rel"], worst["gw_rel"])
    if args.oracle and rank == 0:
        from d3_blocked import dftd3_blocked
        from helpers import tables

        refcn, refc6, rcov, r4r2, rv = tables(dt)
        x0, eb, gb = last
        z = numbers.cpu()
        o = dftd3_blocked(z, x0.cpu(), param, refcn=refcn, refc6=refc6, rcov=rcov[z], r4r2=r4r2[z], rvdw_table=rv,
                          cutoff=50.0 if args.cutoff is None else args.cutoff)
        ea_, er_ = rel(eb.cpu(), torch.from_numpy(o["energy"]))
        ga_, gr_ = rel(gb.cpu(), torch.from_numpy(o["grad"]))
        out["sharded_vs_blocked_oracle"] = {"e_abs": ea_, "e_rel": er_, "g_abs": ga_, "g_rel": gr_}
        out["max_rel_err_vs_oracle"] = max(er_, gr_)
    dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    from tad_dftd3_b200.engine import shutdown

    shutdown()          # graphs with captured NCCL work go before the communicator
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
