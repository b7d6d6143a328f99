#!/usr/bin/env python
"""
Benchmark of the D3(BJ) hot path (BASELINE.json metric: "D3(BJ) energy+grad pair-terms/s
fp64 at 1/2/4/8 B200; ATM triples/s").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload all|c3|c4|c5a|c5b] [--scaling strong|weak]

One "step" = one pass of the hot path (atom-resolved energies AND the complete analytic
gradient) over one batch of synthetic input.  The printed JSON line (rank 0) is the c3
workload -- BASELINE config 3: ONE fixed batch of 4096 chain-grown drug-like molecules of
50..120 atoms, zero-padded to 120, D3(BJ) two-body + CN, fp64 -- and, with the default
`--workload all`, carries the other two throughput configs as sub-records measured in the
same run:  "c4" (config 4: one 50 001-atom water cluster, rows sharded over the GPUs, CN /
dL/dCN / E / gradient exchanged over NCCL) and "c5a" (config 5a: ATM three-body term on a
10 002-atom cluster, cutoff 25 Bohr -> triples/s).

Multi-GPU (torchrun, one rank per GPU): STRONG scaling.  The fixed 4096-molecule batch is
split by structure over the ranks (balanced on the pair count, no data-path collective); the
fixed large structures are split by rows / centres (NCCL exchanges inside a CUDA graph).
Every N > 1 record carries `checks.max_rel_err_vs_single_gpu`: the sharded result against
the same input evaluated on ONE GPU in the same run.  `--scaling weak` gives every rank its
own 4096-molecule batch (round 1's definition).

Keys of the line:
  value        pair-terms/s, inputs resident in HBM: all pairs of the batch / (median over the
               K timed steps of the step time, max over ranks); one fused launch per step through
               the C ABI (`d3b200_batch_vjp_f64`), CUDA events on the launch stream, L2 flushed
               (256 MB write) between steps;
  api_resident the same step through the drop-in API on resident tensors: `dftd3()` +
               `torch.autograd.grad` (SURVEY 8d's definition of the timed region);
  e2e          the same metric through the public host-buffer API (`dftd3_host`): pinned HOST
               numbers/positions in, pinned host energies + gradient out, every step synchronised;
  roofline     FP64 FMA pipe: algorithmic flops per launch (140 per pair-term within
               25 Bohr, 89 beyond) / kernel time, against an FP64 FMA-chain peak measured live on
               the same GPU; `traffic` is read from the named ncu summary under profiles/ (null
               when no summary of the current kernel build is committed);
  cpu_baseline the reference's own torch CPU path (unmodified reference source over the tad-mctc
               stand-in when `baseline/_ref` travelled, else the oracle port) on a bounded sample
               of the same batch, all host threads (rank 0 at N = 1 only).
`--impl reference` times that CPU path alone on the same config (same `config` object).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PARAM = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}  # r2SCAN-D3(BJ)
FLOPS_NEAR, FLOPS_FAR = 140.0, 89.0


# ------------------------------------------------------------------ inputs
def make_batch(nmol, seed):
    from tad_dftd3_b200 import synthetic as syn

    return syn.drug_like_batch(nmol, 50, 120, seed=seed)


def count_pairs(numbers, positions, cutoff=50.0, cn_cutoff=25.0):
    """Exact number of unordered real pairs within the cutoffs (on `positions.device`)."""
    near = far = 0
    for lo in range(0, numbers.shape[0], 512):
        z, x = numbers[lo:lo + 512], positions[lo:lo + 512]
        real = z != 0
        m = real.unsqueeze(-1) & real.unsqueeze(-2)
        d = torch.cdist(x, x)
        iu = torch.triu(torch.ones(z.shape[-1], z.shape[-1], dtype=torch.bool, device=z.device), 1)
        m = m & iu
        near += int((m & (d <= cn_cutoff)).sum())
        far += int((m & (d > cn_cutoff) & (d <= cutoff)).sum())
    return near, far


# ------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.th.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def count_pairs_single(positions, cutoff=50.0, cn_cutoff=25.0, chunk=2048):
    """Exact unordered pair counts of one structure (chunked, on the GPU)."""
    n = positions.shape[0]
    near = far = 0
    idx = torch.arange(n, device=positions.device)
    for lo in range(0, n, chunk):
        d = torch.cdist(positions[lo:lo + chunk], positions)
        upper = idx[None, :] > idx[lo:lo + chunk, None]
        near += int((upper & (d <= cn_cutoff)).sum())
        far += int((upper & (d > cn_cutoff) & (d <= cutoff)).sum())
    return near, far


def fp64_peak(lib, dev, stream):
    out = torch.zeros(1, dtype=torch.double, device=dev)
    blocks, iters = 148 * 8, 4096
    for _ in range(2):
        lib.d3b200_probe_fma_f64(blocks, iters, out.data_ptr(), C.c_void_p(stream.cuda_stream))
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record(stream)
    lib.d3b200_probe_fma_f64(blocks, iters, out.data_ptr(), C.c_void_p(stream.cuda_stream))
    pb.record(stream)
    torch.cuda.synchronize(dev)
    return blocks * 256 * iters * 128 / (pa.elapsed_time(pb) * 1e-3) / 1e12


def count_triples(positions, cutoff):
    """Unordered triples with at least two edges <= cutoff (the ones the reference's
    ordered rule includes for some ordering): sum_i C(n_i, 2) - 2 * #triangles."""
    n = positions.shape[0]
    adj = (torch.cdist(positions, positions) <= cutoff)
    adj.fill_diagonal_(False)
    deg = adj.sum(1).double()
    wedges = float((deg * (deg - 1) / 2).sum())
    a = adj.float()
    tri = 0.0
    for lo in range(0, n, 2048):
        tri += float(((a[lo:lo + 2048] @ a) * a[lo:lo + 2048]).double().sum())
    tri /= 6.0
    return int(round(wedges - 2.0 * tri)), int(round(tri)), int(adj.sum().item() // 2)


def _measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        return None



C3_NMIN, C3_NMAX = 50, 120


def c3_config(nmol, dtype):
    """Workload description shared VERBATIM by both arms (`--impl reference` prints the same object)."""
    return {
        "workload": f"c3: one batch of {nmol} drug-like molecules ({C3_NMIN}-{C3_NMAX} atoms, padded to {C3_NMAX}), "
                    f"D3(BJ) two-body + CN (cutoffs 50 / 25 Bohr), energy+grad {dtype}",
        "nmol": nmol, "seed": 0, "param": PARAM,
        "tables": "synthetic C6 table (reference-c6.pt unavailable offline), D3 covalent radii and r4r2",
    }


def shard_structures(numbers, world):
    """Strong scaling of a batch: structures sorted by size (largest first) and dealt round-robin,
    so every rank gets the same number of structures and nearly the same pair count."""
    counts = (numbers != 0).sum(-1)
    order = torch.argsort(counts, descending=True, stable=True)
    return [order[r::world] for r in range(world)]


# ------------------------------------------------------------------ reference arm
def all_host_cores():
    """The CPU baseline gets every host core: undo the GPU-local affinity of bind_host_thread_to_gpu."""
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except (AttributeError, OSError, ValueError):
        pass
    torch.set_num_threads(os.cpu_count() or 1)


def reference_energy_fn():
    """(energy(numbers, positions) -> per-atom energies on CPU through the reference's own code, kind)."""
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    dt = torch.double
    c6 = syn.synthetic_c6_table(dt)
    rcov_t, r4r2_t = COV_D3(dt), R4R2(dt)
    par = {k: torch.tensor(v, dtype=dt) for k, v in PARAM.items()}
    from refimport import reference_available

    if reference_available() is not None:
        from refimport import import_reference

        d3ref = import_reference()
        from tad_dftd3.reference import Reference

        refobj = Reference(cn=REFCN(dt), c6=c6)

        def energy(numbers, pos, **kw):
            return d3ref.dftd3(numbers, pos, dict(par, **kw.pop("param", {})), ref=refobj, rcov=rcov_t[numbers],
                               r4r2=r4r2_t[numbers], **kw)

        return energy, "reference"
    import d3_oracle as orc

    refcn = REFCN(dt)

    def energy(numbers, pos, **kw):
        kw.pop("chunk_size", None)
        return orc.dftd3(numbers, pos, dict(par, **kw.pop("param", {})), refcn=refcn, refc6=c6, rcov=rcov_t[numbers],
                         r4r2=r4r2_t[numbers], **{k: v for k, v in kw.items() if v is not None})

    return energy, "port"


def reference_chunk_step(numbers, positions, energy):
    """energy + autograd gradient of one chunk of structures through the reference path."""
    pos = positions.clone().requires_grad_(True)
    e = energy(numbers, pos)
    (g,) = torch.autograd.grad(e.sum(), pos)
    return float(e.detach().sum()), g


def reference_single(numbers, positions, param, *, cutoff=None, rvdw_table=None, hessian=False, chunk_size=None):
    """The reference's CPU path on ONE structure: (seconds for energy + gradient [or the Hessian], kind)."""
    energy, kind = reference_energy_fn()
    dt = torch.double
    numbers, positions = numbers.cpu(), positions.cpu().to(dt)
    extra = {k: torch.tensor(float(v), dtype=dt) for k, v in param.items() if k not in PARAM}
    kw = {"param": extra}
    if rvdw_table is not None:
        kw["rvdw"] = rvdw_table.cpu()[numbers.unsqueeze(-1), numbers.unsqueeze(-2)]
    if cutoff is not None:
        kw["cutoff"] = torch.tensor(cutoff, dtype=dt)
    if chunk_size is not None:
        kw["chunk_size"] = chunk_size
    t0 = time.perf_counter()
    if hessian:
        torch.func.jacrev(torch.func.jacrev(lambda p: energy(numbers, p, **dict(kw)).sum(-1)))(positions)
    else:
        pos = positions.clone().requires_grad_(True)
        torch.autograd.grad(energy(numbers, pos, **dict(kw)).sum(), pos)
    return time.perf_counter() - t0, kind


REF_CHUNK = 256          # structures per reference call (the N^2 x 49 gather of 256 molecules is 1.4 GB)
REF_BUDGET_S = 110.0     # wall-clock budget of the timed steps of `--impl reference`


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on the c3 batch, all host threads.
    Each step runs as many 256-structure chunks of the SAME batch as fit the time budget (all 16
    when K is small); consecutive steps walk through the batch."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    all_host_cores()
    energy, kind = reference_energy_fn()
    numbers, positions = make_batch(args.nmol, seed=0)
    chunks = [(numbers[lo:lo + REF_CHUNK], positions[lo:lo + REF_CHUNK]) for lo in range(0, args.nmol, REF_CHUNK)]
    pairs = [sum(count_pairs(z, x)) for z, x in chunks]
    reference_chunk_step(*chunks[0], energy)                       # thread pools, allocator
    t0 = time.perf_counter()
    reference_chunk_step(*chunks[0], energy)
    t_chunk = time.perf_counter() - t0
    per_step = int(max(1, min(len(chunks), REF_BUDGET_S / max(args.steps, 1) / t_chunk)))
    for _ in range(min(args.warmup, 2)):
        reference_chunk_step(*chunks[-1], energy)
    cursor, times, done = 0, [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        n = 0
        for _ in range(per_step):
            reference_chunk_step(*chunks[cursor % len(chunks)], energy)
            n += pairs[cursor % len(chunks)]
            cursor += 1
        times.append(time.perf_counter() - t0)
        done.append(n)
    v = float(np.sum(done) / np.sum(times))
    desc = (f"{per_step} chunk(s) of {REF_CHUNK} structures of the c3 batch per step (of {len(chunks)}; steps walk "
            f"through the batch), energy + autograd gradient, torch {torch.__version__} CPU fp64, "
            f"{torch.get_num_threads()} threads")
    line = {
        "impl": "reference", "metric": "D3(BJ) energy+grad pair-terms/s fp64", "value": v,
        "unit": "pair-terms/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": float(np.mean(times)) * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": c3_config(args.nmol, "f64"),
        "cpu_baseline": {"value": v, "unit": "pair-terms/s", "cores": torch.get_num_threads(), "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": "pair-terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
class Ctx:
    """Rank / device / collective helpers shared by the workloads."""

    def __init__(self, args):
        import torch.distributed as dist

        self.args = args
        self.dist = dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL prints its version banner (and warnings) to stdout: send them to a file so
            # that stdout carries exactly the one JSON line
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/d3b200_nccl_%h_%p.log")
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)  # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize(self.dev)

    def reduce_max(self, *vals):
        t = torch.tensor([float(v) for v in vals], dtype=torch.double, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def reduce_sum(self, *vals):
        t = torch.tensor([float(v) for v in vals], dtype=torch.double, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.tolist()

    def timed(self, fn, steps, flush=True):
        """K steps, each bracketed by CUDA events on the launch stream (L2 flushed before each);
        barrier + synchronize on both sides.  Returns the per-step times in ms."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        self.barrier()
        for a, b in ev:
            if flush:
                self.flush.zero_()
            a.record(self.stream)
            fn()
            b.record(self.stream)
        self.barrier()
        return [a.elapsed_time(b) for a, b in ev]

    def timed_sync(self, fn, steps):
        """K steps, each followed by a stream synchronize (host-visible result), wall clock per step."""
        self.barrier()
        out = []
        for _ in range(steps):
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize(self.dev)
            out.append((time.perf_counter() - t0) * 1e3)
        self.barrier()
        return out


def max_rel_err(a, b, atol=1e-12):
    """Residual of the parity gate |a - b| <= atol + rtol |b|: the smallest rtol that passes
    (entries where the comparison value `b` is not finite are skipped; see `reference_nan_structures`)."""
    ok = torch.isfinite(b)
    d = (a - b).abs()[ok]
    return float(((d - atol).clamp_min(0) / b[ok].abs().clamp_min(1e-300)).max()) if d.numel() else 0.0


def profile_numbers(name):
    """ncu figures of a kernel from a committed summary (profiles/<name>.json), labelled with their source."""
    path = os.path.join(ROOT, "profiles", name + ".json")
    try:
        d = json.load(open(path))
    except (OSError, ValueError):
        return None
    d["source"] = "profiles/" + name + ".json"
    return d


def run_c3(args, cx, clocks_out):
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import _lib, ops
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

    dev, world, rank = cx.dev, cx.world, cx.rank
    lib = _lib.lib()
    dt = torch.double if args.dtype == "f64" else torch.float32
    if args.scaling == "weak":
        nums_all, pos_all = make_batch(args.nmol, seed=rank)
        mine = torch.arange(args.nmol)
    else:
        nums_all, pos_all = make_batch(args.nmol, seed=0)          # the same fixed batch on every rank
        mine = shard_structures(nums_all, world)[rank]
    numbers_h = nums_all[mine].contiguous().pin_memory()
    positions_h = pos_all[mine].to(dt).contiguous().pin_memory()
    numbers, positions = numbers_h.to(dev), positions_h.to(dev)
    near, far = count_pairs(numbers, positions.double())
    pairs, flops = near + far, near * FLOPS_NEAR + far * FLOPS_FAR
    ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
    refcn, refc6, _ = ref._prepared(dev, dt)
    par = {k: torch.tensor(v, dtype=dt) for k, v in PARAM.items()}
    p = torch.tensor([PARAM["s6"], PARAM["s8"], PARAM["a1"], PARAM["a2"], 0.0], dtype=dt)
    st = ops.Settings(rcov_per_element=True, r4r2_per_element=True)
    rcov_t, r4r2_t = COV_D3(dt, dev), R4R2(dt, dev)
    g = torch.ones(numbers.shape, dtype=dt, device=dev)

    # ---- the hot path, inputs resident: one fused launch = energies + gradient
    call = ops._Call(numbers, positions, p, rcov_t, r4r2_t, None, refcn, refc6, st, 0)
    # scheduling hint, as dftd3() applies it from the second call with the same numbers tensor on:
    # largest structures first (computed once, outside the timed region)
    order_hint = ops.batch_order(call.z) if os.environ.get("D3B200_NO_ORDER_HINT") != "1" else None
    call.model.batch_order = order_hint.data_ptr() if order_hint is not None else None
    energy = torch.empty(numbers.shape, dtype=dt, device=dev)
    grad = torch.empty(*numbers.shape, 3, dtype=dt, device=dev)
    fn = getattr(lib, f"d3b200_batch_vjp_{args.dtype}")
    stream = cx.stream

    def step_resident():
        _lib.launch_count += 1
        _lib.check(fn(call.nflat, call.n, call.z.data_ptr(), call.x.data_ptr(), C.byref(call.model),
                      C.byref(call.par), g.data_ptr(), energy.data_ptr(), grad.data_ptr(), None,
                      None, 0, C.c_void_p(stream.cuda_stream)))

    # the same step through the drop-in API (SURVEY 8d's timed region)
    x_leaf = positions.clone().requires_grad_(True)

    def step_api():
        e = d3.dftd3(numbers, x_leaf, par, ref=ref)
        (gx,) = torch.autograd.grad(e.sum(), x_leaf)
        return e, gx

    for _ in range(args.warmup):
        step_resident()
    cx.barrier()
    # sanity of what is being timed (outside the timed region): every output finite,
    # exact zeros on padding, vanishing net force per structure
    assert bool(torch.isfinite(energy).all()) and bool(torch.isfinite(grad).all()), "non-finite output"
    assert bool((energy[numbers == 0] == 0).all())
    net_force = float(grad.sum(1).abs().max())
    checks = {"all_outputs_finite": True, "max_net_force_per_structure": net_force}
    e_api, g_api = step_api()
    checks["api_equals_c_abi_bitwise"] = bool(torch.equal(e_api.detach(), energy) and torch.equal(g_api, grad))
    if world > 1 and args.scaling == "strong":
        # the sharded result against the same batch evaluated on ONE GPU (this one), in the same run
        zf, xf = nums_all.to(dev), pos_all.to(dt).to(dev)
        xf.requires_grad_(True)
        ef = d3.dftd3(zf, xf, par, ref=ref)
        (gf,) = torch.autograd.grad(ef.sum(), xf)
        sel = mine.to(dev)
        err = max(max_rel_err(energy, ef.detach()[sel]), max_rel_err(grad, gf[sel]))
        checks["max_rel_err_vs_single_gpu"] = cx.reduce_max(err)[0]
        del zf, xf, ef, gf

    with ClockSampler(cx.local) as clocks:
        # untimed pre-heat (~1 s of the same step) so that nvidia-smi samples the
        # clocks UNDER LOAD: the K timed steps alone last only a few milliseconds
        t_heat = time.perf_counter()
        while time.perf_counter() - t_heat < 1.0:
            for _ in range(50):
                step_resident()
            torch.cuda.synchronize(dev)
        n_launch = -_lib.launch_count
        t_steps = cx.timed(step_resident, args.steps)
        n_launch += _lib.launch_count
        for _ in range(args.warmup):
            step_api()
        l0 = _lib.launch_count
        t_api = cx.timed(step_api, args.steps)
        api_launches = _lib.launch_count - l0
        # ---- end-to-end through the public host-buffer API from pinned host buffers
        e_h = torch.empty(numbers.shape, dtype=dt).pin_memory()
        g_h = torch.empty(*numbers.shape, 3, dtype=dt).pin_memory()
        host_chunk = args.e2e_chunk
        n_e2e_launch = 1 if host_chunk == 0 else d3.HostPipeline.chunks(numbers.shape[0], host_chunk)

        def step_e2e():
            d3.dftd3_host(numbers_h, positions_h, par, ref=ref, energy_out=e_h, grad_out=g_h,
                          chunk=host_chunk, device=dev, sync=True)

        for _ in range(args.warmup):
            step_e2e()
        n_launch -= _lib.launch_count
        t_e2e = cx.timed_sync(step_e2e, args.steps)
        n_launch += _lib.launch_count
        checks["e2e_equals_resident_bitwise"] = bool(torch.equal(e_h.to(dev), energy) and torch.equal(g_h.to(dev), grad))
    clocks_out.update(clocks.summary())

    peak_tflops = fp_peak(lib, dev, stream, args.dtype)
    ms_med, ms_mean = float(np.median(t_steps)), float(np.mean(t_steps))
    ms, ms_api, ms_e2e = cx.reduce_max(ms_med, float(np.median(t_api)), float(np.median(t_e2e)))
    pairs_all, flops_all, near_all = cx.reduce_sum(pairs, flops, near)
    achieved = flops / (ms_med * 1e-3) / 1e12                          # this GPU's kernel
    prof = profile_numbers("r2_c3_kernel_ncu") if args.dtype == "f64" else None
    rec = {
        "metric": "D3(BJ) energy+grad pair-terms/s fp64",
        "value": pairs_all / (ms * 1e-3), "unit": "pair-terms/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": c3_config(args.nmol, args.dtype),
        "details": {
            "pair_terms_total": int(pairs_all), "pairs_within_25_bohr_total": int(near_all),
            "pair_terms_this_rank": pairs, "structures_this_rank": int(numbers.shape[0]),
            "ms_per_step_mean_rank0": ms_mean,
            "entry": f"d3b200_batch_vjp_{args.dtype} (one fused launch: CN, weights, C6, BJ energy, analytic gradient)",
            "l2": "256 MB buffer written between timed steps",
            "timing": "CUDA events around each of the K steps on the launch stream; median over steps, max over ranks; "
                      "barrier + synchronize on both sides; 1 s untimed pre-heat under the clock sampler",
            "parallelism": (f"the fixed batch split by structure over {world} GPU(s) (largest first, round-robin), "
                            "no data-path collective" if args.scaling == "strong" else
                            f"every rank owns its own batch of {args.nmol} structures (weak scaling)"),
            "cuda_graph": "not needed: the resident step is ONE kernel launch; with 8 GPUs a rank's 512 structures "
                          "are a single wave of CTAs (592 resident per GPU), so the step time is that of its largest "
                          "structures and strong scaling is bound by wave quantisation + launch latency",
            "schedule": "CTAs take the structures largest first (d3b200_model.batch_order, computed once per numbers "
                        "tensor; dftd3() does the same from the second call with the same tensor on; not used by the zero-copy "
                        "end-to-end form, where it bunches the PCIe requests)",
            "checks": checks,
        },
        "api_resident": {
            "value": pairs_all / (ms_api * 1e-3), "unit": "pair-terms/s", "ms_per_step": ms_api,
            "api": "tad_dftd3_b200.dftd3(numbers, positions, param) + torch.autograd.grad(energy.sum(), positions) on resident "
                   "tensors; CUDA events, median over steps, max over ranks",
            "launches_per_step": api_launches / max(args.steps, 1),
        },
        "e2e": {
            "value": pairs_all / (ms_e2e * 1e-3), "unit": "pair-terms/s",
            "h2d_bytes_per_step": int(numbers_h.numel() * 8 + positions_h.numel() * positions_h.element_size()) * world,
            "d2h_bytes_per_step": int((e_h.numel() + g_h.numel()) * e_h.element_size()) * world,
            "ms_per_step": ms_e2e, "ms_per_step_mean_rank0": float(np.mean(t_e2e)),
            "api": f"tad_dftd3_b200.dftd3_host(): pinned host buffers in, pinned host energies + gradient out "
                   f"(d3b200_batch_vjp_host_{args.dtype}); "
                   + ("one launch of the fused kernel reading and writing host memory directly (zero-copy, "
                      "coalesced), " if host_chunk == 0 else
                      (f"{n_e2e_launch} chunks pipelined H2D / kernel / D2H, " if host_chunk > 0 else
                       f"hybrid: inputs staged by the copy engine in {n_e2e_launch} chunks, kernels write the results "
                       "directly into host memory (no D2H copy), "))
                   + "synchronised every step; wall clock per step, median over steps, max over ranks; bytes summed over ranks",
        },
        "gpu_launches": n_launch,
        "roofline": {
            "bound": "fp64" if args.dtype == "f64" else "fp32", "achieved": achieved, "peak": peak_tflops,
            "unit": "TFLOP/s", "frac": achieved / peak_tflops,
            "traffic": None if prof is None else prof.get("dram_bytes_per_launch"),
            "kernel": "d3::molecule_kernel_v2<%s, true, true>" % ("double" if args.dtype == "f64" else "float"),
            "peak_source": "FMA-chain probe of the same type measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
            "from_profile": prof,
            "flops_model": "140/pair-term (r<=25 Bohr), 89 (25<r<=50), SURVEY.md 8(d); this rank's launch",
            "hbm_gbs_measured_peak": _measured_hbm(),
        },
    }
    if world == 1 and not args.no_cpu_baseline and rank == 0:
        all_host_cores()
        energy_fn, kind = reference_energy_fn()
        zc, xc = nums_all[:REF_CHUNK], pos_all[:REF_CHUNK]
        reference_chunk_step(zc, xc, energy_fn)  # warm-up (thread pools, allocator)
        t0 = time.perf_counter()
        e_sum, g_ref = reference_chunk_step(zc, xc, energy_fn)
        dtc = time.perf_counter() - t0
        rp = sum(count_pairs(zc, xc))
        # the sample doubles as a parity check of what was timed (the shard is a permutation of the batch)
        inv = torch.empty_like(mine)
        inv[mine] = torch.arange(mine.numel())
        rows = inv[:REF_CHUNK].to(dev)
        rec["cpu_baseline"] = {
            "value": rp / dtc, "unit": "pair-terms/s", "cores": torch.get_num_threads(), "kind": kind, "seconds": dtc,
            "sample": f"first {REF_CHUNK} structures of the same batch, energy + autograd gradient, "
                      f"torch {torch.__version__} CPU fp64",
            "max_rel_err_gpu_vs_this_run": max_rel_err(grad[rows].cpu().double(), g_ref,
                                                       1e-12 if args.dtype == "f64" else 1e-7),
            # The reference's autograd gradient is NaN for a structure in which every Gaussian weight of an atom
            # underflows (crowded sulfur atoms of the chain-growth generator, CN > 15: the all-zero fallback of
            # model/weights.py:162-174 divides 0 by 0 in backward); the kernels return the finite limit (dw = 0).
            "reference_nan_structures": int((~torch.isfinite(g_ref)).any(-1).any(-1).sum()),
        }
    return rec


def fp_peak(lib, dev, stream, dtype="f64"):
    dt = torch.double if dtype == "f64" else torch.float32
    out = torch.zeros(1, dtype=dt, device=dev)
    blocks, iters = 148 * 8, 4096
    probe = getattr(lib, f"d3b200_probe_fma_{dtype}")
    for _ in range(2):
        probe(blocks, iters, out.data_ptr(), C.c_void_p(stream.cuda_stream))
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record(stream)
    probe(blocks, iters, out.data_ptr(), C.c_void_p(stream.cuda_stream))
    pb.record(stream)
    torch.cuda.synchronize(dev)
    return blocks * 256 * iters * 128 / (pa.elapsed_time(pb) * 1e-3) / 1e12


def _system_params(cutoff, s9):
    from tad_dftd3_b200 import _lib

    par = _lib.Params()
    par.s6, par.s8, par.a1, par.a2 = PARAM["s6"], PARAM["s8"], PARAM["a1"], PARAM["a2"]
    par.s9, par.rs9, par.alp = s9, 1.3333333730697632, 14.0
    par.cutoff, par.cn_cutoff, par.kcn = cutoff, 25.0, 16.0
    return par


def blocked_oracle_baseline(numbers, positions, param, cutoff, rvdw):
    """CPU baseline of kind "port" for structures the reference cannot hold: the blocked fp64 C restatement
    (oracle/d3_blocked.c, OpenMP on all host cores) on the FULL structure.  Returns (seconds, energy, grad)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from d3_blocked import dftd3_blocked
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

    dt = torch.double
    z = numbers.cpu()
    t0 = time.perf_counter()
    o = dftd3_blocked(z, positions.cpu(), param, refcn=REFCN(dt), refc6=syn.synthetic_c6_table(dt), rcov=COV_D3(dt)[z],
                      r4r2=R4R2(dt)[z], rvdw_table=rvdw, cutoff=cutoff)
    return time.perf_counter() - t0, torch.from_numpy(o["energy"]), torch.from_numpy(o["grad"])


def run_system(args, cx, which):
    """One large structure with its rows (c4) / centres (c5a) sharded over the ranks: strong scaling.
    c4: 50 001-atom water cluster, two-body + CN, cutoffs 50 / 25 Bohr  -> pair-terms/s
    c5a: 10 002-atom water cluster, ATM term (s9 = 1), cutoff `--cutoff`  -> triples/s (whole dftd3() step)."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import _lib
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN
    from tad_dftd3_b200.distributed import LOCAL, dftd3_sharded
    from tad_dftd3_b200.engine import SystemEngine

    dev, world, rank, dt = cx.dev, cx.world, cx.rank, torch.double
    lib = _lib.lib()
    atm = which == "c5a"
    nwater = (3334 if args.nwater == 16667 else args.nwater) if atm else args.nwater
    cutoff = args.cutoff if atm else 50.0
    numbers_h, positions_h = syn.water_cluster(nwater, seed=0)
    numbers_h, positions_h = numbers_h.pin_memory(), positions_h.pin_memory()
    numbers, positions = numbers_h.to(dev), positions_h.to(dev)
    n = int(numbers.shape[0])
    ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
    refcn, refc6, _ = ref._prepared(dev, dt)
    rcov_t, r4r2_t = COV_D3(dt, dev), R4R2(dt, dev)
    rvdw = syn.synthetic_rvdw_table(dt, dev) if atm else None
    if atm:
        d3.data.set_vdw_pairwise(rvdw)
    par = _system_params(cutoff, 1.0 if atm else 0.0)
    param_f = dict(PARAM, s9=1.0) if atm else dict(PARAM)
    param = {k: torch.tensor(v, dtype=dt) for k, v in param_f.items()}
    cut_t = torch.tensor(cutoff, dtype=dt)
    if atm:
        units, triangles, edges = count_triples(positions, cutoff)
        flops = units * 140.0
        work = {"triples": units, "all_short_triangles": triangles, "short_edges": edges}
    else:
        near, far = count_pairs_single(positions)
        units, flops = near + far, near * FLOPS_NEAR + far * FLOPS_FAR
        work = {"pair_terms": units, "pairs_within_25_bohr": near}

    eng = SystemEngine(numbers, positions, rcov_t, r4r2_t, refcn, refc6, par, tables=(True, True), rvdw=rvdw)

    def step_resident():
        eng.step(positions)

    x_leaf = positions.clone().requires_grad_(True)

    def step_api():
        e = dftd3_sharded(numbers, x_leaf, param, ref=ref, cutoff=cut_t)
        (gx,) = torch.autograd.grad(e.sum(), x_leaf)
        return e, gx

    e_h = torch.empty(n, dtype=dt).pin_memory()
    g_h = torch.empty(n, 3, dtype=dt).pin_memory()

    def step_e2e():
        eng.step_host(numbers_h, positions_h, e_h, g_h, sync=False)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    cx.barrier()
    energy, grad = eng.energy.clone(), eng.grad.clone()
    checks = {"all_outputs_finite": bool(torch.isfinite(energy).all() and torch.isfinite(grad).all()),
              "max_abs_net_force": float(grad.sum(0).abs().max()), "cuda_graph_captured": eng._graph is not None,
              "total_energy": float(energy.sum())}
    if world > 1:
        solo = SystemEngine(numbers, positions, rcov_t, r4r2_t, refcn, refc6, par, tables=(True, True), rvdw=rvdw,
                            group=LOCAL, graph=False)
        e1, g1 = solo.step(positions)
        err = max(max_rel_err(energy, e1), max_rel_err(grad, g1))
        checks["max_rel_err_vs_single_gpu"] = cx.reduce_max(err)[0]
        del solo
    steps = args.steps if not atm else max(3, min(args.steps, 10))
    l0 = _lib.launch_count
    t_steps = cx.timed(step_resident, steps)
    launches = _lib.launch_count - l0
    for _ in range(3):
        step_api()
    e_api, g_api = step_api()
    # (pairs-once sweeps add with fp64 atomics: equal to rounding, not bitwise)
    checks["max_rel_err_api_vs_engine"] = max(max_rel_err(e_api.detach(), energy), max_rel_err(g_api, grad))
    t_api = cx.timed(step_api, steps)
    for _ in range(3):
        step_e2e()
    t_e2e = cx.timed_sync(step_e2e, steps)
    checks["e2e_max_abs_grad_diff_vs_resident"] = float((g_h.to(dev) - grad).abs().max())
    # the dominant sweep alone (same arrays; results of this extra launch are discarded)
    run = eng.runner
    rows = run.atm_partition(rank, world) if atm else run.partition(rank, world)
    sweep = (lambda: run.sweep_atm(True, rows)) if atm else (lambda: run.sweep_pair(True, rows))
    t_sweep = cx.timed(sweep, min(steps, 5))
    ms, ms_api, ms_e2e, ms_sweep = cx.reduce_max(float(np.median(t_steps)), float(np.median(t_api)),
                                                 float(np.median(t_e2e)), float(np.median(t_sweep)))
    eng.profile_step(positions)
    phases = eng.profile_step(positions)           # eager launches with events at the phase boundaries (this rank)
    peak = fp64_peak(lib, dev, cx.stream)
    unit = "triples/s" if atm else "pair-terms/s"
    achieved = flops / (ms * 1e-3) / 1e12
    prof = profile_numbers("r2_c5a_atm_kernel_ncu" if atm else "r2_c4_pair_kernel_ncu")
    rec = {
        "metric": "ATM energy+grad triples/s fp64" if atm else "D3(BJ) energy+grad pair-terms/s fp64",
        "value": units / (ms * 1e-3), "unit": unit, "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": dict({
            "workload": (f"c5a: dftd3() with the ATM term (s9=1) on a {n}-atom water cluster, cutoff {cutoff} Bohr, "
                         "energy+grad fp64, centres sharded over the GPUs" if atm else
                         f"c4: {n}-atom water cluster, D3(BJ) two-body + CN (cutoffs 50 / 25 Bohr), energy+grad fp64, "
                         "rows sharded over the GPUs")}, **work),
        "details": {
            "entry": "SystemEngine.step(): one CUDA-graph launch = d3b200_system_pack + memset + d3b200_system_{cn,weights,tvec,pair"
                     + (",atm" if atm else "") + ",cnbwd}_f64 + NCCL exchanges (CN, dL/dCN, grad, E) + scatter to the caller's order",
            "sweeps": "pairs-once (fp64 atomics)" if run.symmetric else "full-row (no atomics)",
            "exchange": "none (one GPU)" if world == 1 else
                        ("all-reduce" if (run.symmetric or atm) else "in-place all-gather") + " over NCCL, captured in the graph",
            "dominant_sweep_ms": ms_sweep, "phases_ms_eager_rank0": {k: round(v, 4) for k, v in phases.items()}, "dominant_sweep": "d3b200_system_atm_f64" if atm else "d3b200_system_pair_f64",
            "l2": "256 MB buffer written between timed steps",
            "timing": "CUDA events around each step, median over steps, max over ranks",
            "checks": checks,
        },
        "api_resident": {"value": units / (ms_api * 1e-3), "unit": unit, "ms_per_step": ms_api,
                         "api": "tad_dftd3_b200.distributed.dftd3_sharded() (= dftd3() at one GPU) + torch.autograd.grad on "
                                "resident tensors (cached engine of the structure)"},
        "e2e": {"value": units / (ms_e2e * 1e-3), "unit": unit, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(positions_h.numel() * 8) * world,
                "d2h_bytes_per_step": int(e_h.numel() * 8 + g_h.numel() * 8) * world,
                "api": "SystemEngine.step_host(): pinned host numbers (validated on the host) + positions in, pinned host "
                       "energies + gradient out on every rank, one stream synchronize per step; bytes summed over ranks"},
        "gpu_launches": launches,
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak * world, "unit": "TFLOP/s",
                     "frac": achieved / (peak * world), "traffic": None if prof is None else prof.get("dram_bytes_per_launch"),
                     "from_profile": prof,
                     "kernel": "whole step, all ranks (dominant: " + ("d3::system_atm3_kernel" if atm else "d3::sym_pair_t_kernel / system_pair_t_kernel") + ")",
                     "flops_model": "140 per unordered triple (SURVEY.md 8d)" if atm else
                                    "140/pair-term (r<=25 Bohr), 89 (25<r<=50), SURVEY.md 8(d)",
                     "peak_source": "FP64 FMA-chain probe x n_gpus, measured in this run"},
    }
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        all_host_cores()
        # (1) the reference itself on the largest sub-cluster it can hold
        centre = positions.mean(0)
        nsub = 300 if atm else 2000
        sub = torch.argsort((positions - centre).norm(dim=1))[:nsub]
        if atm:
            t_sub, _, _ = count_triples(positions[sub], cutoff)
            sec, kind = reference_single(numbers[sub], positions[sub], dict(PARAM, s9=1.0), cutoff=cutoff, rvdw_table=rvdw)
            sec0, _ = reference_single(numbers[sub], positions[sub], PARAM, cutoff=cutoff)
            rec["cpu_baseline"] = {"value": t_sub / max(sec - sec0, 1e-9), "unit": unit, "cores": torch.get_num_threads(),
                                   "kind": kind, "seconds": sec - sec0,
                                   "sample": "central 300-atom sub-cluster (the reference's ATM is dense N^3), (two-body + ATM) "
                                             "minus (two-body) wall time, energy + autograd gradient"}
        else:
            sn, sf = count_pairs_single(positions[sub])
            sec, kind = reference_single(numbers[sub], positions[sub], PARAM, chunk_size=256)
            rec["cpu_baseline"] = {"value": (sn + sf) / sec, "unit": unit, "cores": torch.get_num_threads(), "kind": kind,
                                   "seconds": sec,
                                   "sample": "central 2000-atom sub-cluster of the same structure, energy + autograd gradient, "
                                             "chunk_size=256 (the reference cannot hold 50k atoms)"}
        # (2) the blocked C port on the FULL structure: a second CPU baseline and the parity check of this run
        sec, e_o, g_o = blocked_oracle_baseline(numbers, positions, param_f, cutoff, None if rvdw is None else rvdw.cpu())
        rec["cpu_baseline_port"] = {"value": units / sec, "unit": unit, "cores": os.cpu_count(), "kind": "port", "seconds": sec,
                                    "sample": "the FULL structure through oracle/d3_blocked.c (gcc -O2, OpenMP)",
                                    "max_rel_err_gpu_vs_port": max(max_rel_err(energy.cpu(), e_o), max_rel_err(grad.cpu(), g_o))}
    if atm:
        d3.data.set_vdw_pairwise(None)
    eng.release_graph()
    return rec


def run_c5b(args, rank, local, world, dev):
    """BASELINE config 5b: Hessian of a 200-atom molecule by the examples/hessian.py construction."""
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import REFCN

    if rank != 0:
        return
    dt = torch.double
    numbers, positions = syn.drug_like_batch(1, 200, 200, seed=1)
    numbers, positions = numbers[0].to(dev), positions[0].to(dev)
    ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
    param = {k: torch.tensor(v, dtype=dt) for k, v in PARAM.items()}

    def energy(pos):
        return d3.dftd3(numbers, pos, param, ref=ref).sum(-1)

    hess = torch.func.jacrev(torch.func.jacrev(energy))
    times = []
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        h = hess(positions)
        torch.cuda.synchronize(dev)
        if it >= args.warmup:
            times.append(time.perf_counter() - t0)
    sym = float((h.reshape(600, 600) - h.reshape(600, 600).T).abs().max())
    # the same Hessian from one launch of the second-order kernel (tad_dftd3_b200.hessian), without functorch
    direct = []
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        hd = d3.hessian(numbers, positions, param, ref=ref)
        torch.cuda.synchronize(dev)
        if it >= args.warmup:
            direct.append(time.perf_counter() - t0)
    direct_err = float((hd - h).abs().max())
    # end to end: positions from pinned host memory, Hessian back into pinned host memory
    from tad_dftd3_b200 import _lib

    pos_h = positions.cpu().pin_memory()
    h_h = torch.empty(200, 3, 200, 3, dtype=dt).pin_memory()
    e2e, l0 = [], 0
    for it in range(args.warmup + args.steps):
        torch.cuda.synchronize(dev)
        if it == args.warmup:
            l0 = _lib.launch_count
        t0 = time.perf_counter()
        h_h.copy_(hess(pos_h.to(dev, non_blocking=True)), non_blocking=True)
        torch.cuda.synchronize(dev)
        if it >= args.warmup:
            e2e.append(time.perf_counter() - t0)
    launches = _lib.launch_count - l0
    line = {"metric": "Hessian (jacrev o jacrev) wall time, 200 atoms, two-body + CN, fp64", "value": float(np.mean(times)),
            "unit": "s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(times)) * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c5b: examples/hessian.py construction on one 200-atom chain-grown molecule",
                       "hessian_shape": list(h.shape), "max_asymmetry": sym,
                       "direct_hessian_ms": float(np.mean(direct)) * 1e3,
                       "direct_hessian_note": "tad_dftd3_b200.hessian(): one d3b200_batch_hvp_f64 launch on the 3N unit "
                                              "tangents; the jacrev o jacrev value above is dominated by ~3 ms of functorch "
                                              "host overhead (tools/hessian_probe.py)",
                       "direct_vs_jacrev_max_abs_diff": direct_err,
                       "entry": "one d3b200_batch_energy_f64, one d3b200_batch_vjp_f64 and one d3b200_batch_hvp_f64 "
                                "launch (cotangent batch of 600) per Hessian",
                       "reference_cpu_seconds_survey": 10.6},
            "e2e": {"value": float(np.mean(e2e)), "unit": "s", "h2d_bytes_per_step": int(pos_h.numel() * 8),
                    "d2h_bytes_per_step": int(h_h.numel() * 8), "ms_per_step": float(np.mean(e2e)) * 1e3,
                    "api": "torch.func.jacrev(jacrev(dftd3(...).sum)) on positions copied from pinned host memory, "
                           "Hessian copied back to pinned host memory"},
            "gpu_launches": launches}
    if not args.no_cpu_baseline:
        all_host_cores()
        sec, kind = reference_single(numbers, positions, PARAM, hessian=True)
        line["cpu_baseline"] = {"value": sec, "unit": "s", "cores": torch.get_num_threads(), "kind": kind,
                                "sample": "the same 200-atom molecule, jacrev o jacrev through the reference"}
    print(json.dumps(line))



def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nmol", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="c3 under torchrun: strong = the fixed batch split over the ranks (default); weak = one batch per rank")
    ap.add_argument("--e2e-chunk", type=int, default=0,
                    help="c3 end-to-end: 0 = zero-copy form of dftd3_host (default); > 0 = staged form with this many "
                         "structures per pipeline chunk; < 0 = hybrid form (inputs staged, outputs written directly)")
    ap.add_argument("--cutoff", type=float, default=25.0, help="c5a: dispersion/triple cutoff in Bohr")
    ap.add_argument("--workload", default="all", choices=["all", "c3", "c4", "c5a", "c5b"],
                    help="all: c3 line with c4 and c5a sub-records (default); c3 / c4 / c5a / c5b: that workload alone")
    ap.add_argument("--nwater", type=int, default=16667)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="c3 only: arithmetic type of the path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    cx = Ctx(args)
    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import engine

    # one process per GPU: keep this rank's host thread and its page-locked buffers on the
    # NUMA node of its GPU (matters for the end-to-end number when 8 ranks share the host)
    numa_bound = d3.bind_host_thread_to_gpu(cx.local) if os.environ.get("D3B200_NO_NUMA_BIND") != "1" else False
    clocks = {}
    if args.workload == "c5b":
        run_c5b(args, cx.rank, cx.local, cx.world, cx.dev)
        line = None
    elif args.workload in ("c4", "c5a"):
        with ClockSampler(cx.local) as cs:
            line = run_system(args, cx, args.workload)
        line["clocks"] = cs.summary()
    else:
        line = run_c3(args, cx, clocks)
        line["clocks"] = clocks
        line["e2e"]["host_thread_bound_to_gpu_numa_node"] = bool(numa_bound)
        if args.workload == "all" and args.dtype == "f64":
            for which in ("c4", "c5a"):
                with ClockSampler(cx.local) as cs:
                    sub = run_system(args, cx, which)
                sub["clocks"] = cs.summary()
                line[which] = sub
                line["gpu_launches"] += sub["gpu_launches"]
    if cx.rank == 0 and line is not None:
        print(json.dumps(line))
    if cx.world > 1:
        engine.shutdown()
        cx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
