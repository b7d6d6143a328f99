#!/usr/bin/env python
"""
Benchmark of the D3(BJ) hot path (BASELINE.json metric: energy+gradient
pair-terms/s, fp64).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload c3|c4|c5a] [--nmol 4096]

One "step" = one pass of the hot path (atom-resolved energies AND the complete
analytic gradient) over one batch of synthetic structures.  Default workload
(`c3`): BASELINE config 3, 4096 chain-grown drug-like molecules of 50..120
atoms, zero-padded to 120, D3(BJ) two-body + CN, fp64; under torchrun every
rank owns its own 4096-molecule shard (structures are independent: no
data-path collective; weak scaling).

Printed JSON line (rank 0):
  value     pair-terms/s, inputs resident in HBM, one fused launch per step
            (`d3b200_batch_vjp_f64` through the Python binding), CUDA events,
            L2 flushed between steps, max over ranks;
  e2e       the same metric through the public API (`dftd3()` +
            `autograd.grad`) starting from pinned HOST buffers: H2D of numbers
            and positions and D2H of energies and gradients inside the timed
            region;
  roofline  FP64 FMA pipe: algorithmic flops per launch (140 per
            pair-term within 25 Bohr, 89 beyond) / event time of the kernel,
            against an FP64 FMA-chain peak measured live on the same GPU;
  cpu_baseline  the reference's own torch CPU path (unmodified reference
            source over the tad-mctc stand-in when `baseline/_ref` travelled,
            else the oracle port), on a bounded sample, all host threads.
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


# ------------------------------------------------------------------ reference arm
def all_host_cores():
    """The CPU baseline gets every host core: undo the GPU-local affinity of bind_host_thread_to_gpu."""
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except (AttributeError, OSError, ValueError):
        pass
    torch.set_num_threads(os.cpu_count() or 1)


def reference_step_fn(nmol_sample, seed=0):
    """Callable running energy+gradient of the reference CPU path on a sample; returns
    (fn, kind, pair_terms, sample description)."""
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    numbers, positions = make_batch(nmol_sample, seed)
    near, far = count_pairs(numbers, positions)
    dt = torch.double
    c6 = syn.synthetic_c6_table(dt)
    rcov, r4r2 = COV_D3(dt)[numbers], R4R2(dt)[numbers]
    par = {k: torch.tensor(v, dtype=dt) for k, v in PARAM.items()}
    from refimport import reference_available

    if reference_available() is not None:
        from refimport import import_reference

        d3ref = import_reference()
        from tad_dftd3.reference import Reference

        refobj = Reference(cn=REFCN(dt), c6=c6)
        kind = "reference"

        def energy(pos):
            return d3ref.dftd3(numbers, pos, par, ref=refobj, rcov=rcov, r4r2=r4r2)
    else:
        import d3_oracle as orc

        refcn = REFCN(dt)
        kind = "port"

        def energy(pos):
            return orc.dftd3(numbers, pos, par, refcn=refcn, refc6=c6, rcov=rcov, r4r2=r4r2)

    def step():
        pos = positions.clone().requires_grad_(True)
        e = energy(pos)
        (g,) = torch.autograd.grad(e.sum(), pos)
        return float(e.detach().sum()), g

    desc = (f"{nmol_sample} molecules of the c3 batch (seed {seed}), energy + autograd gradient, "
            f"torch {torch.__version__} CPU fp64")
    return step, kind, near + far, desc


def reference_single(numbers, positions, param, *, cutoff=None, rvdw_table=None, hessian=False, chunk_size=None):
    """The reference's own CPU path (unmodified source over the stand-in when `baseline/_ref`
    or /root/reference is present, else the oracle port) on ONE structure: returns
    (seconds for energy + gradient [or the Hessian], kind)."""
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from refimport import reference_available

    dt = torch.double
    numbers, positions = numbers.cpu(), positions.cpu().to(dt)
    c6 = syn.synthetic_c6_table(dt)
    rcov, r4r2 = COV_D3(dt)[numbers], R4R2(dt)[numbers]
    par = {k: torch.tensor(float(v), dtype=dt) for k, v in param.items()}
    rvdw = None if rvdw_table is None else rvdw_table.cpu()[numbers.unsqueeze(-1), numbers.unsqueeze(-2)]
    cut = None if cutoff is None else torch.tensor(cutoff, dtype=dt)
    if reference_available() is not None:
        from refimport import import_reference

        d3ref = import_reference()
        from tad_dftd3.reference import Reference

        refobj = Reference(cn=REFCN(dt), c6=c6)
        kind = "reference"

        def energy(pos):
            return d3ref.dftd3(numbers, pos, par, ref=refobj, rcov=rcov, r4r2=r4r2, rvdw=rvdw, cutoff=cut,
                               chunk_size=chunk_size)
    else:
        import d3_oracle as orc

        kind = "port"

        def energy(pos):
            return orc.dftd3(numbers, pos, par, refcn=REFCN(dt), refc6=c6, rcov=rcov, r4r2=r4r2, rvdw=rvdw,
                             cutoff=50.0 if cutoff is None else cutoff)

    t0 = time.perf_counter()
    if hessian:
        torch.func.jacrev(torch.func.jacrev(lambda p: energy(p).sum(-1)))(positions)
    else:
        pos = positions.clone().requires_grad_(True)
        torch.autograd.grad(energy(pos).sum(), pos)
    return time.perf_counter() - t0, kind


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    step, kind, pairs, desc = reference_step_fn(args.ref_nmol)
    for _ in range(max(args.warmup, 1) if args.warmup < 3 else 1):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = pairs / dt
    line = {
        "impl": "reference", "metric": "D3(BJ) energy+grad pair-terms/s fp64", "value": v,
        "unit": "pair-terms/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c3: 4096 drug-like molecules (50-120 atoms, padded to 120), D3(BJ) two-body + CN, energy+grad fp64",
                   "sample_per_step": desc},
        "cpu_baseline": {"value": v, "unit": "pair-terms/s", "cores": torch.get_num_threads(), "kind": kind, "sample": desc},
        "e2e": {"value": v, "unit": "pair-terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nmol", type=int, default=4096)
    ap.add_argument("--ref-nmol", type=int, default=192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=0,
                    help="c3 end-to-end: 0 = zero-copy form of dftd3_host (default); > 0 = staged form with this many "
                         "structures per pipeline chunk")
    ap.add_argument("--cutoff", type=float, default=25.0, help="c5a: dispersion/triple cutoff in Bohr")
    ap.add_argument("--workload", default="c3", choices=["c3", "c4", "c5a", "c5b"],
                    help="c3: 4096-molecule batch (default, the headline metric); "
                         "c4: one 50k-atom water cluster, rows sharded over the GPUs")
    ap.add_argument("--nwater", type=int, default=16667)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="c3 only: arithmetic type of the path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner (and warnings) to stdout: send them to a file so
        # that stdout carries exactly the one JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/d3b200_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)

    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import _lib, ops
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

    # one process per GPU: keep this rank's host thread and its page-locked buffers on the
    # NUMA node of its GPU (matters for the end-to-end number when 8 ranks share the host)
    numa_bound = d3.bind_host_thread_to_gpu(local) if os.environ.get("D3B200_NO_NUMA_BIND") != "1" else False
    lib = _lib.lib()
    dt = torch.double if args.dtype == "f64" else torch.float32
    if args.workload == "c5a":
        run_c5a(args, rank, local, world, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "c5b":
        run_c5b(args, rank, local, world, dev)
        return
    if args.workload == "c4":
        run_c4(args, rank, local, world, dev)
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- this rank's shard (weak scaling: every rank owns nmol structures)
    numbers_h, positions_h = make_batch(args.nmol, seed=rank)
    numbers_h, positions_h = numbers_h.pin_memory(), positions_h.pin_memory()
    positions_h = positions_h.to(dt).pin_memory()
    numbers, positions = numbers_h.to(dev), positions_h.to(dev)
    near, far = count_pairs(numbers, positions.double())
    pairs = near + far
    flops = near * FLOPS_NEAR + far * FLOPS_FAR
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
    stream = torch.cuda.current_stream(dev)

    def step_resident():
        _lib.launch_count += 1
        _lib.check(fn(call.nflat, call.n, call.z.data_ptr(), call.x.data_ptr(), C.byref(call.model),
                      C.byref(call.par), g.data_ptr(), energy.data_ptr(), grad.data_ptr(), None,
                      None, 0, C.c_void_p(stream.cuda_stream)))

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    # sanity of what is being timed (outside the timed region): every output finite,
    # exact zeros on padding, vanishing net force per structure
    assert bool(torch.isfinite(energy).all()) and bool(torch.isfinite(grad).all()), "non-finite output"
    assert bool((energy[numbers == 0] == 0).all())
    net_force = float(grad.sum(1).abs().max())
    launches0 = _lib.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local) as clocks:
        # untimed pre-heat (~1 s of the same step) so that nvidia-smi samples the
        # clocks UNDER LOAD: the K timed steps alone last only a few milliseconds
        t_heat = time.perf_counter()
        while time.perf_counter() - t_heat < 1.0:
            for _ in range(50):
                step_resident()
            torch.cuda.synchronize(dev)
        barrier()
        n_launch = -_lib.launch_count
        for a, b in ev:
            flush.zero_()
            a.record(stream)
            step_resident()
            b.record(stream)
        barrier()
        n_launch += _lib.launch_count
        t_steps = [a.elapsed_time(b) for a, b in ev]  # ms, kernel only
        # ---- end-to-end through the public API from pinned host buffers
        e_h = torch.empty(numbers.shape, dtype=dt).pin_memory()
        g_h = torch.empty(*numbers.shape, 3, dtype=dt).pin_memory()

        # End to end through the public host-buffer API: `dftd3_host` cuts the batch
        # into chunks of structures and pipelines H2D copy / fused kernel / D2H copy
        # (d3b200_batch_vjp_host_f64); it returns after the results are in e_h / g_h.
        host_chunk = max(0, args.e2e_chunk)
        n_e2e_launch = 1 if host_chunk == 0 else d3.HostPipeline.chunks(args.nmol, host_chunk)
        positions_e2e, ref_e2e = positions_h, ref

        def step_e2e():
            d3.dftd3_host(numbers_h, positions_e2e, par, ref=ref_e2e, energy_out=e_h, grad_out=g_h,
                          chunk=host_chunk, device=dev, sync=True)

        for _ in range(args.warmup):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        n_launch -= _lib.launch_count
        for _ in range(args.steps):
            step_e2e()
        n_launch += _lib.launch_count
        e1.record(stream)
        barrier()
        t_e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3) / args.steps
    launches = n_launch  # our kernels inside the two timed regions (K resident steps + K end-to-end steps)

    # ---- FP64 FMA-chain peak on this GPU (roofline denominator)
    out = torch.zeros(1, dtype=dt, device=dev)
    blocks, iters = 148 * 8, 4096
    probe = getattr(lib, f"d3b200_probe_fma_{args.dtype}")
    for _ in range(2):
        probe(blocks, iters, out.data_ptr(), C.c_void_p(stream.cuda_stream))
    pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pa.record(stream)
    probe(blocks, iters, out.data_ptr(), C.c_void_p(stream.cuda_stream))
    pb.record(stream)
    torch.cuda.synchronize(dev)
    peak_tflops = blocks * 256 * iters * 128 / (pa.elapsed_time(pb) * 1e-3) / 1e12

    ms = float(np.mean(t_steps))
    t = torch.tensor([ms, t_e2e_ms], dtype=torch.double, device=dev)
    tot = torch.tensor([float(pairs), float(flops)], dtype=torch.double, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, e2e_ms = float(t[0]), float(t[1])
    pairs_all, flops_all = float(tot[0]), float(tot[1])

    if rank == 0:
        achieved = flops / (float(np.mean(t_steps)) * 1e-3) / 1e12  # this GPU's kernel
        line = {
            "metric": "D3(BJ) energy+grad pair-terms/s fp64",
            "value": pairs_all / (ms * 1e-3),
            "unit": "pair-terms/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": f"c3: {args.nmol} drug-like molecules per GPU (50-120 atoms, padded to 120), "
                            f"D3(BJ) two-body + CN, energy+grad {args.dtype}",
                "pair_terms_per_gpu": pairs, "pairs_within_25_bohr": near,
                "entry": "d3b200_batch_vjp_f64 (one fused launch: CN, weights, C6, BJ energy, analytic gradient)",
                "l2": "256 MB buffer written between timed steps",
                "timing": "CUDA events around each step on the launch stream, mean of steps, max over ranks; "
                          "1 s untimed pre-heat under the clock sampler",
                "parallelism": f"structures sharded over {world} GPU(s), no data-path collective",
                "tables": "synthetic C6 table (reference-c6.pt unavailable offline)",
                "schedule": "CTAs take the structures largest first (d3b200_model.batch_order, computed once per numbers "
                            "tensor; dftd3() does the same from the second call with the same tensor on; not used by the zero-copy "
                            "end-to-end form, where it bunches the PCIe requests)",
                "checks": {"all_outputs_finite": True, "max_net_force_per_structure": net_force},
            },
            "e2e": {
                "value": pairs_all / (e2e_ms * 1e-3), "unit": "pair-terms/s",
                "h2d_bytes_per_step": int(numbers_h.numel() * 8 + positions_h.numel() * positions_h.element_size()),
                "d2h_bytes_per_step": int((e_h.numel() + g_h.numel()) * e_h.element_size()),
                "ms_per_step": e2e_ms,
                "api": f"tad_dftd3_b200.dftd3_host(): pinned host buffers in, pinned host energies + gradient out "
                       f"(d3b200_batch_vjp_host_{args.dtype}); "
                       + ("one launch of the fused kernel reading and writing host memory directly (zero-copy, "
                          "coalesced), " if host_chunk == 0 else
                          f"{n_e2e_launch} chunks pipelined H2D / kernel / D2H, ")
                       + "synchronised every step",
                "host_thread_bound_to_gpu_numa_node": bool(numa_bound),
            },
            "gpu_launches": launches,
            "roofline": {
                "bound": "fp64" if args.dtype == "f64" else "fp32", "achieved": achieved, "peak": peak_tflops,
                "unit": "TFLOP/s", "frac": achieved / peak_tflops, "traffic": 15.85e6 if args.dtype == "f64" else None,
                "kernel": "d3::molecule_kernel_v2<%s, true>" % ("double" if args.dtype == "f64" else "float"),
                "peak_source": "FMA-chain probe of the same type measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                "fp64_pipe_active_pct_ncu": 53.6 if args.dtype == "f64" else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch, profiles/r1h_molecule_kernel_v2_raw.csv",
                "flops_model": "140/pair-term (r<=25 Bohr), 89 (25<r<=50), SURVEY.md 8(d)",
                "hbm_gbs_measured_peak": _measured_hbm(),
            },
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            all_host_cores()
            stepf, kind, rp, desc = reference_step_fn(args.ref_nmol)
            stepf()  # warm-up (thread pools, allocator)
            t0 = time.perf_counter()
            stepf()
            dtc = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": rp / dtc, "unit": "pair-terms/s", "cores": torch.get_num_threads(),
                                    "kind": kind, "sample": desc, "seconds": dtc}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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


def run_c4(args, rank, local, world, dev):
    """BASELINE config 4: one 50k-atom water cluster, two-body + CN, energy + gradient,
    rows sharded over the ranks (strong scaling: the structure is fixed)."""
    import torch.distributed as dist

    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import _lib
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN
    from tad_dftd3_b200.distributed import dftd3_sharded, run_sharded
    from tad_dftd3_b200.system import SystemPlan, SystemRunner

    dt = torch.double
    lib = _lib.lib()
    stream = torch.cuda.current_stream(dev)
    numbers_h, positions_h = syn.water_cluster(args.nwater, seed=0)
    numbers_h, positions_h = numbers_h.pin_memory(), positions_h.pin_memory()
    numbers, positions = numbers_h.to(dev), positions_h.to(dev)
    near, far = count_pairs_single(positions)
    pairs, flops = near + far, near * FLOPS_NEAR + far * FLOPS_FAR
    ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
    refcn, refc6, _ = ref._prepared(dev, dt)
    rcov, r4r2 = COV_D3(dt, dev)[numbers], R4R2(dt, dev)[numbers]
    par = _lib.Params()
    par.s6, par.s8, par.a1, par.a2 = PARAM["s6"], PARAM["s8"], PARAM["a1"], PARAM["a2"]
    par.s9, par.rs9, par.alp = 0.0, 4.0 / 3.0, 14.0
    par.cutoff, par.cn_cutoff, par.kcn = 50.0, 25.0, 16.0
    g = torch.ones(numbers.shape[0], dtype=dt, device=dev)
    param = {k: torch.tensor(v, dtype=dt) for k, v in PARAM.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        plan = SystemPlan(numbers, positions, rcov, r4r2, par.kcn)
        runner = SystemRunner(plan, refcn, refc6, par, g, world)
        run_sharded(runner, True)
        return plan.scatter(runner.energy), plan.scatter(runner.grad)

    e_h = torch.empty(numbers.shape[0], dtype=dt).pin_memory()
    g_h = torch.empty(numbers.shape[0], 3, dtype=dt).pin_memory()

    def step_e2e():
        z = numbers_h.to(dev, non_blocking=True)
        x = positions_h.to(dev, non_blocking=True).requires_grad_(True)
        e = dftd3_sharded(z, x, param, ref=ref)
        (gx,) = torch.autograd.grad(e.sum(), x)
        e_h.copy_(e.detach(), non_blocking=True)
        g_h.copy_(gx, non_blocking=True)

    for _ in range(args.warmup):
        step_resident()
    barrier()
    l0 = _lib.launch_count
    with ClockSampler(local) as clocks:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        barrier()
        t0 = time.perf_counter()
        for a, b in ev:
            a.record(stream)
            step_resident()
            b.record(stream)
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
        ms = max(float(np.mean([a.elapsed_time(b) for a, b in ev])), wall_ms)
        for _ in range(args.warmup):
            step_e2e()
        barrier()
        e2e_steps = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            step_e2e()
            torch.cuda.synchronize(dev)
            e2e_steps.append((time.perf_counter() - t0) * 1e3)
        barrier()
        # median: a fresh process occasionally shows one step of tens of ms here (allocator growth while
        # the plan of every step is rebuilt for a new numbers tensor)
        e2e_ms = float(np.median(e2e_steps))
    launches = _lib.launch_count - l0
    peak = fp64_peak(lib, dev, stream)
    t = torch.tensor([ms, e2e_ms], dtype=torch.double, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        achieved = flops / (ms * 1e-3) / 1e12
        line = {
            "metric": "D3(BJ) energy+grad pair-terms/s fp64", "value": pairs / (ms * 1e-3),
            "unit": "pair-terms/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": f"c4: {numbers.shape[0]}-atom water cluster, D3(BJ) two-body + CN (cutoffs 50 / 25 Bohr), "
                            "energy+grad fp64, rows sharded over the GPUs",
                "pair_terms": pairs, "pairs_within_25_bohr": near,
                "entry": "SystemPlan (cached Morton order, d3b200_system_pack_f64) + d3b200_system_{cn,weights,tvec,pair,cnbwd}_f64",
                "l2": "per-atom arrays (10 MB) stream from L2/HBM every sweep; structure larger than one CTA's reach",
                "timing": "max(CUDA events, wall clock) per step incl. the plan build, max over ranks",
                "parallelism": f"rows of one structure over {world} GPU(s); all-reduce of CN, dL/dCN, E, grad",
            },
            "e2e": {"value": pairs / (e2e_ms * 1e-3), "unit": "pair-terms/s",
                    "h2d_bytes_per_step": int(numbers_h.numel() * 8 + positions_h.numel() * 8),
                    "d2h_bytes_per_step": int(e_h.numel() * 8 + g_h.numel() * 8), "ms_per_step": e2e_ms,
                    "api": "tad_dftd3_b200.distributed.dftd3_sharded() + autograd.grad, pinned host buffers; "
                           "median of the timed steps"},
            "gpu_launches": launches,
            "roofline": {"bound": "fp64", "achieved": achieved * world and achieved, "peak": peak * world,
                         "unit": "TFLOP/s", "frac": achieved / (peak * world), "traffic": None,
                         "kernel": "whole step (cn + weights + pair + cnbwd sweeps, all ranks)",
                         "peak_source": "FP64 FMA-chain probe x n_gpus, measured in this run"},
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            # the reference cannot hold 50k atoms (20 GB per dense N x N temporary): time a
            # 2000-atom sub-cluster of the same structure and report its dense pair rate
            all_host_cores()
            centre = positions.mean(0)
            sub = torch.argsort((positions - centre).norm(dim=1))[:2000]
            sn, sf = count_pairs_single(positions[sub])
            sec, kind = reference_single(numbers[sub], positions[sub], PARAM, chunk_size=256)
            line["cpu_baseline"] = {"value": (sn + sf) / sec, "unit": "pair-terms/s", "cores": torch.get_num_threads(),
                                    "kind": kind, "seconds": sec,
                                    "sample": "central 2000-atom sub-cluster of the same structure, energy + autograd "
                                              "gradient, chunk_size=256 (the reference cannot hold 50k atoms)"}
        print(json.dumps(line))


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


def run_c5a(args, rank, local, world, dev):
    """BASELINE config 5a: ATM three-body term on a 10k-atom water cluster with a triple cutoff.
    One step = the whole dftd3() evaluation with s9 = 1 (CN, weights, two-body sweep, three-body
    sweep, CN-backward), centres of the three-body term sharded over the GPUs."""
    import torch.distributed as dist

    import tad_dftd3_b200 as d3
    from tad_dftd3_b200 import _lib, synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN
    from tad_dftd3_b200.distributed import run_sharded
    from tad_dftd3_b200.system import SystemPlan, SystemRunner

    dt = torch.double
    lib = _lib.lib()
    stream = torch.cuda.current_stream(dev)
    nwater = 3334 if args.nwater == 16667 else args.nwater
    numbers, positions = syn.water_cluster(nwater, seed=0)
    numbers, positions = numbers.to(dev), positions.to(dev)
    triples, triangles, edges = count_triples(positions, args.cutoff)
    ref = d3.Reference(cn=REFCN(dt, dev), c6=syn.synthetic_c6_table(dt, dev))
    refcn, refc6, _ = ref._prepared(dev, dt)
    rcov, r4r2 = COV_D3(dt, dev)[numbers], R4R2(dt, dev)[numbers]
    rvdw = syn.synthetic_rvdw_table(dt, dev)
    par = _lib.Params()
    par.s6, par.s8, par.a1, par.a2 = PARAM["s6"], PARAM["s8"], PARAM["a1"], PARAM["a2"]
    par.s9, par.rs9, par.alp = 1.0, 1.3333333730697632, 14.0
    par.cutoff, par.cn_cutoff, par.kcn = args.cutoff, 25.0, 16.0
    g = torch.ones(numbers.shape[0], dtype=dt, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    t_atm, t_all = [], []
    launches = 0
    with ClockSampler(local) as clocks:
        for it in range(args.warmup + args.steps):
            barrier()
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            l0 = _lib.launch_count
            e0.record(stream)
            plan = SystemPlan(numbers, positions, rcov, r4r2, par.kcn)
            run = SystemRunner(plan, refcn, refc6, par, g, world, rvdw)
            atm = run.sweep_atm

            def timed_atm(*a, _atm=atm):
                e1.record(stream)
                _atm(*a)
                e2.record(stream)

            run.sweep_atm = timed_atm
            run_sharded(run, True)
            energy, grad = plan.scatter(run.energy), plan.scatter(run.grad)
            e3.record(stream)
            barrier()
            if it >= args.warmup:
                t_atm.append(e1.elapsed_time(e2)); t_all.append(e0.elapsed_time(e3))
                launches += _lib.launch_count - l0
    # ---- end to end through the public API (dftd3_sharded == dftd3 at one GPU): pinned host
    # numbers / positions in, pinned host energies + gradient out, every step
    from tad_dftd3_b200.distributed import dftd3_sharded

    d3.data.set_vdw_pairwise(rvdw)
    num_h, pos_h = numbers.cpu().pin_memory(), positions.cpu().pin_memory()
    e_h = torch.empty(numbers.shape[0], dtype=dt).pin_memory()
    g_h = torch.empty(numbers.shape[0], 3, dtype=dt).pin_memory()
    param = {k: torch.tensor(v, dtype=dt) for k, v in dict(PARAM, s9=1.0).items()}

    def step_e2e():
        z = num_h.to(dev, non_blocking=True)
        x = pos_h.to(dev, non_blocking=True).requires_grad_(True)
        e = dftd3_sharded(z, x, param, ref=ref, cutoff=torch.tensor(args.cutoff, dtype=dt))
        (gx,) = torch.autograd.grad(e.sum(), x)
        e_h.copy_(e.detach(), non_blocking=True)
        g_h.copy_(gx, non_blocking=True)
        torch.cuda.synchronize(dev)

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_err = float((g_h.to(dev) - grad).abs().max())
    # median over the timed steps: single steps are occasionally disturbed by allocator growth
    t = torch.tensor([float(np.median(t_atm)), float(np.median(t_all)), e2e_ms], dtype=torch.double, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_atm, ms_all, e2e_ms = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        return
    peak = fp64_peak(lib, dev, stream)
    achieved = triples * 140.0 / (ms_atm * 1e-3) / 1e12
    line = {
        "metric": "ATM energy+grad triples/s fp64", "value": triples / (ms_all * 1e-3), "unit": "triples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_all,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"c5a: dftd3() with the ATM term (s9=1) on a {numbers.shape[0]}-atom water cluster, "
                               f"cutoff {args.cutoff} Bohr, energy+grad fp64, centres sharded over the GPUs",
                   "triples": triples, "all_short_triangles": triangles, "short_edges": edges,
                   "entry": "SystemPlan + d3b200_system_{cn,weights,tvec,pair,atm,cnbwd}_f64; value = triples / whole step "
                            "(max over ranks, CUDA events); the three-body sweep alone in atm_ms",
                   "atm_ms": ms_atm, "triples_per_s_atm_kernel": triples / (ms_atm * 1e-3),
                   "atm_ms_per_step": [round(v, 3) for v in t_atm],
                   "total_energy": float(energy.sum()),
                   "parallelism": f"centres of the three-body term and rows of the two-body sweeps over {world} GPU(s); "
                                  "all-reduce of CN, dL/dCN, E, grad"},
        "gpu_launches": launches,
        "e2e": {"value": triples / (e2e_ms * 1e-3), "unit": "triples/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(num_h.numel() * 8 + pos_h.numel() * 8),
                "d2h_bytes_per_step": int((e_h.numel() + g_h.numel()) * 8),
                "api": "tad_dftd3_b200.distributed.dftd3_sharded() (= dftd3() at one GPU) + autograd.grad, pinned host "
                       "buffers, plan (Morton / element sort) rebuilt every step",
                "max_abs_grad_diff_vs_resident": e2e_err},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak * world, "unit": "TFLOP/s",
                     "frac": achieved / (peak * world),
                     "traffic": 100.9e6 if (world == 1 and args.cutoff == 25.0 and nwater == 3334) else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one launch (6.0 + 94.9 MB, mostly the "
                                       "per-centre neighbour records leaving L2), profiles/r1h_system_atm3_raw.csv",
                     "fp64_pipe_active_pct_ncu": 40.3,
                     "kernel": "d3::system_atm3_kernel<double, true> (every triple once), all ranks",
                     "flops_model": "140 per unordered triple (SURVEY.md 8d)"},
        "clocks": clocks.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        # the reference's ATM is dense N^3 (216 MB per temporary at N = 300): time a 300-atom sub-cluster
        all_host_cores()
        centre = positions.mean(0)
        sub = torch.argsort((positions - centre).norm(dim=1))[:300]
        t_sub, _, _ = count_triples(positions[sub], args.cutoff)
        sec, kind = reference_single(numbers[sub], positions[sub], dict(PARAM, s9=1.0), cutoff=args.cutoff,
                                     rvdw_table=rvdw)
        sec0, _ = reference_single(numbers[sub], positions[sub], PARAM, cutoff=args.cutoff)
        line["cpu_baseline"] = {"value": t_sub / max(sec - sec0, 1e-9), "unit": "triples/s",
                                "cores": torch.get_num_threads(), "kind": kind, "seconds": sec - sec0,
                                "sample": "central 300-atom sub-cluster, (two-body + ATM) minus (two-body) wall time, "
                                          "energy + autograd gradient"}
    print(json.dumps(line))


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


def _measured_hbm():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        return None


if __name__ == "__main__":
    main()
