"""
Generate the committed golden fixtures under `tests/golden/`.

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs
`/root/reference`); the fixtures travel to the GPU box, the reference does not.

1. `reference_pins.npz` -- golden values held by the reference's OWN tests
   (`test/test_model/samples.py`, `test/test_disp/samples.py`): cn, weights,
   dense c6, disp2, disp3 per molecule, plus the geometries that exist in the
   reference tree (`examples/batch.py`).  These pin the oracle (and the CUDA
   path) independently of any stand-in table where possible.
2. `reference_runs.npz` -- outputs of the UNMODIFIED reference source
   (`/root/reference/src/tad_dftd3`, imported over the tad-mctc stand-in) on
   seeded inputs with the closed-form synthetic C6 / r0ab tables of
   `tad_dftd3_b200.synthetic`: energies, gradients for random cotangents,
   parameter gradients, ATM, Hessians, edge cases.

    python oracle/make_golden.py
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from refimport import import_reference  # noqa: E402

d3 = import_reference()
from tad_dftd3.reference import Reference, _load_cn  # noqa: E402

from tad_dftd3_b200 import synthetic as syn  # noqa: E402
from tad_dftd3_b200.data import COV_D3  # noqa: E402

REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


# ---------------------------------------------------------------- part 1
def load_samples(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.refs


def example_geometries():
    """sample1/sample2 dict literals of examples/batch.py, evaluated safely."""
    src = open(os.path.join(REF, "examples", "batch.py")).read()
    tree = ast.parse(src)
    out = {}
    for node in tree.body:
        if isinstance(node, ast.Assign) and getattr(node.targets[0], "id", "") in ("sample1", "sample2"):
            kw = {k.arg: k.value for k in node.value.keywords}
            symbols = ast.literal_eval(kw["numbers"].args[0].func.value)  # "Pb H ...".split()
            from tad_mctc.data.pse import S2Z

            numbers = [S2Z[s] for s in symbols.split()]
            pos = ast.literal_eval(kw["positions"].args[0])
            out[node.targets[0].id] = (numbers, pos)
    return out


def part1():
    model = load_samples(os.path.join(REF, "test/test_model/samples.py"), "ref_model_samples")
    disp = load_samples(os.path.join(REF, "test/test_disp/samples.py"), "ref_disp_samples")
    geo = example_geometries()
    # our literal copies of the README geometries must be the reference's
    for (numbers, pos), mine in zip((geo["sample1"], geo["sample2"]), syn.README_BATCH):
        assert numbers == mine["numbers"] and pos == mine["positions"]
    known_numbers = {
        "SiH4": [14, 1, 1, 1, 1],
        "AmF3": [95, 9, 9, 9],
        "MB16_43_01": [11, 1, 8, 1, 9, 1, 1, 8, 7, 1, 1, 17, 5, 5, 7, 13],
        "PbH4-BiH3": geo["sample1"][0],
        "C6H5I-CH3SH": geo["sample2"][0],
    }
    known_pos = {"PbH4-BiH3": geo["sample1"][1], "C6H5I-CH3SH": geo["sample2"][1]}
    store = {}
    for name, numbers in known_numbers.items():
        key = name.replace("-", "_").replace("@", "_")
        store[f"{key}/numbers"] = np.asarray(numbers, dtype=np.int64)
        if name in known_pos:
            store[f"{key}/positions"] = np.asarray(known_pos[name], dtype=np.float64)
        for src in (model, disp):
            if name in src:
                for field, val in src[name].items():
                    if isinstance(val, torch.Tensor) and val.numel() > 0:
                        store[f"{key}/{field}"] = val.double().numpy()
    np.savez_compressed(os.path.join(OUT, "reference_pins.npz"), **store)
    print("reference_pins.npz:", sorted(store))


# ---------------------------------------------------------------- part 2
def ref_tables(dtype):
    refobj = Reference(cn=_load_cn(dtype), c6=syn.synthetic_c6_table(dtype))
    return refobj, COV_D3(dtype), d3.data.R4R2(dtype), syn.synthetic_rvdw_table(dtype)


def run_reference(numbers, positions, param, dtype, cutoff=None, g=None, want_param_grad=False,
                  want_hessian=False, chunk_size=None):
    """Energy (+ VJP for cotangent g, + parameter gradients, + Hessian of the sum)."""
    refobj, rcov_t, r4r2_t, rvdw_t = ref_tables(dtype)
    pos = positions.to(dtype).clone().requires_grad_(True)
    par = {k: torch.tensor(v, dtype=dtype, requires_grad=want_param_grad and k in ("s6", "s8", "a1", "a2"))
           for k, v in param.items()}
    kw = dict(
        ref=refobj,
        rcov=rcov_t[numbers],
        rvdw=rvdw_t[numbers.unsqueeze(-1), numbers.unsqueeze(-2)],
        r4r2=r4r2_t[numbers],
        chunk_size=chunk_size,
    )
    if cutoff is not None:
        kw["cutoff"] = torch.tensor(cutoff, dtype=dtype)
    e = d3.dftd3(numbers, pos, par, **kw)
    out = {"energy": e.detach().numpy()}
    if g is not None:
        gg = g.to(dtype)
        wrt = [pos] + ([par[k] for k in ("s6", "s8", "a1", "a2") if k in par] if want_param_grad else [])
        grads = torch.autograd.grad((e * gg).sum(), wrt)
        out["grad"] = grads[0].numpy()
        if want_param_grad:
            out["pgrad"] = np.asarray([float(x) for x in grads[1:]])
    if want_hessian:
        def f(p):
            return d3.dftd3(numbers, p, {k: v.detach() for k, v in par.items()}, **kw).sum(-1)
        h = torch.func.jacrev(torch.func.jacrev(f))(positions.to(dtype))
        out["hessian"] = h.detach().numpy()
    return out


def tens(sample, dtype=torch.double):
    return torch.tensor(sample["numbers"]), torch.tensor(sample["positions"], dtype=dtype)


def crowded_cluster():
    """He + 30 H on a 1.7 Bohr sphere + far Ar: every Gaussian weight underflows."""
    rng = np.random.RandomState(7)
    u = rng.normal(size=(30, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    pos = np.concatenate([np.zeros((1, 3)), 1.7 * u, np.array([[9.0, 0.5, -0.3]])])
    numbers = np.array([2] + [1] * 30 + [18])
    return torch.from_numpy(numbers), torch.from_numpy(pos)


PAR_R2SCAN = {"a1": 0.49484001, "s8": 0.78981345, "a2": 5.73083694}
PAR_TPSS0 = {"s6": 1.0, "s8": 1.2576, "s9": 0.0, "alp": 14.0, "a1": 0.3768, "a2": 4.5865}
PAR_TPSS0_ATM = dict(PAR_TPSS0, s9=1.0)
PAR_GFN1 = {"s6": 1.0, "s8": 2.4, "s9": 0.0, "a1": 0.63, "a2": 5.0}
PAR_FULL = {"s6": 0.93, "s8": 0.78981345, "s9": 1.0, "a1": 0.49484001, "a2": 5.73083694}


def part2():
    store = {}
    rng = np.random.RandomState(11)

    def add(case, numbers, positions, param, dtype=torch.double, **kw):
        g = torch.from_numpy(rng.uniform(0.5, 1.5, size=tuple(numbers.shape)))
        res = run_reference(numbers, positions, param, dtype, g=g, **kw)
        store[f"{case}/numbers"] = numbers.numpy()
        store[f"{case}/positions"] = positions.double().numpy()
        store[f"{case}/g"] = g.numpy()
        store[f"{case}/param_keys"] = np.asarray(list(param.keys()))
        store[f"{case}/param_vals"] = np.asarray(list(param.values()), dtype=np.float64)
        store[f"{case}/cutoff"] = np.asarray(kw.get("cutoff") or 50.0)
        store[f"{case}/dtype"] = np.asarray(str(dtype))
        for k, v in res.items():
            store[f"{case}/{k}"] = v
        print(f"  {case}: E_tot = {res['energy'].sum():.12e}")

    # config 1: README single structure (fp64 and the README's own fp32)
    n1, p1 = tens(syn.README_SINGLE)
    add("c1_f64", n1, p1, PAR_R2SCAN, want_param_grad=True)
    add("c1_f32", n1, p1, PAR_R2SCAN, dtype=torch.float32)
    add("c1_atm_f64", n1, p1, PAR_FULL, want_param_grad=True)
    # config 2: README batch
    na, pa = tens(syn.README_BATCH[0])
    nb, pb = tens(syn.README_BATCH[1])
    n2, p2 = syn.pack([na, nb]), syn.pack([pa, pb])
    add("c2_f64", n2, p2, PAR_R2SCAN, want_param_grad=True)
    add("c2_f32", n2, p2, PAR_R2SCAN, dtype=torch.float32)
    add("c2_tpss0_f64", n2, p2, PAR_TPSS0)
    add("c2_atm_f64", n2, p2, PAR_TPSS0_ATM, want_param_grad=True)
    add("c2_atm_cut8_f64", n2, p2, PAR_TPSS0_ATM, cutoff=8.0)
    add("c2_cut6_f64", n2, p2, PAR_TPSS0, cutoff=6.0)
    add("c2_atm_f32", n2, p2, PAR_TPSS0_ATM, dtype=torch.float32)
    # Hessians (two-body + CN, and with ATM) on PbH4-BiH3
    add("pbh4_hess_f64", na, pa, PAR_GFN1, want_hessian=True)
    add("pbh4_hess_atm_f64", na, pa, PAR_FULL, want_hessian=True)
    add("pbh4_chunk2_f64", na, pa, PAR_GFN1, chunk_size=2)
    # small chain-grown batch, ragged sizes 5..40
    n3, p3 = syn.drug_like_batch(6, 5, 40, seed=3)
    add("mols6_f64", n3, p3, PAR_R2SCAN, want_param_grad=True)
    add("mols6_atm_f64", n3, p3, PAR_FULL, want_param_grad=True)
    add("mols6_atm_cut12_f64", n3, p3, PAR_FULL, cutoff=12.0)
    add("mols6_f32", n3, p3, PAR_R2SCAN, dtype=torch.float32)
    # ghost atoms (Z = 0 in the middle of a structure), as in the disp.py docstring
    n4 = n3[1].clone()
    n4[[1, 4, 7]] = 0
    add("ghost_f64", n4, p3[1], PAR_FULL)
    # all-Gaussians-underflow fallback of the weights
    n5, p5 = crowded_cluster()
    add("crowded_f64", n5, p5, PAR_TPSS0)
    # close contacts
    n6, p6 = tens(syn.NAN_SAMPLE)
    add("close_f64", n6, p6, PAR_FULL)
    add("close_f32", n6, p6, PAR_FULL, dtype=torch.float32)
    # degenerate sizes
    add("one_atom_f64", torch.tensor([6]), torch.zeros(1, 3, dtype=torch.double), PAR_FULL)
    add("two_atoms_f64", torch.tensor([3, 1]), torch.tensor([[0.0, 0, -1.5], [0.0, 0, 1.5]], dtype=torch.double), PAR_FULL,
        want_hessian=True)
    # a water cluster that spans the CN cutoff (25) but not the dispersion cutoff
    n7, p7 = syn.water_cluster(64, seed=5, spacing=9.0)
    add("water64_f64", n7, p7, PAR_TPSS0)
    add("water64_cut20_atm_f64", n7[:90], p7[:90], PAR_TPSS0_ATM, cutoff=20.0)

    # stage outputs for the stage-level drop-ins (README.rst:238-261)
    refobj, rcov_t, r4r2_t, rvdw_t = ref_tables(torch.double)
    cn = d3.ncoord.cn_d3(n2, p2, counting_function=d3.ncoord.exp_count, rcov=rcov_t[n2])
    w = d3.model.weight_references(n2, cn, refobj)
    c6 = d3.model.atomic_c6(n2, w, refobj)
    e2 = d3.disp.dispersion(n2, p2, {k: torch.tensor(v, dtype=torch.double) for k, v in PAR_TPSS0_ATM.items()}, c6,
                            rvdw_t[n2.unsqueeze(-1), n2.unsqueeze(-2)], r4r2_t[n2])
    store["stages_c2/cn"] = cn.numpy()
    store["stages_c2/weights"] = w.numpy()
    store["stages_c2/c6"] = c6.numpy()
    store["stages_c2/energy"] = e2.numpy()

    np.savez_compressed(os.path.join(OUT, "reference_runs.npz"), **store)
    sz = os.path.getsize(os.path.join(OUT, "reference_runs.npz"))
    print(f"reference_runs.npz: {len(store)} arrays, {sz/1024:.0f} kB")


# ---------------------------------------------------------------- part 3
def heavy_cluster(natoms=36, nelem=12, seed=21, box=13.0, dmin=2.3):
    """Random cluster with `nelem` distinct elements up to uranium (more than the 8 that the
    pre-contracted T-vector sweeps take: exercises the general tiled kernels)."""
    rng = np.random.RandomState(seed)
    pool = [1, 6, 7, 8, 9, 15, 16, 17, 26, 35, 53, 79, 82, 92]
    zs = rng.choice(pool, size=nelem, replace=False)
    numbers = np.concatenate([zs, rng.choice(zs, size=natoms - nelem)])
    pos = []
    while len(pos) < natoms:
        c = rng.uniform(0.0, box, size=3)
        if all(np.linalg.norm(c - q) >= dmin for q in pos):
            pos.append(c)
    return torch.from_numpy(numbers.astype(np.int64)), torch.from_numpy(np.asarray(pos))


def part3():
    """Second fixture file (kept separate so that `reference_runs.npz` stays byte-identical)."""
    store = {}
    rng = np.random.RandomState(23)

    def add(case, numbers, positions, param, dtype=torch.double, **kw):
        g = torch.from_numpy(rng.uniform(0.5, 1.5, size=tuple(numbers.shape)))
        res = run_reference(numbers, positions, param, dtype, g=g, **kw)
        store[f"{case}/numbers"] = numbers.numpy()
        store[f"{case}/positions"] = positions.double().numpy()
        store[f"{case}/g"] = g.numpy()
        store[f"{case}/param_keys"] = np.asarray(list(param.keys()))
        store[f"{case}/param_vals"] = np.asarray(list(param.values()), dtype=np.float64)
        store[f"{case}/cutoff"] = np.asarray(kw.get("cutoff") or 50.0)
        store[f"{case}/dtype"] = np.asarray(str(dtype))
        for k, v in res.items():
            store[f"{case}/{k}"] = v
        print(f"  {case}: E_tot = {res['energy'].sum():.12e}")

    nh, ph = heavy_cluster()
    assert len(set(nh.tolist())) == 12
    add("heavy12_f64", nh, ph, PAR_R2SCAN, want_param_grad=True)
    add("heavy12_atm_f64", nh, ph, PAR_FULL, want_param_grad=True)
    add("heavy12_atm_cut9_f64", nh, ph, PAR_FULL, cutoff=9.0)
    add("heavy12_f32", nh, ph, PAR_R2SCAN, dtype=torch.float32)
    add("heavy12_hess_atm_f64", nh[:10], ph[:10], PAR_FULL, want_hessian=True)
    # ragged batch: one atom, forty atoms, three atoms, and a row of padding only
    n40, p40 = syn.drug_like_batch(1, 40, 40, seed=9)
    na, pa = tens(syn.README_BATCH[0])
    rows_n = [torch.tensor([8]), n40[0], na[:3], torch.zeros(2, dtype=torch.int64)]
    rows_p = [torch.zeros(1, 3, dtype=torch.double), p40[0], pa[:3], torch.zeros(2, 3, dtype=torch.double)]
    add("ragged_atm_f64", syn.pack(rows_n), syn.pack(rows_p), PAR_TPSS0_ATM, want_param_grad=True)
    # 300-atom water cluster, three-body term with a cutoff that cuts many triples partly: neighbour
    # lists of 100-250 atoms, i.e. several 64-entry tiles per centre in the triples-once kernel
    nw, pw = syn.water_cluster(100, seed=13)
    add("water100_atm_cut18_f64", nw, pw, PAR_TPSS0_ATM, cutoff=18.0)
    # every damping parameter left at the reference's defaults (defaults.py:41-60), without and with s9
    n3, p3 = syn.drug_like_batch(4, 6, 25, seed=17)
    add("defaults_f64", n3, p3, {})
    add("defaults_s9_f64", n3, p3, {"s9": 1.0})
    # stage outputs on the 12-element cluster (stage kernels on heavy elements)
    refobj, rcov_t, r4r2_t, rvdw_t = ref_tables(torch.double)
    cn = d3.ncoord.cn_d3(nh, ph, counting_function=d3.ncoord.exp_count, rcov=rcov_t[nh])
    w = d3.model.weight_references(nh, cn, refobj)
    c6 = d3.model.atomic_c6(nh, w, refobj)
    e2 = d3.disp.dispersion(nh, ph, {k: torch.tensor(v, dtype=torch.double) for k, v in PAR_FULL.items()}, c6,
                            rvdw_t[nh.unsqueeze(-1), nh.unsqueeze(-2)], r4r2_t[nh])
    store["stages_heavy12/cn"] = cn.numpy()
    store["stages_heavy12/weights"] = w.numpy()
    store["stages_heavy12/c6"] = c6.numpy()
    store["stages_heavy12/energy"] = e2.numpy()
    np.savez_compressed(os.path.join(OUT, "reference_runs2.npz"), **store)
    sz = os.path.getsize(os.path.join(OUT, "reference_runs2.npz"))
    print(f"reference_runs2.npz: {len(store)} arrays, {sz/1024:.0f} kB")


if __name__ == "__main__":
    torch.manual_seed(0)
    if "--only-part3" not in sys.argv:
        part1()
        part2()
    part3()
