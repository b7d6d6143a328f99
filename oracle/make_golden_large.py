"""
Golden fixtures at the largest sizes the UNMODIFIED reference can hold in this
container (`tests/golden/reference_large.npz`): they pin the blocked C oracle
(`oracle/d3_blocked.c`) -- the checker of the full-size GPU tests -- against the
reference's own energies and autograd gradients.

TEST INFRASTRUCTURE ONLY.  Needs `/root/reference`; run once in the build
container:   python oracle/make_golden_large.py

Cases (fp64, synthetic C6 / r0ab tables of `tad_dftd3_b200.synthetic`, random
cotangents g so that every per-atom energy is pinned individually):
  water2001    667 rigid waters (2001 atoms), two-body + CN, default cutoffs 50 / 25 Bohr
  heavy600     600-atom random cluster, 12 elements up to uranium, 8 ghost atoms (Z = 0) interleaved
  water300_atm 100 waters, s9 = 1, cutoff 12 Bohr (most triples partly cut: ordered rule atm.py:147-149)
  heavy150_atm 150 atoms, 12 elements, s9 = 1, cutoff 9 Bohr, two ghost atoms
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

from make_golden import heavy_cluster, run_reference  # noqa: E402  (imports the reference over the stand-in)

from tad_dftd3_b200 import synthetic as syn  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_large.npz")
PARAM = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}


def with_ghosts(numbers, positions, every):
    """Insert a ghost atom (Z = 0, position of its predecessor shifted) after every `every`-th atom."""
    zs, xs = [], []
    for k in range(numbers.shape[0]):
        zs.append(int(numbers[k]))
        xs.append(positions[k].tolist())
        if (k + 1) % every == 0:
            zs.append(0)
            xs.append((positions[k] + 0.37).tolist())
    return torch.tensor(zs), torch.tensor(xs, dtype=torch.double)


def main():
    torch.manual_seed(0)
    store = {}

    def add(case, numbers, positions, param, cutoff=None, chunk_size=None):
        gen = torch.Generator().manual_seed(len(store) + 11)
        g = torch.rand(numbers.shape, dtype=torch.double, generator=gen) + 0.5
        out = run_reference(numbers, positions, param, torch.double, cutoff=cutoff, g=g, chunk_size=chunk_size)
        store[f"{case}/numbers"] = numbers.numpy().astype(np.int64)
        store[f"{case}/positions"] = positions.numpy()
        store[f"{case}/g"] = g.numpy()
        store[f"{case}/param_keys"] = np.asarray(list(param.keys()))
        store[f"{case}/param_vals"] = np.asarray([float(v) for v in param.values()])
        store[f"{case}/cutoff"] = np.asarray(50.0 if cutoff is None else cutoff)
        store[f"{case}/energy"] = out["energy"]
        store[f"{case}/grad"] = out["grad"]
        print(case, numbers.shape[0], "atoms, E =", float(out["energy"].sum()))

    z, x = syn.water_cluster(667, seed=5)
    add("water2001", z, x, PARAM, chunk_size=256)
    z, x = heavy_cluster(natoms=600, nelem=12, seed=33, box=34.0, dmin=2.4)
    z, x = with_ghosts(z, x, 75)
    add("heavy600", z, x, PARAM, chunk_size=128)
    z, x = syn.water_cluster(100, seed=6)
    add("water300_atm", z, x, dict(PARAM, s9=1.0), cutoff=12.0)
    z, x = heavy_cluster(natoms=150, nelem=12, seed=34, box=21.0, dmin=2.4)
    z, x = with_ghosts(z, x, 75)
    add("heavy150_atm", z, x, dict(PARAM, s9=1.0), cutoff=9.0)
    np.savez_compressed(OUT, **store)
    print(OUT, os.path.getsize(OUT) // 1024, "kB")


if __name__ == "__main__":
    main()
