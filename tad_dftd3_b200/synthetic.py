"""
Synthetic tables and structures for parity tests and benchmarks.

Neither the reference-C6 table (`reference-c6.pt`, a missing large blob of the
reference checkout) nor the D3 r0ab van-der-Waals table (lives in tad-mctc) is
available offline, and there is no network for molecule data sets.  Everything
here is closed-form or seeded with numpy's `RandomState` (bit-stable across
numpy versions), so the build container, the GPU box and the committed golden
fixtures all see identical inputs.

All of it enters the product only through the public keyword arguments
(`ref=`, `rvdw=`) exactly as a user-supplied table would.
"""
from __future__ import annotations

import numpy as np
import torch

from .data import COV_2009_ANGSTROM, ANGSTROM_TO_BOHR

__all__ = [
    "synthetic_c6_table",
    "synthetic_rvdw_table",
    "drug_like_batch",
    "water_cluster",
    "README_SINGLE",
    "README_BATCH",
    "NAN_SAMPLE",
]

MAX_ELEMENT = 104


def synthetic_c6_table(dtype=torch.double, device=None, refcn=None) -> torch.Tensor:
    """
    Closed-form stand-in for the `[104, 104, 7, 7]` reference-C6 table.

    K[z1,z2,a,b] = sqrt(c[z1,a] c[z2,b]) (1 + 0.05 s[z1,a] s[z2,b]) with
    c[z,a] = (1.5 + 0.9 z^1.3)(1 - 0.07 a) and s = sin(1.7 z + 2.3 a + 0.5):
    positive, of realistic magnitude (H-H ~ 2.4, C-C ~ 10..100), and symmetric
    under (z1,a) <-> (z2,b) like the real table.  Row/column 0 (padding) is 0.
    """
    z = np.arange(MAX_ELEMENT, dtype=np.float64)[:, None]
    a = np.arange(7, dtype=np.float64)[None, :]
    c = (1.5 + 0.9 * z**1.3) * (1.0 - 0.07 * a)
    s = np.sin(1.7 * z + 2.3 * a + 0.5)
    k = np.sqrt(c[:, None, :, None] * c[None, :, None, :]) * (
        1.0 + 0.05 * s[:, None, :, None] * s[None, :, None, :]
    )
    k = 0.5 * (k + k.transpose(1, 0, 3, 2))  # exactly symmetric in floating point
    k[0] = 0.0
    k[:, 0] = 0.0
    return torch.from_numpy(k).to(dtype=dtype, device=device)


def synthetic_rvdw_table(dtype=torch.double, device=None) -> torch.Tensor:
    """Closed-form symmetric stand-in for the `[104, 104]` r0ab table (Bohr)."""
    r = np.asarray(COV_2009_ANGSTROM[:MAX_ELEMENT], dtype=np.float64) * ANGSTROM_TO_BOHR
    z = np.arange(MAX_ELEMENT, dtype=np.float64)
    wob = 0.05 * np.sin(0.37 * z[:, None] * z[None, :] + 0.11 * (z[:, None] + z[None, :]))
    t = 1.15 * (r[:, None] + r[None, :]) + 2.1 + wob
    t[0, :] = 0.0
    t[:, 0] = 0.0
    return torch.from_numpy(t).to(dtype=dtype, device=device)


# --------------------------------------------------------------------------
# structure generators
# --------------------------------------------------------------------------

_ELEMS = np.array([1, 6, 7, 8, 16, 9, 17])
_PROBS = np.array([0.45, 0.35, 0.07, 0.08, 0.02, 0.015, 0.015])


def drug_like_batch(nmol: int, nmin: int = 50, nmax: int = 120, seed: int = 0):
    """
    Zero-padded batch of chain-grown organic-like molecules (BASELINE config 3).

    Heavy atoms first (each bonded at the sum of covalent radii to a random
    earlier heavy atom), hydrogens last; a trial position is rejected when it
    is closer than 0.8 x (bond sum) or 1.3 Bohr to an existing atom.  All
    molecules grow in lock-step, so 4096 molecules take about a second.

    Returns `(numbers [B, nmax] int64, positions [B, nmax, 3] float64)`,
    each molecule centred at the origin, padding = 0 / 0.0.
    """
    rng = np.random.RandomState(seed)
    rad = np.asarray(COV_2009_ANGSTROM, dtype=np.float64) * ANGSTROM_TO_BOHR
    n = rng.randint(nmin, nmax + 1, size=nmol)
    nmax_real = int(n.max())
    numbers = np.zeros((nmol, nmax), dtype=np.int64)
    draw = rng.choice(len(_ELEMS), size=(nmol, nmax), p=_PROBS)
    z = _ELEMS[draw]
    z[:, 0] = 6  # at least one heavy atom to grow from
    real = np.arange(nmax)[None, :] < n[:, None]
    z = np.where(real, z, 0)
    # heavy atoms first, H afterwards, padding last (stable)
    key = np.where(z == 0, 2, np.where(z == 1, 1, 0))
    order = np.argsort(key, axis=1, kind="stable")
    numbers = np.take_along_axis(z, order, axis=1)
    nheavy = (key == 0).sum(1)

    pos = np.zeros((nmol, nmax, 3), dtype=np.float64)
    rz = rad[numbers]
    for t in range(1, nmax_real):
        todo = n > t
        tries = 0
        while todo.any():
            idx = np.nonzero(todo)[0]
            m = len(idx)
            hi = np.minimum(t, nheavy[idx])
            parent = (rng.random_sample(m) * hi).astype(np.int64)
            u = rng.normal(size=(m, 3))
            u /= np.linalg.norm(u, axis=1, keepdims=True)
            bond = rz[idx, parent] + rz[idx, t]
            trial = pos[idx, parent] + bond[:, None] * u
            d = np.linalg.norm(pos[idx, :t] - trial[:, None, :], axis=2)
            lim = np.maximum(0.8 * (rz[idx, :t] + rz[idx, t][:, None]), 1.3)
            # relax slowly if a molecule is badly crowded
            ok = (d >= lim * (1.0 if tries < 200 else 0.9)).all(1)
            acc = idx[ok]
            pos[acc, t] = trial[ok]
            todo[acc] = False
            tries += 1
            if tries > 2000:  # pragma: no cover
                raise RuntimeError("chain growth did not converge")
    realf = (numbers != 0)[..., None]
    centre = (pos * realf).sum(1, keepdims=True) / n[:, None, None]
    pos = np.where(realf, pos - centre, 0.0)
    return torch.from_numpy(numbers), torch.from_numpy(pos)


def water_cluster(nwater: int, seed: int = 0, spacing: float = 5.86, sigma: float = 0.3):
    """
    Rigid waters (O-H 1.81 Bohr, 104.5 deg, random orientation) on a jittered
    cubic lattice (BASELINE configs 4 and 5a).  Returns
    `(numbers [3 nwater] int64, positions [3 nwater, 3] float64)` centred at
    the origin, atoms ordered O,H,H per molecule.
    """
    rng = np.random.RandomState(seed)
    m = int(np.ceil(nwater ** (1.0 / 3.0) - 1e-9))
    grid = np.stack(np.meshgrid(*(np.arange(m),) * 3, indexing="ij"), -1).reshape(-1, 3)
    keep = np.sort(rng.permutation(len(grid))[:nwater])
    sites = grid[keep].astype(np.float64) * spacing + rng.normal(scale=sigma, size=(nwater, 3))
    # random rotations from normalised quaternions
    q = rng.normal(size=(nwater, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, zq = q.T
    rot = np.stack(
        [
            1 - 2 * (y * y + zq * zq), 2 * (x * y - zq * w), 2 * (x * zq + y * w),
            2 * (x * y + zq * w), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - x * w),
            2 * (x * zq - y * w), 2 * (y * zq + x * w), 1 - 2 * (x * x + y * y),
        ],
        -1,
    ).reshape(nwater, 3, 3)
    half = np.deg2rad(104.5) / 2
    oh = 1.81
    local = np.array(
        [
            [0.0, 0.0, 0.0],
            [oh * np.sin(half), 0.0, oh * np.cos(half)],
            [-oh * np.sin(half), 0.0, oh * np.cos(half)],
        ]
    )
    pos = sites[:, None, :] + np.einsum("nij,aj->nai", rot, local)
    pos = pos.reshape(-1, 3)
    pos -= pos.mean(0, keepdims=True)
    numbers = np.tile(np.array([8, 1, 1], dtype=np.int64), nwater)
    return torch.from_numpy(numbers), torch.from_numpy(pos)


# --------------------------------------------------------------------------
# literal README / example structures (BASELINE configs 1 and 2, Bohr)
# --------------------------------------------------------------------------

README_SINGLE = dict(  # C4H5NS, `examples/single.py:7-25` of the reference
    numbers=[6, 6, 6, 6, 7, 6, 16, 1, 1, 1, 1, 1],
    positions=[
        [-2.56745685564671, -0.02509985979910, 0.00000000000000],
        [-1.39177582455797, +2.27696188880014, 0.00000000000000],
        [+1.27784995624894, +2.45107479759386, 0.00000000000000],
        [+2.62801937615793, +0.25927727028120, 0.00000000000000],
        [+1.41097033661123, -1.99890996077412, 0.00000000000000],
        [-1.17186102298849, -2.34220576284180, 0.00000000000000],
        [-2.39505990368378, -5.22635838332362, 0.00000000000000],
        [+2.41961980455457, -3.62158019253045, 0.00000000000000],
        [-2.51744374846065, +3.98181713686746, 0.00000000000000],
        [+2.24269048384775, +4.24389473203647, 0.00000000000000],
        [+4.66488984573956, +0.17907568006409, 0.00000000000000],
        [-4.60044244782237, -0.17794734637413, 0.00000000000000],
    ],
)

README_BATCH = (  # PbH4-BiH3 and C6H5I-CH3SH, `examples/batch.py:7-49`
    dict(
        numbers=[82, 1, 1, 1, 1, 83, 1, 1, 1],
        positions=[
            [-0.00000020988889, -4.98043478877778, +0.00000000000000],
            [+3.06964045311111, -6.06324400177778, +0.00000000000000],
            [-1.53482054188889, -6.06324400177778, -2.65838526500000],
            [-1.53482054188889, -6.06324400177778, +2.65838526500000],
            [-0.00000020988889, -1.72196703577778, +0.00000000000000],
            [-0.00000020988889, +4.77334244722222, +0.00000000000000],
            [+1.35700257511111, +6.70626379422222, -2.35039772300000],
            [-2.71400388988889, +6.70626379422222, +0.00000000000000],
            [+1.35700257511111, +6.70626379422222, +2.35039772300000],
        ],
    ),
    dict(
        numbers=[6, 6, 6, 6, 6, 6, 53, 1, 1, 1, 1, 1, 16, 1, 6, 1, 1, 1],
        positions=[
            [-1.42754169820131, -1.50508961850828, -1.93430551124333],
            [+1.19860572924150, -1.66299114873979, -2.03189643761298],
            [+2.65876001301880, +0.37736955363609, -1.23426391650599],
            [+1.50963368042358, +2.57230374419743, -0.34128058818180],
            [-1.12092277855371, +2.71045691257517, -0.25246348639234],
            [-2.60071517756218, +0.67879949508239, -1.04550707592673],
            [-2.86169588073340, +5.99660765711210, +1.08394899986031],
            [+2.09930989272956, -3.36144811062374, -2.72237695164263],
            [+2.64405246349916, +4.15317840474646, +0.27856972788526],
            [+4.69864865613751, +0.26922271535391, -1.30274048619151],
            [-4.63786461351839, +0.79856258572808, -0.96906659938432],
            [-2.57447518692275, -3.08132039046931, -2.54875517521577],
            [-5.88211879210329, 11.88491819358157, +2.31866455902233],
            [-8.18022701418703, 10.95619984550779, +1.83940856333092],
            [-5.08172874482867, 12.66714386256482, -0.92419491629867],
            [-3.18311711399702, 13.44626574330220, -0.86977613647871],
            [-5.07177399637298, 10.99164969235585, -2.10739192258756],
            [-6.35955320518616, 14.08073002965080, -1.68204314084441],
        ],
    ),
)

NAN_SAMPLE = dict(  # close-contact regression geometry, `test/test_grad/test_nan.py:34-55`
    numbers=[6, 6, 6, 6, 6, 6, 6, 6, 1, 1, 1, 1, 1, 7, 8, 8, 8],
    positions=[
        [-1.0981, +0.1496, +0.1346],
        [-0.4155, +1.2768, +0.3967],
        [+0.9426, +0.7848, +0.1307],
        [+2.1708, +1.3814, -0.0347],
        [+3.3234, +0.5924, -0.1535],
        [+3.1564, -0.8110, -0.0285],
        [+1.8929, -1.4673, +0.0373],
        [+0.8498, -0.5613, +0.0109],
        [-0.7751, +2.2970, +0.5540],
        [+2.3079, +2.4725, -0.1905],
        [+4.3031, +0.9815, -0.4599],
        [+4.0011, -1.4666, -0.0514],
        [+1.8340, -2.5476, -0.1587],
        [-2.5629, -0.0306, -0.1458],
        [-3.0792, +1.0280, -0.3225],
        [-3.0526, -1.1594, +0.1038],
        [-0.4839, -0.9612, -0.0048],
    ],
)


def pack(tensors, value=0):
    """Zero-pad a sequence of tensors to a common shape and stack (tad-mctc `pack`)."""
    tensors = list(tensors)
    shape = [max(t.shape[d] for t in tensors) for d in range(tensors[0].dim())]
    out = torch.full((len(tensors), *shape), value, dtype=tensors[0].dtype, device=tensors[0].device)
    for n, t in enumerate(tensors):
        out[(n, *[slice(0, s) for s in t.shape])] = t
    return out
