"""
Periodic images for `dftd3()` -- SURVEY.md 8(f) rank 2, caller side of the hot path.

The reference has no periodic boundary conditions: its cutoffs act on dense
distance matrices of ONE finite structure (disp.py:276-295, damping/atm.py:111-149).
This front end keeps exactly those semantics -- hard cutoffs, no switching
function, the ordered ATM cutoff rule, CN with the fixed 25 Bohr cutoff of
`cn_d3` (disp.py:150-152) -- and applies them to the infinite lattice:

    E_i(cell) = E_i of atom i inside the finite cluster made of the unit cell and
                every lattice image that can influence the atoms of the cell.

Which images can: a pair term of a cell atom i needs the partners j within
`cutoff` and their coordination numbers, i.e. every atom within `cutoff + 25`;
an ATM term (i; j, k) has r_ij <= cutoff and r_jk <= cutoff (the mask never
tests r_ik, atm.py:147-149), so k lies within 2 cutoff of i and its C6 needs
the atoms within `2 cutoff + 25`.  Atoms beyond that radius change nothing
(the cutoffs are step functions), so the result does not depend on how much
more than the necessary shell is included.

The extended cluster goes through the SAME kernels as any other structure
(fused kernel up to 960 atoms, tiled Morton-sorted sweeps beyond: the box
culling of the tiles is the neighbour search).  Image positions are
`positions[index] + shifts @ cell`, built with torch ops, so autograd folds the
gradients of the images back onto the atoms of the cell and also yields
dE/dcell (the strain derivative).
"""
from __future__ import annotations

import math

import torch
from torch import Tensor

from . import defaults

__all__ = ["lattice_images", "dftd3_periodic", "image_radius"]


def image_radius(cutoff: float, atm: bool, cn_cutoff: float = defaults.D3_CN_CUTOFF) -> float:
    """Distance from the atoms of the cell up to which lattice images can change their energies."""
    return (2.0 * cutoff if atm else cutoff) + cn_cutoff


def lattice_images(numbers: Tensor, positions: Tensor, cell: Tensor, radius: float,
                   pbc=(True, True, True)):
    """Atoms of the extended cluster as (index [M] int64, shifts [M, 3] int64): extended atom m is atom
    `index[m]` of the cell translated by `shifts[m] @ cell`.  The N atoms of the cell come first, in order
    (padding atoms, Z = 0, included; they have no images); then every image of a real atom that lies within
    `radius` of some real atom of the cell.  Works on detached values; nothing here is differentiated."""
    if numbers.dim() != 1 or positions.shape != (numbers.shape[0], 3):
        raise ValueError("lattice_images takes one structure: numbers [N], positions [N, 3]")
    if cell.shape != (3, 3):
        raise ValueError("cell must be [3, 3] (rows are the lattice vectors)")
    dev = positions.device
    x = positions.detach().to(torch.double)
    a = cell.detach().to(device=dev, dtype=torch.double)
    n = numbers.shape[0]
    real = numbers.to(dev) != 0
    vol = float(torch.det(a).abs())
    if any(pbc) and not vol > 0.0:
        raise ValueError("cell is singular")
    central = torch.arange(n, device=dev)
    zero = torch.zeros(n, 3, dtype=torch.int64, device=dev)
    if not any(pbc) or int(real.sum()) == 0:
        return central, zero
    xr = x[real]
    frac = xr @ torch.linalg.inv(a)
    span = (frac.max(0).values - frac.min(0).values).tolist()
    nmax = []
    for k in range(3):
        if not pbc[k]:
            nmax.append(0)
            continue
        l, m = (k + 1) % 3, (k + 2) % 3
        height = vol / float(torch.linalg.cross(a[l], a[m]).norm())
        nmax.append(int(math.ceil(radius / height + span[k])))
    rng = [torch.arange(-m, m + 1, device=dev) for m in nmax]
    grid = torch.stack(torch.meshgrid(*rng, indexing="ij"), -1).reshape(-1, 3)
    grid = grid[(grid != 0).any(-1)]                             # the cell itself is already there
    if grid.shape[0] == 0:
        return central, zero
    ridx = central[real]
    lo, hi = xr.min(0).values, xr.max(0).values
    r2 = radius * radius
    keep_i, keep_s = [], []
    # translations in chunks: bounding-box filter, then the exact distance to the nearest atom of the cell
    step = max(1, int(4.0e6 // max(1, xr.shape[0])))
    for s0 in range(0, grid.shape[0], step):
        g = grid[s0:s0 + step]
        t = g.to(torch.double) @ a                               # [T, 3]
        p = xr.unsqueeze(0) + t.unsqueeze(1)                     # [T, R, 3]
        d = (p - p.clamp(min=lo, max=hi))
        near = (d * d).sum(-1) <= r2                             # [T, R]
        ti, ai = near.nonzero(as_tuple=True)
        if ti.numel() == 0:
            continue
        q = p[ti, ai]                                            # [C, 3]
        ok = torch.zeros(q.shape[0], dtype=torch.bool, device=dev)
        cstep = max(1, int(4.0e6 // max(1, xr.shape[0])))           # [cstep, R, 3] doubles: <= 96 MB
        for c0 in range(0, q.shape[0], cstep):
            qq = q[c0:c0 + cstep]
            d2 = ((qq.unsqueeze(1) - xr.unsqueeze(0)) ** 2).sum(-1)   # exact differences (no |x|^2 expansion)
            ok[c0:c0 + cstep] = d2.min(-1).values <= r2
        ti, ai = ti[ok], ai[ok]
        keep_i.append(ridx[ai])
        keep_s.append(g[ti])
    if not keep_i:
        return central, zero
    return torch.cat([central] + keep_i), torch.cat([zero] + keep_s)


def _gather(t, index, n):
    if t is None:
        return None
    if t.dim() == 1 and t.shape[0] == n:
        return t[index]
    if t.dim() == 2 and t.shape == (n, n):
        return t[index][:, index]
    return t                                                     # element tables pass through


def dftd3_periodic(numbers: Tensor, positions: Tensor, cell: Tensor, param: dict, *, ref=None, rcov=None,
                   rvdw=None, r4r2=None, cutoff=None, pbc=(True, True, True), images=None,
                   _energy_fn=None, **kwargs) -> Tensor:
    """Per-atom DFT-D3 energies `[N]` of the atoms of one unit cell in the periodic lattice spanned by the rows of
    `cell`, with the reference's finite-structure semantics applied to the lattice (module docstring).
    Differentiable with respect to `positions`, `cell` and whatever `dftd3()` differentiates.

    `images` = a `(index, shifts)` pair from `lattice_images` to reuse over the steps of an optimisation /
    MD run while no atom moves by more than the slack that was put on its radius; by default it is rebuilt from
    the current positions with the exact radius.  Keyword arguments of `dftd3()` pass through (`rcov`, `r4r2`
    per atom `[N]` and a dense `rvdw` `[N, N]` are expanded to the images)."""
    from .disp import _atm_requested, _scalar, dftd3

    if numbers.dim() != 1:
        raise ValueError("dftd3_periodic takes one structure: numbers [N], positions [N, 3], cell [3, 3]")
    if numbers.shape != positions.shape[:-1]:
        raise ValueError(
            f"Shape of positions ({tuple(positions.shape[:-1])}) is not consistent with atomic numbers ({tuple(numbers.shape)})."
        )
    n = numbers.shape[0]
    cut = defaults.D3_DISP_CUTOFF if cutoff is None else _scalar(cutoff)
    if images is None:
        images = lattice_images(numbers, positions, cell, image_radius(cut, _atm_requested(param)), pbc)
    index, shifts = images
    index = index.to(positions.device)
    shifts = shifts.to(positions.device)
    if index.shape[0] < n or not bool((index[:n] == torch.arange(n, device=index.device)).all()) \
            or bool((shifts[:n] != 0).any()):
        raise ValueError("images must list the atoms of the cell first (lattice_images)")
    cellp = cell.to(device=positions.device, dtype=positions.dtype)
    ext_positions = positions[index] + shifts.to(positions.dtype) @ cellp
    ext_numbers = numbers.to(positions.device)[index]
    kw = dict(kwargs, ref=ref, cutoff=cutoff, rcov=_gather(rcov, index, n), r4r2=_gather(r4r2, index, n),
              rvdw=_gather(rvdw, index, n))
    if _energy_fn is not None:                                   # test seam (CPU oracle as the evaluator)
        energy = _energy_fn(ext_numbers, ext_positions, param, **kw)
    else:
        energy = dftd3(ext_numbers, ext_positions, param, **kw)
    return energy[:n]
