"""
Host side of the tiled large-structure path (`d3b200_system_*`).

`SystemPlan` does what the kernels expect of their caller (see
`include/d3b200.h`): drop ghost atoms, order the real atoms along a Morton
curve so that 32-atom sub-tiles are spatially compact, sort by element inside
every 128-atom tile (the pair kernel re-contracts the 7x7 reference block only
when the partner element changes), pad to a multiple of 128 with far-away
`Z = 0` atoms, pack the AoS arrays and the sub-tile bounding boxes.

`SystemRunner` owns the per-atom work arrays and issues the sweeps; row ranges
make it the unit of work of the i-row sharded multi-GPU mode
(`tad_dftd3_b200.distributed`).
"""
from __future__ import annotations

import ctypes as C
import os

import torch
from torch import Tensor

from . import _lib

TILE = 128
SUB = 32
_FAR = 1.0e8


def _part1by2(x: Tensor) -> Tensor:
    x = x & 0x3FF
    x = (x | (x << 16)) & 0x30000FF
    x = (x | (x << 8)) & 0x300F00F
    x = (x | (x << 4)) & 0x30C30C3
    x = (x | (x << 2)) & 0x9249249
    return x


def morton_order(pos: Tensor) -> Tensor:
    """argsort of the atoms along a 30-bit Morton curve (cells of >= 4 Bohr)."""
    lo = pos.min(0).values
    extent = float((pos.max(0).values - lo).max())
    cell = max(extent / 1023.0, 4.0)
    q = ((pos - lo) / cell).to(torch.int64).clamp_(0, 1023)
    code = _part1by2(q[:, 0]) | (_part1by2(q[:, 1]) << 1) | (_part1by2(q[:, 2]) << 2)
    return torch.argsort(code, stable=True)


_atm_work_cache: dict = {}  # (device, stream) -> scratch of the triples-once ATM kernel
_static_cache: dict = {}   # key -> _Static
ORDER_MAX_AGE = 20         # calls after which the Morton / element order is rebuilt


class _Static:
    """Everything of a plan that does not depend on the positions: which atoms are
    real, their sorted order, sorted Z / kcn*rcov / sqrt(r4r2), padding, element
    list.  The ORDER depends on the positions only through efficiency (the boxes
    are recomputed from the current positions on every call), so it is reused for
    `ORDER_MAX_AGE` calls while the same `numbers` tensor keeps coming back
    (MD / optimisation loops)."""

    def __init__(self, numbers, positions, rcov, r4r2, kcn, rcov_table, r4r2_table, pad_tiles=1):
        dev, dt = positions.device, positions.dtype
        idx = torch.nonzero(numbers != 0).reshape(-1)
        n = int(idx.numel())
        self.n = n
        unit = TILE * max(1, int(pad_tiles))          # equal tile counts per rank: pad to world * 128
        self.npad = npad = max(unit, -(-n // unit) * unit)
        z = numbers[idx]
        if n:
            o1 = morton_order(positions.detach()[idx])
            # With <= 8 distinct elements the sweeps read pre-contracted T vectors and do not
            # care about the element order, so the tiles stay purely spatial (tighter 32-atom
            # boxes, better culling).  Otherwise sort by element inside each 128-atom tile: the
            # general pair kernel re-contracts the 7x7 block only when the partner element changes.
            few = int(torch.unique(z).numel()) <= 8 and os.environ.get("TAD_DFTD3_B200_NO_TVEC") != "1"
            if few and os.environ.get("TAD_DFTD3_B200_ELEMENT_SORT") != "1":
                order = o1
            else:
                tile = torch.arange(n, device=dev) // TILE
                o2 = torch.argsort(tile * 128 + z[o1], stable=True)
                order = o1[o2]
        else:
            order = idx
        self.perm = idx[order]                      # sorted slot -> original atom index
        zs = z[order]
        self.z = torch.zeros(npad, dtype=torch.int32, device=dev)
        self.z[:n] = zs.to(torch.int32)
        rc = rcov.detach().to(dt)
        q = r4r2.detach().to(dt)
        rc = rc[zs] if rcov_table else rc[self.perm]
        q = q[zs] if r4r2_table else q[self.perm]
        template = torch.zeros(npad, 4, dtype=dt, device=dev)
        template[n:, :3] = (_FAR + torch.arange(npad - n, device=dev, dtype=dt) * 100.0).unsqueeze(-1)
        template[:n, 3] = kcn * rc
        self.template = template
        self.sqrtq = torch.ones(npad, dtype=dt, device=dev)
        self.sqrtq[:n] = torch.sqrt(q)
        self.aux1 = torch.stack([self.sqrtq, torch.ones_like(self.sqrtq)], 1).contiguous()
        self.real = (self.z != 0).reshape(npad // SUB, SUB, 1)
        self.empty = ~self.real.any(1)
        zlist = torch.unique(self.z[self.z > 0])
        lut = torch.full((128,), -1, dtype=torch.int8, device=dev)
        lut[zlist.long()] = torch.arange(zlist.numel(), dtype=torch.int8, device=dev)
        self.elements = (zlist.to(torch.int32).contiguous(), lut[self.z.long()].contiguous())
        self.age = 0


def _static_for(numbers, positions, rcov, r4r2, kcn, rcov_table, r4r2_table, pad_tiles=1) -> _Static:
    import weakref

    def tkey(t, table):
        return ("table" if table else "atom", id(t), t._version)

    key = (id(numbers), numbers._version, positions.dtype, str(positions.device), float(kcn),
           tkey(rcov, rcov_table), tkey(r4r2, r4r2_table), int(pad_tiles))
    hit = _static_cache.get(key)
    if hit is not None and hit[0]() is numbers and hit[1]() is rcov and hit[2]() is r4r2 \
            and hit[3].age < ORDER_MAX_AGE:
        hit[3].age += 1
        return hit[3]
    st = _Static(numbers, positions, rcov, r4r2, kcn, rcov_table, r4r2_table, pad_tiles)
    if len(_static_cache) > 32:
        _static_cache.clear()
    try:
        _static_cache[key] = (weakref.ref(numbers), weakref.ref(rcov), weakref.ref(r4r2), st)
    except TypeError:  # pragma: no cover
        pass
    return st


class SystemPlan:
    """Sorted / padded / packed view of one structure.

    `rcov` / `r4r2` are per-atom arrays `[N]` (the reference's keyword arrays) or,
    with `*_table=True`, per-element tables indexed by Z."""

    def __init__(self, numbers: Tensor, positions: Tensor, rcov: Tensor, r4r2: Tensor, kcn: float,
                 rcov_table: bool = False, r4r2_table: bool = False, pad_tiles: int = 1):
        dev, dt = positions.device, positions.dtype
        self.natoms_in = numbers.shape[0]
        st = _static_for(numbers, positions, rcov, r4r2, kcn, rcov_table, r4r2_table, pad_tiles)
        self.static = st
        self.n, self.npad, self.perm, self.z, self.sqrtq = st.n, st.npad, st.perm, st.z, st.sqrtq
        n, npad = st.n, st.npad
        self.dtype, self.device = dt, dev
        if dev.type == "cuda" and os.environ.get("TAD_DFTD3_B200_TORCH_PACK") != "1":
            # one launch: gather + AoS packing + sub-tile boxes (d3b200_system_pack_*)
            self.atoms = torch.empty(npad, 4, dtype=dt, device=dev)
            self.box = torch.empty(npad // SUB, 6, dtype=dt, device=dev)
            self.repack(positions)
            return
        # host tensors (the CPU tests of the sharding logic): the same arrays with tensor operations
        atoms = st.template.clone()
        atoms[:n, :3] = positions.detach()[st.perm]
        self.atoms = atoms
        # sub-tile boxes over the real atoms (all-padding sub-tiles sit far away)
        xyz = atoms[:, :3].reshape(npad // SUB, SUB, 3)
        big = torch.finfo(dt).max
        lo = torch.where(st.real, xyz, xyz.new_tensor(big)).amin(1)
        hi = torch.where(st.real, xyz, xyz.new_tensor(-big)).amax(1)
        lo = torch.where(st.empty, xyz[:, 0, :], lo)
        hi = torch.where(st.empty, xyz[:, 0, :], hi)
        self.box = torch.cat([lo, hi], 1).contiguous()

    def repack(self, positions: Tensor) -> None:
        """Refresh `atoms` / `box` IN PLACE from new positions of the same structure (same
        sorted order; the boxes are recomputed, so the order only matters for efficiency).
        One launch of `d3b200_system_pack_*`; the arrays keep their addresses, which is what
        lets a captured CUDA graph of the sweeps be replayed for every step."""
        st, dev, dt = self.static, self.device, self.dtype
        x = positions.detach()
        x = x if x.is_contiguous() else x.contiguous()
        fn = getattr(_lib.lib(), "d3b200_system_pack_" + ("f64" if dt == torch.float64 else "f32"))
        _lib.launch_count += 1
        with torch.cuda.device(dev):
            _lib.check(fn(self.npad, self.n, C.c_void_p(st.perm.data_ptr()), C.c_void_p(x.data_ptr()),
                          C.c_void_p(st.template.data_ptr()), C.c_void_p(st.z.data_ptr()),
                          C.c_void_p(self.atoms.data_ptr()), C.c_void_p(self.box.data_ptr()),
                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
        self._keep = x

    def elements(self):
        """(zlist int32 [E], eidx int8 [npad]) -- distinct elements and each atom's index (-1 = padding)."""
        return self.static.elements

    def aux(self, g: Tensor | None) -> Tensor:
        """[npad, 2] = sqrt(r4r2), cotangent g (ones when None)."""
        if g is None:
            return self.static.aux1               # read-only use
        a = self.static.aux1.clone()
        a[: self.n, 1] = g.detach().to(self.dtype)[self.perm]
        return a

    def scatter(self, sorted_vals: Tensor) -> Tensor:
        """Sorted per-atom values -> original atom order (zeros on ghosts)."""
        out = torch.zeros(self.natoms_in, *sorted_vals.shape[1:], dtype=sorted_vals.dtype, device=self.device)
        out[self.perm] = sorted_vals[: self.n]
        return out


class SystemRunner:
    """Work arrays + launches of the four sweeps for one plan."""

    # CTAs wanted per launch: enough to give every SM several waves of 128-thread CTAs
    TARGET_CTAS = 148 * 12

    def __init__(self, plan: SystemPlan, refcn: Tensor, refc6: Tensor, par: _lib.Params, g: Tensor | None = None,
                 world: int = 1, rvdw: Tensor | None = None):
        self.plan = plan
        self.rvdw = rvdw                      # [104, 104] element-pair table (ATM only)
        dt, dev, npad = plan.dtype, plan.device, plan.npad
        self.par = par
        self.sfx = "f64" if dt == torch.float64 else "f32"
        self.auxt = plan.aux(g)
        buf = torch.zeros(22 * npad, dtype=dt, device=dev)      # one allocation, one fill
        self.buf = buf
        self.cn, self.decn, self.energy = buf[:npad], buf[npad:2 * npad], buf[2 * npad:3 * npad]
        self.grad = buf[3 * npad:6 * npad].view(npad, 3)
        self.energy_grad = buf[2 * npad:6 * npad]               # contiguous [E | dL/dx]: one message on several GPUs
        self.w, self.dw = buf[6 * npad:14 * npad].view(npad, 8), buf[14 * npad:22 * npad].view(npad, 8)
        self.refcn, self.refc6 = refcn, refc6
        s = _lib.System()
        s.natoms = npad
        s.atoms, s.aux, s.z, s.box = (t.data_ptr() for t in (plan.atoms, self.auxt, plan.z, plan.box))
        s.refcn, s.refc6 = refcn.data_ptr(), refc6.data_ptr()
        s.cn, s.w, s.dw, s.decn = (t.data_ptr() for t in (self.cn, self.w, self.dw, self.decn))
        s.energy, s.grad = self.energy.data_ptr(), self.grad.data_ptr()
        ntiles = npad // TILE
        rows_tiles = max(1, ntiles // max(world, 1))
        # with four or more ranks a rank's ~50-100 row tiles need a finer j-split to fill 148 SMs:
        # 50k atoms on 8 GPUs 0.86 ms at 12 CTAs per SM, 0.79 ms at 24, 1.01 ms at 8 (profiles/r2_summary.md)
        target = self.TARGET_CTAS * (2 if world >= 4 else 1)
        target = int(os.environ.get("TAD_DFTD3_B200_TARGET_CTAS", target))             # experiment knob
        self.jsplit = int(max(1, min(ntiles, -(-target // rows_tiles))))
        self.scratch = torch.empty(5 * self.jsplit * npad, dtype=dt, device=dev)
        s.jsplit, s.scratch = self.jsplit, self.scratch.data_ptr()
        # pre-contracted reference vectors (structures with <= 8 distinct elements)
        s.nelem = 0
        self.use_tvec = False
        if os.environ.get("TAD_DFTD3_B200_NO_TVEC") != "1":
            zlist, eidx = plan.elements()
            ne = int(zlist.numel())
            if 0 < ne <= 8:
                self._zlist, self._eidx = zlist, eidx
                self.tvec = torch.empty(npad * ne * 16 + 64, dtype=dt, device=dev)
                s.nelem, s.zlist, s.eidx, s.tvec = ne, zlist.data_ptr(), eidx.data_ptr(), self.tvec.data_ptr()
                self.use_tvec = True
        # Pairs-once sweeps need the T vectors and pay a fixed cost per CTA (staging
        # tiles, atomics), so they win only while a CTA still walks many j-tiles:
        # measured on the 50k-atom cluster, 5.75 vs 6.47 ms at jsplit 5 (one GPU) but
        # 1.31 vs 1.14 ms per rank at jsplit 37 (eight GPUs).
        # TAD_DFTD3_B200_FULLROW=1 / =0 forces one or the other.
        # Round 2, with the plan, the work arrays and the launch sequence cached in a CUDA graph (engine.py) the
        # fixed costs are gone and the pairs-once sweeps win at every GPU count measured: 3.98 / 2.24 / 1.31 /
        # 0.89 ms at 1 / 2 / 4 / 8 GPUs against 5.48 / 3.08 / - / 1.09 ms full-row (profiles/r2_summary.md).
        force = os.environ.get("TAD_DFTD3_B200_FULLROW")
        self.symmetric = self.use_tvec and force != "1"
        # experiment knob: pairs-once sweeps with a capped j-split (fewer, longer CTAs)
        cap = os.environ.get("TAD_DFTD3_B200_SYM_JSPLIT")
        if cap and self.use_tvec:
            self.symmetric = True
            self.jsplit = int(max(1, min(self.jsplit, int(cap))))
            s.jsplit = self.jsplit
        s.symmetric, s.tile_stride = int(self.symmetric), 1
        self.sys = s

    def partition(self, rank: int, world: int):
        """Rows this rank sweeps.  Full-row kernels: a contiguous tile-aligned range.
        Pairs-once kernels: every `world`-th row tile starting at tile `rank`
        (the triangle J >= I is balanced that way)."""
        if self.symmetric:
            self.sys.tile_stride = max(1, world)
            return (min(rank * TILE, self.plan.npad), self.plan.npad)
        self.sys.tile_stride = 1
        return row_partition(self.plan.npad, world)[rank]

    def _call(self, name, rows, *extra):
        self.sys.row_begin, self.sys.row_end = rows if rows is not None else (0, self.plan.npad)
        fn = getattr(_lib.lib(), f"d3b200_system_{name}_{self.sfx}")
        _lib.launch_count += 1
        with torch.cuda.device(self.plan.device):
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(fn(C.byref(self.sys), C.byref(self.par), *extra, stream))

    def sweep_cn(self, rows=None):
        self._call("cn", rows)

    def weights(self):
        self._call("weights", None)
        if self.use_tvec:
            self._call("tvec", None)

    def sweep_pair(self, want_grad: bool, rows=None):
        self._call("pair", rows, int(want_grad))

    def sweep_cnbwd(self, rows=None):
        self._call("cnbwd", rows)

    def atm_partition(self, rank: int, world: int):
        """Centres of the three-body term this rank owns: every `world`-th atom of the sorted
        order starting at atom `rank`; needs the triples-once kernel."""
        if world <= 1 or not self.use_tvec or os.environ.get("TAD_DFTD3_B200_ATM2") == "1":
            self._atm_stride = 1
            return row_partition(self.plan.npad, world)[rank]
        # atom-by-atom round robin (negative stride, see d3b200.h): neighbouring atoms of the sorted order
        # cost about the same, so the ranks finish together (sub-tile round robin left 10 % imbalance at 8 GPUs)
        self._atm_stride = -world
        return (min(rank, self.plan.npad), self.plan.npad)

    def sweep_atm(self, want_grad: bool, rows=None):
        """Three-body term (adds to energy / grad / decn); no-op when s9 == 0."""
        if self.par.s9 == 0.0:
            return
        if self.rvdw is None:
            raise ValueError("the ATM term needs the [104,104] rvdw table")
        if self.use_tvec and self.sys.atm_work is None and os.environ.get("TAD_DFTD3_B200_ATM2") != "1":
            # scratch of the triples-once kernel (per-CTA neighbour lists and records)
            nbytes = int(getattr(_lib.lib(), f"d3b200_system_atm_workspace_bytes_{self.sfx}")(self.plan.npad))
            # several hundred MB: one buffer per device and stream, reused by every runner (work on one
            # stream is ordered, so consecutive calls cannot overlap in it)
            dev = self.plan.device
            key = (str(dev), torch.cuda.current_stream(dev).cuda_stream)
            buf = _atm_work_cache.get(key)
            if buf is None or buf.numel() < nbytes:
                buf = _atm_work_cache[key] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._atm_work = buf
            self.sys.atm_work, self.sys.atm_work_bytes = buf.data_ptr(), buf.numel()
        keep = self.sys.tile_stride
        self.sys.tile_stride = getattr(self, "_atm_stride", 1)
        try:
            self._call("atm", rows, C.c_void_p(self.rvdw.data_ptr()), int(want_grad))
        finally:
            self.sys.tile_stride = keep

    def param_gradient(self) -> Tensor:
        """dL/d(s6, s8, a1, a2, s9) with L = sum_i g_i E_i for this structure (single GPU; after
        `weights()`).  Two-body part: one more full-row sweep (`d3b200_system_pgrad_*`) whose
        per-row sums are contracted with -g/2; s9: the three-body energies alone, divided by s9."""
        if not self.use_tvec:
            raise NotImplementedError(
                "damping-parameter gradients of structures beyond the fused-kernel limit need "
                "the T-vector sweeps (at most 8 distinct elements)")
        npad, js = self.plan.npad, self.jsplit
        self.scratch.zero_()
        keep = self.sys.symmetric
        self.sys.symmetric = 0
        try:
            self._call("pgrad", None)
        finally:
            self.sys.symmetric = keep
        rows = self.scratch[: 4 * js * npad].view(4, js, npad).sum(1)           # [4, npad]
        g = self.auxt[:, 1]
        out = torch.zeros(5, dtype=self.plan.dtype, device=self.plan.device)
        out[:4] = rows @ (-0.5 * g)
        if self.par.s9 != 0.0:
            e_atm = torch.zeros(npad, dtype=self.plan.dtype, device=self.plan.device)
            keep_e = self.sys.energy
            self.sys.energy = e_atm.data_ptr()
            try:
                self.sweep_atm(False)
            finally:
                self.sys.energy = keep_e
            out[4] = (e_atm * g).sum() / self.par.s9
        return out

    # single-GPU drivers -----------------------------------------------------
    def run_energy(self) -> Tensor:
        self.sweep_cn()
        self.weights()
        self.sweep_pair(False)
        self.sweep_atm(False)
        return self.plan.scatter(self.energy)

    def run_vjp(self):
        self.sweep_cn()
        self.weights()
        self.sweep_pair(True)
        self.sweep_atm(True)
        self.sweep_cnbwd()
        return self.plan.scatter(self.energy), self.plan.scatter(self.grad)


def row_partition(npad: int, world: int):
    """Tile-aligned contiguous row ranges, one per rank."""
    ntiles = npad // TILE
    bounds = [TILE * ((ntiles * r) // world) for r in range(world + 1)]
    return [(bounds[r], bounds[r + 1]) for r in range(world)]
