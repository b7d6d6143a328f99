"""
Multi-rank logic of the i-row sharded path on CPU: world_size 2 over gloo.

The CUDA sweeps cannot run here, so `run_sharded` (the code that partitions the
rows, orders the sweeps and exchanges CN / dL/dCN / E / gradient) is driven
with a stand-in runner whose four sweeps are dense torch expressions built on
the CPU oracle, restricted to the rank's rows exactly like the kernels.  The
assembled energy and gradient must equal the oracle's full result.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleRunner:
    """`SystemRunner` interface on CPU (rows of the full-row quantities)."""

    def __init__(self, plan, refcn, refc6, par, g=None, rvdw=None):
        import d3_oracle as orc

        self.orc, self.plan, self.par = orc, plan, par
        n = plan.npad
        self.z = plan.z.long()
        self.x = plan.atoms[:, :3].clone()
        self.rcov = plan.atoms[:, 3] / par.kcn
        self.q = plan.sqrtq ** 2
        self.g = plan.aux(g)[:, 1]
        self.refcn, self.refc6, self.rvdw = refcn, refc6, rvdw
        z = lambda *s: torch.zeros(*s, dtype=torch.double)  # noqa: E731
        self.cn, self.decn, self.energy, self.grad = z(n), z(n), z(n), z(n, 3)
        self.w = None

    def sweep_cn(self, rows):
        full = self.orc.cn_d3(self.z, self.x, self.rcov)
        self.cn[rows[0]:rows[1]] = full[rows[0]:rows[1]]

    def weights(self):
        self.w_ready = True  # weights are recomputed from the gathered CN below

    def _pair_terms(self, x, cn):
        w = self.orc.weight_references(self.z, cn, self.refcn)
        c6 = self.orc.atomic_c6(self.z, w, self.refc6)
        p = self.par
        return self.orc.dispersion2(self.z, x, c6, self.q, p.s6, p.s8, p.a1, p.a2, p.cutoff)

    def sweep_pair(self, want_grad, rows):
        x = self.x.clone().requires_grad_(True)
        cn = self.cn.clone().requires_grad_(True)
        e = self._pair_terms(x, cn)
        sl = slice(rows[0], rows[1])
        self.energy[sl] = e.detach()[sl]
        if want_grad:
            gx, gcn = torch.autograd.grad((e * self.g).sum(), (x, cn))
            self.grad[sl] = gx[sl]
            self.decn[sl] = gcn[sl]

    def sweep_atm(self, want_grad, rows):
        """Three-body term of the atoms this rank owns.  Like the CUDA kernel it ADDS
        to arrays that other ranks own as well (gradient and dL/dCN of every atom of
        a triple), which is what the all-reduces of `run_sharded` have to get right."""
        p = self.par
        if p.s9 == 0.0:
            return
        real = torch.nonzero(self.z).reshape(-1)
        z = self.z[real]
        x = self.x[real].clone().requires_grad_(True)
        cn = self.cn[real].clone().requires_grad_(True)
        w = self.orc.weight_references(z, cn, self.refcn)
        c6 = self.orc.atomic_c6(z, w, self.refc6)
        rv = self.rvdw[z.unsqueeze(-1), z.unsqueeze(-2)]
        e = self.orc.dispersion_atm(z, x, c6, rv, p.cutoff, p.s9, p.rs9, p.alp)
        own = (real >= rows[0]) & (real < rows[1])
        self.energy[real[own]] += e.detach()[own]
        if want_grad and bool(own.any()):
            gx, gcn = torch.autograd.grad((e * self.g[real])[own].sum(), (x, cn))
            self.grad[real] += gx
            self.decn[real] += gcn

    def sweep_cnbwd(self, rows):
        x = self.x.clone().requires_grad_(True)
        cn = self.orc.cn_d3(self.z, x, self.rcov)
        (gx,) = torch.autograd.grad((cn * self.decn).sum(), x)
        self.grad[rows[0]:rows[1]] += gx[rows[0]:rows[1]]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, atm=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tad_dftd3_b200 import _lib, synthetic as syn
        from tad_dftd3_b200.data import COV_D3, R4R2, REFCN
        from tad_dftd3_b200.distributed import run_sharded
        from tad_dftd3_b200.system import SystemPlan, row_partition

        dt = torch.double
        # 300 atoms -> 3 tiles of 128; with the (dense N^3) three-body term 150 atoms -> 2 tiles
        numbers, positions = syn.water_cluster(50 if atm else 100, seed=4)
        numbers = numbers.clone()
        numbers[[5, 17]] = 0                                       # ghost atoms
        par = _lib.Params()
        par.s6, par.s8, par.a1, par.a2 = 1.0, 0.78981345, 0.49484001, 5.73083694
        par.cutoff, par.cn_cutoff, par.kcn = (12.0 if atm else 30.0), 25.0, 16.0
        par.s9, par.rs9, par.alp = (1.0 if atm else 0.0), 1.3333333730697632, 14.0
        g = torch.linspace(0.5, 1.5, numbers.shape[0], dtype=dt)
        plan = SystemPlan(numbers, positions, COV_D3(dt)[numbers], R4R2(dt)[numbers], par.kcn)
        runner = OracleRunner(plan, REFCN(dt), syn.synthetic_c6_table(dt), par, g, syn.synthetic_rvdw_table(dt))
        rows = run_sharded(runner, True)
        assert rows == row_partition(plan.npad, world)[rank]
        e, gr = plan.scatter(runner.energy), plan.scatter(runner.grad)
        # every rank holds the full result
        gathered = [torch.zeros_like(e) for _ in range(world)]
        dist.all_gather(gathered, e)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        if rank == 0:
            np.savez(out, energy=e.numpy(), grad=gr.numpy(), rows=np.asarray(rows), npad=plan.npad)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_two_ranks_matches_oracle(tmp_path):
    import d3_oracle as orc
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    assert res["npad"] == 384 and tuple(res["rows"]) == (0, 128)

    dt = torch.double
    numbers, positions = syn.water_cluster(100, seed=4)
    numbers = numbers.clone()
    numbers[[5, 17]] = 0
    g = torch.linspace(0.5, 1.5, numbers.shape[0], dtype=dt)
    pos = positions.clone().requires_grad_(True)
    param = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694}
    e = orc.dftd3(numbers, pos, param, refcn=REFCN(dt), refc6=syn.synthetic_c6_table(dt),
                  rcov=COV_D3(dt)[numbers], r4r2=R4R2(dt)[numbers], cutoff=30.0)
    (gr,) = torch.autograd.grad((e * g).sum(), pos)
    np.testing.assert_allclose(res["energy"], e.detach().numpy(), rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(res["grad"], gr.numpy(), rtol=1e-10, atol=1e-13)
    assert (res["energy"][[5, 17]] == 0).all()


@pytest.mark.timeout(600)
def test_sharded_two_ranks_with_three_body_term(tmp_path):
    """s9 != 0: each rank adds the three-body term of its own atoms into arrays the
    other rank owns too; the exchanges of `run_sharded` must sum them (cutoff 12 Bohr
    cuts triples partly, so the reference's ordered cutoff rule is exercised)."""
    import d3_oracle as orc
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN

    out = str(tmp_path / "res_atm.npz")
    mp.spawn(_worker, args=(2, _free_port(), out, True), nprocs=2, join=True)
    res = np.load(out)
    assert res["npad"] == 256 and tuple(res["rows"]) == (0, 128)

    dt = torch.double
    numbers, positions = syn.water_cluster(50, seed=4)
    numbers = numbers.clone()
    numbers[[5, 17]] = 0
    g = torch.linspace(0.5, 1.5, numbers.shape[0], dtype=dt)
    pos = positions.clone().requires_grad_(True)
    param = {"s6": 1.0, "s8": 0.78981345, "a1": 0.49484001, "a2": 5.73083694, "s9": 1.0}
    rvdw = syn.synthetic_rvdw_table(dt)
    e = orc.dftd3(numbers, pos, param, refcn=REFCN(dt), refc6=syn.synthetic_c6_table(dt),
                  rcov=COV_D3(dt)[numbers], r4r2=R4R2(dt)[numbers],
                  rvdw=rvdw[numbers.unsqueeze(-1), numbers.unsqueeze(-2)], cutoff=12.0)
    e2 = orc.dftd3(numbers, pos, {k: v for k, v in param.items() if k != "s9"}, refcn=REFCN(dt),
                   refc6=syn.synthetic_c6_table(dt), rcov=COV_D3(dt)[numbers], r4r2=R4R2(dt)[numbers], cutoff=12.0)
    assert float((e - e2).abs().max()) > 1e-9          # the three-body term is really there
    (gr,) = torch.autograd.grad((e * g).sum(), pos)
    np.testing.assert_allclose(res["energy"], e.detach().numpy(), rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(res["grad"], gr.numpy(), rtol=1e-10, atol=1e-13)
    assert (res["energy"][[5, 17]] == 0).all()


def test_row_partition_is_tile_aligned_and_complete():
    from tad_dftd3_b200.system import row_partition

    for npad in (128, 384, 50048):
        for world in (1, 2, 3, 4, 8):
            parts = row_partition(npad, world)
            assert parts[0][0] == 0 and parts[-1][1] == npad
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c
            assert all(a % 128 == 0 and b % 128 == 0 for a, b in parts)


def test_plan_sorting_and_padding(monkeypatch):
    from tad_dftd3_b200 import synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2
    from tad_dftd3_b200.system import SystemPlan

    dt = torch.double
    numbers, positions = syn.water_cluster(60, seed=1)
    numbers = numbers.clone()
    numbers[3] = 0
    plan = SystemPlan(numbers, positions, COV_D3(dt)[numbers], R4R2(dt)[numbers], 16.0)
    assert plan.n == 179 and plan.npad == 256
    # permutation covers exactly the real atoms
    assert sorted(plan.perm.tolist()) == [i for i in range(180) if i != 3]
    # few elements (T-vector sweeps): purely spatial order; with the general pair kernel
    # the elements are sorted inside each 128-atom tile
    monkeypatch.setenv("TAD_DFTD3_B200_ELEMENT_SORT", "1")
    numbers2 = numbers.clone()
    plan2 = SystemPlan(numbers2, positions, COV_D3(dt)[numbers2], R4R2(dt)[numbers2], 16.0)
    z = plan2.z[: plan2.n].reshape(-1).tolist()
    assert z[:128] == sorted(z[:128]) and z[128:] == sorted(z[128:])
    assert sorted(plan2.perm.tolist()) == sorted(plan.perm.tolist())
    # boxes bound their real atoms; padding sits far away
    xyz = plan.atoms[:, :3].reshape(-1, 32, 3)
    for s in range(plan.npad // 32):
        real = plan.z.reshape(-1, 32)[s] != 0
        if real.any():
            assert (xyz[s][real] >= plan.box[s, :3] - 1e-12).all()
            assert (xyz[s][real] <= plan.box[s, 3:] + 1e-12).all()
    assert (plan.atoms[plan.n:, 0] > 1e7).all()
    # scatter puts sorted values back and zeros on ghosts
    vals = torch.arange(plan.npad, dtype=dt)
    back = plan.scatter(vals)
    assert back[3] == 0 and torch.equal(back[plan.perm], vals[: plan.n])


def test_three_body_centre_ownership_covers_every_atom_once():
    """`SystemRunner.atm_partition` + the kernel's index arithmetic (csrc/d3_atm3.cuh): with a negative stride a
    launch owns the atoms row_begin + k*|stride| below row_end, with a positive one the 32-atom sub-tiles
    row_begin/32 + k*stride; over the ranks every centre exactly once."""
    from tad_dftd3_b200 import _lib, synthetic as syn
    from tad_dftd3_b200.data import COV_D3, R4R2, REFCN
    from tad_dftd3_b200.system import SUB, SystemPlan, SystemRunner

    dt = torch.double
    numbers, positions = syn.water_cluster(150, seed=2)               # 450 atoms -> 512 padded
    par = _lib.Params()
    par.s6, par.s8, par.a1, par.a2, par.s9 = 1.0, 0.79, 0.49, 5.73, 1.0
    par.cutoff, par.cn_cutoff, par.kcn = 20.0, 25.0, 16.0
    plan = SystemPlan(numbers, positions, COV_D3(dt)[numbers], R4R2(dt)[numbers], par.kcn)
    for world in (1, 2, 3, 8):
        owned = []
        for rank in range(world):
            run = SystemRunner(plan, REFCN(dt), syn.synthetic_c6_table(dt), par, None, world, syn.synthetic_rvdw_table(dt))
            begin, end = run.atm_partition(rank, world)
            stride = run._atm_stride
            if stride < 0:                                              # atom by atom (the kernel's arithmetic)
                nown = (end - begin + (-stride) - 1) // (-stride)
                owned += [begin + k * (-stride) for k in range(nown)]
            else:                                                       # sub-tiles
                assert begin % SUB == 0 and end % SUB == 0
                nsub_own = ((end - begin) // SUB + stride - 1) // stride
                owned += [s * SUB + a for s in (begin // SUB + k * stride for k in range(nsub_own)) for a in range(SUB)]
        owned = [a for a in owned if a < plan.npad]
        assert sorted(owned) == list(range(plan.npad)), (world, len(owned))
