"""Host logic of the slab-parallel path on CPU: the partition, and a world_size-2 gloo run in which
each rank advances its slab with the ORACLE RHS (test infrastructure) using the exchange schedule of
fluxreconstruction.jl_b200/partition.py; the stitched result must equal the single-domain oracle."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_partition_covers_everything(FR):
    P = FR.partition
    for n, w in ((2048, 8), (2050, 8), (17, 4), (5, 5), (96, 2)):
        slabs = [P.slab(n, w, r) for r in range(w)]
        assert slabs[0].start == 1 and slabs[-1].stop == n
        assert all(a.stop + 1 == b.start for a, b in zip(slabs, slabs[1:]))
        assert max(s.count for s in slabs) - min(s.count for s in slabs) <= 1
        assert all(s.lo == (s.rank - 1) % w and s.hi == (s.rank + 1) % w for s in slabs)
    with pytest.raises(ValueError):
        P.slab(3, 4, 0)


def test_exchange_plans(FR):
    P = FR.partition
    s0, s1, s2 = (P.slab(30, 3, r) for r in range(3))
    assert P.stage_exchange_plan(s0) == [(1, 10, 11)]
    assert P.stage_exchange_plan(s1) == [(0, 1, 0), (2, 10, 11)]
    assert P.stage_exchange_plan(s2) == [(1, 1, 0)]
    assert P.step_ghost_plan(s0, True) == [(2, 1, 0)]
    assert P.step_ghost_plan(s1, True) == []
    assert P.step_ghost_plan(s2, True) == [(0, 10, 11)]
    assert P.step_ghost_plan(s0, False) == []


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist

    import fr_oracle as o
    import frb200 as FR

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = 5.0 / 3.0
    nx, nyg, nsteps, dt = 8, 10, 3, 1e-3
    psg = o.FRPSpace2D(0.0, 1.0, nx, 0.0, 1.0, nyg, 2, 1, 1)
    rng = np.random.default_rng(11)
    ug = np.asfortranarray(o.ic_wave2d(psg, g, "x") * (1 + 0.01 * rng.standard_normal((nx + 2, nyg + 2, 3, 3, 4))))
    sl = FR.partition.slab(nyg, world, rank)
    u = np.asfortranarray(ug[:, sl.start - 1: sl.stop + 2].copy())
    psl = o.FRPSpace2D(0.0, 1.0, nx, 0.0, sl.count / nyg, sl.count, 2, 1, 1)

    def exchange(a, plan, flip_var=None):
        reqs, bufs = [], []
        for peer, send_row, recv_row in plan:
            s = torch.from_numpy(np.ascontiguousarray(a[:, send_row]))
            r = torch.empty_like(s)
            reqs += [dist.isend(s, peer), dist.irecv(r, peer)]
            bufs.append((recv_row, r))
        for rq in reqs:
            rq.wait()
        for recv_row, r in bufs:
            a[:, recv_row] = r.numpy()
            if flip_var is not None:
                a[:, recv_row, :, :, flip_var] *= -1

    def rhs(a):
        return o.rhs_euler2d(a, psl, g)

    exchange(u, FR.partition.stage_exchange_plan(sl))
    for _ in range(nsteps):
        # per-step ghost fill (euler2d_wave.jl:127-132): x locally, the y seam between first/last rank
        u[0] = u[nx]
        u[nx + 1] = u[1]
        exchange(u, FR.partition.step_ghost_plan(sl, True), flip_var=2)
        frozen = [u[:, 0].copy() if sl.is_first else None, u[:, -1].copy() if sl.is_last else None]

        def stage(a):
            exchange(a, FR.partition.stage_exchange_plan(sl))
            if frozen[0] is not None:
                a[:, 0] = frozen[0]
            if frozen[1] is not None:
                a[:, -1] = frozen[1]
            return a

        u1 = stage(u + (0.5 * dt) * rhs(u))  # Midpoint; du = 0 in the x ghosts keeps them frozen
        u = np.asfortranarray(u + dt * rhs(u1))
        u = stage(u)
    parts = [None] * world
    dist.all_gather_object(parts, (sl.start, sl.count, u[:, 1:-1].copy()))
    if rank == 0:
        ref = o.integrate(ug, dt, nsteps, lambda w: o.rhs_euler2d(w, psg, g), "midpoint",
                          lambda w: o.ghost_fill_euler2d(w, "wave_x"))
        got = np.zeros_like(ref)
        for st, cnt, arr in parts:
            got[:, st: st + cnt] = arr
        q.put(float(np.abs(got[1:-1, 1:-1] - ref[1:-1, 1:-1]).max()))
    dist.destroy_process_group()


def test_two_rank_gloo_slab_exchange_matches_single_domain():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-13


def _worker_ns(rank, world, port, q):
    """cfg5 in column slabs (the slowest index of u[4, ns, nr, ny+2, nx+2]): walls only where the slab touches the
    cavity's x walls, the neighbour's boundary column of the current stage everywhere else."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist

    import fr_oracle as o
    import frb200 as FR

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, nxg, ny, deg, nsteps = 5.0 / 3.0, 7, 5, 2, 3
    psg = o.FRPSpace2D(0.0, 1.0, nxg, 0.0, 1.0, ny, deg, 1, 1)
    rng = np.random.default_rng(12)
    ug = o.ic_cavity(psg, g)
    ug = np.asfortranarray(ug * (1 + 0.02 * rng.standard_normal(ug.shape)))
    mu = FR.ref_vhs_vis(1e-3, 1.0, 0.5)
    dt = 0.1 * min(psg.dx, psg.dy) / 3.0
    sl = FR.partition.slab(nxg, world, rank)
    u = np.asfortranarray(ug[..., sl.start - 1: sl.stop + 2].copy())
    psl = o.FRPSpace2D(0.0, sl.count / nxg, sl.count, 0.0, 1.0, ny, deg, 1, 1)
    walls = o.ns_boundary

    def slab_boundary(a, gamma, lam0=1.0, lid=0.15):  # boundary! without the walls a neighbour stands in for
        lo, hi = a[..., 0].copy(), a[..., -1].copy()
        walls(a, gamma, lam0, lid)
        if not sl.is_first:
            a[..., 0] = lo
        if not sl.is_last:
            a[..., -1] = hi
        return a

    o.ns_boundary = slab_boundary

    def exchange(a):
        reqs, bufs = [], []
        for peer, send_col, recv_col in FR.partition.stage_exchange_plan(sl):
            s = torch.from_numpy(np.ascontiguousarray(a[..., send_col]))
            r = torch.empty_like(s)
            reqs += [dist.isend(s, peer), dist.irecv(r, peer)]
            bufs.append((recv_col, r))
        for rq in reqs:
            rq.wait()
        for recv_col, r in bufs:
            a[..., recv_col] = r.numpy()

    exchange(u)
    for _ in range(nsteps):  # Euler forward, ns_cavity.jl:380
        u = np.asfortranarray(u + dt * o.rhs_ns2d(u, psl, 1.0, g, mu, 0.81, dt))
        exchange(u)
    parts = [None] * world
    dist.all_gather_object(parts, (sl.start, sl.count, u[..., 1:-1].copy()))
    if rank == 0:
        o.ns_boundary = walls
        ref = ug.copy(order="F")
        for _ in range(nsteps):
            ref = np.asfortranarray(ref + dt * o.rhs_ns2d(ref, psg, 1.0, g, mu, 0.81, dt))
        got = np.zeros_like(ref)
        for st, cnt, arr in parts:
            got[..., st: st + cnt] = arr
        q.put(float(np.abs(got[..., 1:-1, 1:-1] - ref[..., 1:-1, 1:-1]).max()))
    dist.destroy_process_group()


def test_two_rank_gloo_cavity_column_slabs_match_single_domain():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker_ns, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) < 1e-13
