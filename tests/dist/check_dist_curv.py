"""Test worker (imports the oracle: test infrastructure).  Slab-parallel parity check of the curvilinear path
(run under torchrun, one rank per GPU): N row slabs of a sheared mesh, the periodic ghost fill of
dev/parallelogram.jl:201-205 across the global seam, must equal the single-domain oracle on the same global mesh.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29541 tests/dist/check_dist_curv.py [nx ny_global nsteps scheme deg kernel]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist

import fr_oracle as o
import fr_oracle_curv as c
import frb200 as FR

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 33
nyg = int(sys.argv[2]) if len(sys.argv) > 2 else 16
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
scheme = sys.argv[4] if len(sys.argv) > 4 else "ssprk3"
deg = int(sys.argv[5]) if len(sys.argv) > 5 else 2
kernel = sys.argv[6] if len(sys.argv) > 6 else "auto"
g = 5.0 / 3.0
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()

v = c.parallelogram_vertices(nx, nyg)
pog = c.CurvSpace2D(v, deg)
n1g, n2g = c.parallelogram_normals(nx, nyg)
# a smooth wave in both directions plus 1 % of seeded noise (so that every halo row matters), admissible
x, y = pog.xpg[..., 0], pog.xpg[..., 1]
rng = np.random.default_rng(7)
rho = 1.0 + 0.1 * np.sin(2 * np.pi * (x - y)) + 0.05 * np.cos(4 * np.pi * y)
prim = np.stack([rho, np.ones_like(rho), 0.2 + 0.05 * np.sin(4 * np.pi * y), rho], axis=-1)
ug = o.prim_conserve(prim, g)
ug = np.asfortranarray(ug * (1.0 + 0.01 * rng.standard_normal(ug.shape)))
c.ghost_fill_periodic(ug)

sl = FR.partition.slab(nyg, world, rank)
rows = slice(sl.start - 1, sl.stop + 2)  # the rank's rows with one halo row on either side
zl = np.zeros((nx + 2, sl.count + 2))
psl = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, nx, 0.0, 0.5, sl.count, zl, zl, zl, zl, v[:, rows]), deg)
assert np.array_equal(psl.iJ, pog.iJ[:, rows])  # per-element metric: the slab's space carries the global rows
ul = np.asfortranarray(ug[:, rows].copy())
n1l = np.asfortranarray(n1g[:, sl.start - 1: sl.stop])
n2l = np.asfortranarray(n2g[:, sl.start - 1: sl.stop + 1])
prob = FR.DistributedEuler2DCurv(ul, (0.0, 1.0), psl, g, dist, n1l, n2l, corr="sp", fy_index="k",
                                 ctx=FR.Context(local), ghost="periodic", kernel=kernel)
alg = {"euler": FR.Euler, "midpoint": FR.Midpoint, "ssprk3": FR.SSPRK33}[scheme]()
dt = 2e-4
prob.step(alg, dt, nsteps)
res = prob.download()
parts = [None] * world
dist.all_gather_object(parts, (sl.start, sl.count, res[:, 1:-1].copy()))
ok = True
if rank == 0:
    rhs = lambda w: c.rhs_euler2d_curv(w, pog, n1g, n2g, g, corr="sp", fy_index="k")  # noqa: E731
    ref = o.integrate(ug, dt, nsteps, rhs, scheme, before_step=c.ghost_fill_periodic)
    got = np.zeros_like(ref)
    for st, cnt, arr in parts:
        got[:, st: st + cnt] = arr
    err = np.abs(got[1:-1, 1:-1] - ref[1:-1, 1:-1]).max() / np.abs(ref).max()
    moved = np.abs(ref[1:-1, 1:-1] - ug[1:-1, 1:-1]).max() / np.abs(ref).max()
    print(f"check_dist_curv world={world} nx={nx} ny={nyg} steps={nsteps} {scheme} deg={deg} {kernel}: "
          f"rel err = {err:.3e} (state moved {moved:.2e})")
    ok = bool(err <= 1e-10 and np.isfinite(got).all())
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
prob.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
