"""Test worker (imports the oracle: test infrastructure).  Column-slab parallel cavity (cfg5) parity check (run under torchrun, one rank per GPU): the N-rank result must
equal the single-domain C oracle of example/ns_cavity.jl on the same global mesh.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29513 tests/dist/check_dist_ns.py [nx_global ny deg nsteps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import torch.distributed as dist

import c_oracle
import fr_oracle as o
import frb200 as FR

nxg = int(sys.argv[1]) if len(sys.argv) > 1 else 33
ny = int(sys.argv[2]) if len(sys.argv) > 2 else 20
deg = int(sys.argv[3]) if len(sys.argv) > 3 else 2
nsteps = int(sys.argv[4]) if len(sys.argv) > 4 else 30
g = 5.0 / 3.0
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()

psg = FR.FRPSpace2D(0.0, 1.0, nxg, 0.0, 1.0, ny, deg, 1, 1)
rng = np.random.default_rng(8)
ug = o.ic_cavity(psg, g)
ug = np.asfortranarray(ug * (1.0 + 0.02 * rng.standard_normal(ug.shape)))
mu = FR.ref_vhs_vis(1e-3, 1.0, 0.5)
dt = 0.1 * min(psg.dx, psg.dy) / 3.0
sl = FR.partition.slab(nxg, world, rank)
# local slab with halo columns: global columns start-1 .. stop+1 (i is the last, slowest index)
ul = np.asfortranarray(ug[..., sl.start - 1: sl.stop + 2].copy())
dx = 1.0 / nxg
psl = FR.FRPSpace2D((sl.start - 1) * dx, sl.stop * dx, sl.count, 0.0, 1.0, ny, deg, 1, 1)
assert abs(psl.Jx - psg.Jx) <= 4e-16 * psg.Jx  # (x1 - x0) / count of the slab vs 1 / nx of the mesh: an ulp or two
prob = FR.DistributedNSCavity(ul, (0.0, 1.0), psl, 1.0, g, mu, 0.81, dt, dist, ctx=FR.Context(local))
# one f!(du, u) on the resident slab, then the time loop
du = np.zeros_like(ul, order="F")
prob.rhs_resident(du)
prob.step(FR.Euler(), dt, nsteps)
res = prob.download()
parts = [None] * world
dist.all_gather_object(parts, (sl.start, sl.count, du[..., 1:-1].copy(), res[..., 1:-1].copy()))
ok = True
if rank == 0:
    dref = c_oracle.rhs_ns2d(ug.copy(order="F"), psg, 1.0, g, mu, 0.81, dt)
    ref = c_oracle.integrate_ns2d(ug, psg, 1.0, g, mu, 0.81, dt, nsteps)
    dgot, got = np.zeros_like(dref), np.zeros_like(ref)
    for st, cnt, d, arr in parts:
        dgot[..., st: st + cnt] = d
        got[..., st: st + cnt] = arr
    I = (slice(None),) * 3 + (slice(1, -1), slice(1, -1))
    e1 = np.abs(dgot[I] - dref[I]).max() / np.abs(dref).max()
    e2 = np.abs(got[I] - ref[I]).max() / np.abs(ref).max()
    print(f"check_dist_ns world={world} nx={nxg} ny={ny} deg={deg} steps={nsteps}: rhs rel err = {e1:.3e}, "
          f"after steps = {e2:.3e}")
    ok = bool(e1 <= 1e-12 and e2 <= 1e-10)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
prob.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
