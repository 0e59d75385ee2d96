"""Test worker (imports the oracle: test infrastructure).  Slab-parallel parity check (run under torchrun, one rank per GPU):
the N-rank result must equal the single-domain oracle on the same global mesh.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/dist/check_dist.py [nx ny_global nsteps scheme kernel ghost]
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

nx = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nyg = int(sys.argv[2]) if len(sys.argv) > 2 else 96
nsteps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
scheme = sys.argv[4] if len(sys.argv) > 4 else "ssprk3"
kernel = sys.argv[5] if len(sys.argv) > 5 else "auto"
ghost = sys.argv[6] if len(sys.argv) > 6 else "wave_x"
g = 5.0 / 3.0
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()

psg = FR.FRPSpace2D(0.0, 1.0, nx, 0.0, 1.0, nyg, 3, 1, 1)
rng = np.random.default_rng(5)
ug = o.ic_wave2d(psg, g, "x" if ghost == "wave_x" else "y")
ug = np.asfortranarray(ug * (1.0 + 0.01 * rng.standard_normal(ug.shape)))
ug[..., 2] += 0.05 * ug[..., 0]
sl = FR.partition.slab(nyg, world, rank)
# local slab with halo rows: global rows start-1 .. stop+1
ul = np.asfortranarray(ug[:, sl.start - 1: sl.stop + 2].copy())
dy = 1.0 / nyg
psl = FR.FRPSpace2D(0.0, 1.0, nx, (sl.start - 1) * dy, sl.stop * dy, sl.count, 3, 1, 1)
assert abs(psl.Jy - psg.Jy) < 1e-18
prob = FR.DistributedEuler2D(ul, (0.0, 1.0), psl, g, dist, ctx=FR.Context(local), ghost=ghost, kernel=kernel)
alg = {"euler": FR.Euler, "midpoint": FR.Midpoint, "ssprk3": FR.SSPRK33}[scheme]()
dt = 2e-4
prob.step(alg, dt, nsteps)
res = prob.download()
parts = [None] * world
dist.all_gather_object(parts, (sl.start, sl.count, res[:, 1:-1].copy()))
ok = True
if rank == 0:
    ref = c_oracle.integrate_euler2d(ug, psg, g, dt, nsteps, scheme, ghost)
    got = np.zeros_like(ref)
    for st, cnt, arr in parts:
        got[:, st: st + cnt] = arr
    err = np.abs(got[1:-1, 1:-1] - ref[1:-1, 1:-1]).max() / np.abs(ref).max()
    print(f"check_dist world={world} nx={nx} ny={nyg} steps={nsteps} {scheme} {kernel} {ghost}: rel err = {err:.3e}")
    ok = bool(err <= 1e-10)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
prob.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
