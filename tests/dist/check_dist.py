"""Test worker (imports the oracle: test infrastructure).  Slab-parallel parity check (run under torchrun, one rank per GPU):
the N-rank result must equal the single-domain oracle on the same global mesh.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29511 tests/dist/check_dist.py [nx ny_global nsteps scheme kernel ghost hooks]

hooks: "none" | "step_host" (frb_step_host per step) | "limiter" (positive_limiter before every step, shock-vortex.jl:298-303) | "filter" (modal filter
after every step, :308-321) | "rhs" (also f!(du,u) of the resident slabs against the oracle residual)
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
hooks = sys.argv[7] if len(sys.argv) > 7 else "none"
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
wl = None
if hooks == "limiter":
    # a negative density point in cells either side of every slab boundary: the limiter rewrites exactly the rows
    # the neighbours hold halo copies of (as in test_euler2d_limiter_hook_acts_like_the_oracle_limiter)
    for j in range(1, nyg + 1):
        if j % 12 in (0, 1):  # rows either side of every slab boundary of 2, 4 and 8 slabs of 96 rows
            for i in (5, 30, 31, 47):
                ug[i, j, 1, 2, 0] = -0.02
                ug[i, j, 1, 2, 1:3] = 0.0
    wl = psg.wp / 4.0
sl = FR.partition.slab(nyg, world, rank)
# local slab with halo rows: global rows start-1 .. stop+1
ul = np.asfortranarray(ug[:, sl.start - 1: sl.stop + 2].copy())
dy = 1.0 / nyg
psl = FR.FRPSpace2D(0.0, 1.0, nx, (sl.start - 1) * dy, sl.stop * dy, sl.count, 3, 1, 1)
assert abs(psl.Jy - psg.Jy) <= 4e-16 * psg.Jy  # slab extent / count vs 1 / ny: an ulp or two
prob = FR.DistributedEuler2D(ul, (0.0, 1.0), psl, g, dist, ctx=FR.Context(local), ghost=ghost, kernel=kernel)
alg = {"euler": FR.Euler, "midpoint": FR.Midpoint, "ssprk3": FR.SSPRK33}[scheme]()
dt = 2e-4
if wl is not None:
    prob.set_hooks(ghost=ghost, limiter_weights=wl)
if hooks == "filter":
    prob.modal_filter(psl, 1e-3, when="after")
if hooks == "step_host":
    # the user loop with the state on the host between steps: every rank calls frb_step_host on its slab (upload,
    # halo rows re-sent behind a neighbour barrier, step with the device ghost hook, download)
    res = ul.copy(order="F")
    for _ in range(nsteps):
        prob.step_host(res, res, alg, dt)
else:
    prob.step(alg, dt, nsteps)
    res = prob.download()
du_l = None
if hooks == "rhs":
    du_l = np.zeros_like(ul, order="F")
    prob.rhs_resident(du_l)
parts = [None] * world
dist.all_gather_object(parts, (sl.start, sl.count, res[:, 1:-1].copy(), None if du_l is None else du_l[:, 1:-1].copy()))
ok = True
if rank == 0:
    if hooks == "filter":
        ref = ug.copy(order="F")
        for _ in range(nsteps):
            ref = c_oracle.integrate_euler2d(ref, psg, g, dt, 1, scheme, ghost)
            o.filter_pass_2d(ref, psg.V, psg.iV, 3, 1e-3)
    else:
        ref = c_oracle.integrate_euler2d(ug, psg, g, dt, nsteps, scheme, ghost, limiter_weights=wl)
    got = np.zeros_like(ref)
    for st, cnt, arr, _ in parts:
        got[:, st: st + cnt] = arr
    err = np.abs(got[1:-1, 1:-1] - ref[1:-1, 1:-1]).max() / np.abs(ref).max()
    print(f"check_dist world={world} nx={nx} ny={nyg} steps={nsteps} {scheme} {kernel} {ghost} {hooks}: rel err = {err:.3e}")
    ok = bool(err <= 1e-10)
    if hooks == "rhs":
        # residual of the resident slabs: the halo rows hold the neighbours' rows of the final state; the global
        # seam rows hold the ghost rows of the last step's fill (frozen), which the single-domain state also has
        gotu = ref.copy(order="F")
        gotu[1:-1, 1:-1] = got[1:-1, 1:-1]
        dref = c_oracle.rhs_euler2d(gotu, psg, g)
        dgot = np.zeros_like(dref)
        for st, cnt, _, darr in parts:
            dgot[:, st: st + cnt] = darr
        inner = slice(2, -2)  # rows next to the seam see ghost rows of different age in the two set-ups
        derr = np.abs(dgot[1:-1, inner] - dref[1:-1, inner]).max() / np.abs(dref).max()
        print(f"check_dist world={world} resident RHS rel err = {derr:.3e}")
        ok = ok and bool(derr <= 1e-10)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
prob.close()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
