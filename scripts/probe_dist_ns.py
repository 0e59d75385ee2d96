"""Weak-scaling probe of the cfg5 column slabs: every rank holds nx_local x ny elements of the cavity.

  python scripts/probe_dist_ns.py [nx_local ny deg nsteps lid]                              # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29521 scripts/probe_dist_ns.py [nx_local ny deg nsteps]                 # N GPUs

Prints one JSON line on rank 0: device ms per Euler step (max over ranks) and whole-job DOF-updates/s.
The gas starts at rest; lid (default 0: the state stays at rest and finite, the kernels do the same work
whatever the data) is the wall speed of ns_cavity.jl:337.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import frb200 as FR


class _Solo:
    def get_rank(self):
        return 0

    def get_world_size(self):
        return 1


def main():
    nxl, ny, deg, nsteps = (int(a) for a in (sys.argv[1:5] + ["1024", "1024", "3", "20"][len(sys.argv) - 1:])[:4])
    lid = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank = dist.get_rank()
    else:
        dist, rank = _Solo(), 0
    g = 5.0 / 3.0
    nxg = nxl * world
    dx = 1.0 / nxg
    ps = FR.FRPSpace2D(rank * nxl * dx, (rank + 1) * nxl * dx, nxl, 0.0, 1.0, ny, deg, 1, 1)
    nsp = deg + 1
    u = np.empty((4, nsp, nsp, ny + 2, nxl + 2), order="F")  # gas at rest, ns_cavity.jl:36-47
    u[...] = FR.prim_conserve(np.array([1.0, 0.0, 0.0, 1.0]), g)[:, None, None, None, None]
    mu = FR.ref_vhs_vis(1e-3, 1.0, 0.5)
    dt = 0.1 * min(dx, 1.0 / ny) / 3.0
    prob = FR.DistributedNSCavity(u, (0.0, 1.0), ps, 1.0, g, mu, 0.81, dt, dist, lid=lid, ctx=FR.Context(local))
    # forward Euler at the script's dt amplifies rounding noise ~4 x per step on a fine p3 mesh:
    # the timed steps run in groups of 4 from the rest state, re-uploaded (and re-exchanged) outside the timed region
    prob.step(FR.Euler(), dt, 2)
    ms, launches, done = 0.0, 0, 0
    while done < nsteps:
        prob.upload(u)
        if world > 1:
            prob.resync()
            dist.barrier()
        k = min(4, nsteps - done)
        prob.step(FR.Euler(), dt, k)
        m1, l1 = prob.last_timing()
        ms += m1; launches += l1; done += k
    if world > 1:
        import torch

        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    res = prob.download()
    fin = bool(np.isfinite(res[:, :, :, 1:-1, 1:-1]).all())  # owned elements (ghost corners are never written)
    bad_ring = int((~np.isfinite(res)).sum() - (~np.isfinite(res[:, :, :, 1:-1, 1:-1])).sum())
    if rank == 0:
        dofs = prob.dofs * world
        print(json.dumps({"workload": f"cfg5 cavity {nxl}x{ny} p{deg} per GPU, column slabs", "n_gpus": world,
                          "ms_per_step": round(ms / nsteps, 4), "gdof_per_s": round(dofs * nsteps / ms / 1e6, 2),
                          "kernels_per_step": launches / nsteps, "lid": lid, "finite": fin, "nonfinite_ghost_values": bad_ring,
                          "max_momentum": float(np.nanmax(np.abs(res[1:3, :, :, 1:-1, 1:-1])))}), flush=True)
    prob.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
