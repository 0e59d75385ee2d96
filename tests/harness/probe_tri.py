"""Timing of the triangle residual (dev/sod.jl:31-123) -- thread-per-cell gather kernels, measured so that
the row has a number (profiles/r02_summary.md).  Lives under tests/ because the mesh generator and the space
builder it uses (tri_mesh_rect, tri_space) are the oracle's.

    python tests/harness/probe_tri.py [nx ny deg]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

import fr_oracle as o  # noqa: E402
import fr_oracle_tri as T  # noqa: E402
import frb200 as FR  # noqa: E402

nx, ny, deg = (int(a) for a in (sys.argv[1:4] + ["96", "96", "2"][len(sys.argv) - 1:]))
g = 5.0 / 3.0
t0 = time.time()
pts, cells = T.tri_mesh_rect(nx, ny, jitter=0.2, seed=3)
sp = T.tri_space(pts, cells, deg)
x, y = sp["xpg"][..., 0], sp["xpg"][..., 1]
rho = 1.0 + 0.2 * np.sin(2 * np.pi * x) * np.cos(2 * np.pi * y)
prim = np.stack([rho, 0.5 + 0.1 * y, -0.25 + 0.1 * x, 1.0 / (0.8 + 0.2 * rho)], axis=-1)
u = np.asfortranarray(o.prim_conserve(prim, g))
setup = time.time() - t0
prob = FR.TriEulerProblem(u, (0.0, 0.1), sp["cellType"], sp["J"], sp["lf"], sp["normals"], sp["fpn"], sp["dl"],
                          sp["phi"], g, fpn_base=0)
dofs = prob.dofs
for kind, nb in ((0, 16), (1, 24)):
    prob.time_stage(kind, 3)
    ms = prob.time_stage(kind, 20)
    print(json.dumps({"workload": f"tri euler {len(cells)} cells deg {deg}", "stage_bytes_per_dof": nb,
                      "ms_per_stage": round(ms, 5), "gdof_per_s": round(dofs / ms / 1e6, 3),
                      "algorithmic_GBps": round(dofs * nb / ms / 1e6, 1), "host_setup_s": round(setup, 1)}), flush=True)
prob.close()
