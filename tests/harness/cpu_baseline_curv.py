"""CPU baseline of the curvilinear path: the C / OpenMP restatement of the residual (oracle/fr_oracle_curv.c; Julia is
not in the image, so kind = "port") timed on a bounded sample with all host threads.  Lives under tests/ because it
executes the oracle; the GPU side of the comparison is scripts/probe_curv.py.

    python tests/harness/cpu_baseline_curv.py [deg nx ny evals]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import frb200 as FR  # noqa: E402  (host mirror only: spaces and initial data; no GPU call is made here)


def cpu_baseline(deg, nx=512, ny=256, evals=3):
    import time

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle

    base = FR.PSpace2D(0.0, 1.0, nx, 0.0, 1.0, ny, 1, 1)
    v = base.vertices.copy()
    v[..., 0] += v[..., 1]
    z = np.zeros((nx + 2, ny + 2))
    ps = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, nx, 0.0, 1.0, ny, z, z, z, z, v), deg)
    n1, n2 = FR.face_normals(ps.vertices)
    rho = 1.0 + 0.1 * np.sin(2 * np.pi * (ps.xpg[..., 0] - ps.xpg[..., 1]))
    u = np.asfortranarray(FR.prim_conserve(np.stack([rho, np.ones_like(rho), 0.2 * np.ones_like(rho), rho], -1), 5.0 / 3.0))
    c_oracle.rhs_euler2d_curv(u, ps, n1, n2, 5.0 / 3.0, fy_index="k")
    t0 = time.perf_counter()
    for _ in range(evals):
        c_oracle.rhs_euler2d_curv(u, ps, n1, n2, 5.0 / 3.0, fy_index="k")
    dt = (time.perf_counter() - t0) / evals
    dofs = nx * ny * (deg + 1) ** 2 * 4
    return {"value": dofs / dt, "unit": "DOF-updates/s", "cores": c_oracle.num_threads(), "kind": "port",
            "sample": f"{evals} RHS evaluations of the curvilinear residual on {nx}x{ny} p{deg} sheared elements, "
                      "C/OpenMP restatement of dev/parallelogram.jl:80-165"}


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:5]]
    print(json.dumps({"cpu_baseline": cpu_baseline(*(a if a else [3]))}))
