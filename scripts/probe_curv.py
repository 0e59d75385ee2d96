"""Stage timing of the curvilinear-quadrilateral kernels (frb_euler2d_curv_create) on a sheared mesh.

    python scripts/probe_curv.py [nx ny deg iters]

Prints one JSON line per stage kind: ms per fused stage (face + element kernel), DOF-updates/s and the
algorithmic bandwidth (16 / 24 B of state + 8 B of metric per DOF-update).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frb200 as FR  # noqa: E402


def main():
    nx, ny, deg, iters = (int(a) for a in (sys.argv[1:5] + ["1024", "1024", "3", "20"][len(sys.argv) - 1:]))
    g = 5.0 / 3.0
    base = FR.PSpace2D(0.0, 1.0, nx, 0.0, 1.0, ny, 1, 1)
    v = base.vertices.copy()
    v[..., 0] += v[..., 1]  # 45-degree shear
    z = np.zeros((nx + 2, ny + 2))
    ps = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, nx, 0.0, 1.0, ny, z, z, z, z, v), deg)
    x = ps.xpg[..., 0] - ps.xpg[..., 1]
    rho = 1.0 + 0.1 * np.sin(2 * np.pi * x)
    prim = np.stack([rho, np.ones_like(rho), 0.2 * np.ones_like(rho), rho], axis=-1)
    u = np.asfortranarray(FR.prim_conserve(prim, g))
    for corr, metric in (("sp", "stored"), ("fp", "stored"), ("sp", "vertices")):
        prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, g, corr=corr, metric=metric)
        dofs = prob.dofs
        for kind, state_bytes in ((0, 16), (1, 24)):
            prob.time_stage(kind, 3)
            ms = prob.time_stage(kind, iters)
            print(json.dumps({
                "workload": f"curv euler2d {nx}x{ny} p{deg} corr={corr} metric={metric}",
                "stage_bytes_per_dof": state_bytes + 8,
                "ms_per_stage": round(ms, 4), "gdof_per_s": round(dofs / ms / 1e6, 2),
                "algorithmic_GBps": round(dofs * (state_bytes + 8) / ms / 1e6, 1)}), flush=True)
        prob.close()


if __name__ == "__main__":
    main()
