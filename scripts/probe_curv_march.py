"""A / B of the curvilinear stage kernels on one problem: the two-kernel form against the
one-launch marching kernel (FRB_CURV_MARCH), and the marching kernel over segment lengths (FRB_CURV_ROWS).

    python scripts/probe_curv_march.py [nx ny deg iters]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import frb200 as FR  # noqa: E402
from probe_curv import line  # noqa: E402


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
    prob = FR.Euler2DCurvProblem(u, (0.0, 1.0), ps, g, corr="sp", metric="stored")
    dofs = prob.dofs
    rows = [int(r) for r in os.environ.get("PROBE_ROWS", "0,16,32,43,64,86,128,256").split(",")]
    variants = [("one launch (default)", {}), ("two kernels", {"FRB_CURV_TWO_KERNELS": "1"})] + [
        (f"march rows={r or 'auto'}", {"FRB_CURV_MARCH": "1"} | ({"FRB_CURV_ROWS": str(r)} if r else {})) for r in rows]
    for name, env in variants:
        for k in ("FRB_CURV_MARCH", "FRB_CURV_ROWS", "FRB_CURV_TWO_KERNELS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        for kind, state_bytes in ((0, 16), (1, 24)):
            prob.time_stage(kind, 3)
            ms = prob.time_stage(kind, iters)
            rec = line(f"curv euler2d {nx}x{ny} p{deg} {name}", dofs, state_bytes, ms)
            print(json.dumps({k: rec[k] for k in ("workload", "stage_bytes_per_dof", "ms_per_stage", "gdof_per_s")}
                             | {"frac": rec["roofline"]["frac"]}), flush=True)
    prob.close()


if __name__ == "__main__":
    main()
