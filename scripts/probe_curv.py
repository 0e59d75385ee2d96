"""Stage timing of the curvilinear-quadrilateral kernels (frb_euler2d_curv_create) on a sheared mesh.

    python scripts/probe_curv.py [nx ny deg iters]

Prints one JSON line per stage kind: ms per fused stage (the kernel frb_set_kernel selects), DOF-updates/s and a
`roofline` object like bench.py's: achieved = algorithmic bytes (16 / 24 B of state + 8 B of metric per
DOF-update) / stage time, peak = the measured HBM copy bandwidth of MEASURED_PEAKS.json
(fallback: B200_PROFILING.md), traffic = the DRAM bytes per stage of the committed ncu capture
(profiles/r02_summary.md section F; 1024^2 p3, 16-B stage only).  The CPU restatement of the same residual is timed by
tests/harness/cpu_baseline_curv.py (only tests/ may execute the oracle).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frb200 as FR  # noqa: E402


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def line(workload, dofs, state_bytes, ms, traffic=None):
    """The JSON record of one measurement (pure function: tests/test_bench_contract.py checks it on the CPU)."""
    peak, src = measured_peak()
    achieved = dofs * (state_bytes + 8) / ms / 1e6  # GB/s
    return {"workload": workload, "stage_bytes_per_dof": state_bytes + 8, "ms_per_stage": round(ms, 4),
            "gdof_per_s": round(dofs / ms / 1e6, 2), "algorithmic_GBps": round(achieved, 1),
            "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": src,
                         "kernel": "euler2d_curv_fused_kernel (1 launch per stage; vertex metric: face + element kernel)"}}


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
            captured = (nx, ny, deg, kind, corr, metric) == (1024, 1024, 3, 0, "sp", "stored")
            print(json.dumps(line(f"curv euler2d {nx}x{ny} p{deg} corr={corr} metric={metric}", dofs, state_bytes, ms,
                                  1.625e9 if captured else None)), flush=True)
        prob.close()


if __name__ == "__main__":
    main()
