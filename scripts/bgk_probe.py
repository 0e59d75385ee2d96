"""cfg4 stage timing (for ncu and quick A/B): 1-D BGK p2, ncell cells x 256 velocities.  python scripts/bgk_probe.py [ncell ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, frb200 as FR
o = FR.examples
for ncell in ([int(a) for a in sys.argv[1:]] or [8192]):
    ps = FR.FRPSpace1D(0.0, 1.0, ncell, 2)
    vs = FR.VSpace1D(-5.0, 5.0, 256)
    velo, wts = vs.u, vs.weights
    prob = FR.BGKProblem(o.ic_bgk1d(ps, velo), (0, 1), ps, velo, wts, 1e-2)
    for kind in (0, 1):
        prob.time_stage(kind, 2)
        us = prob.time_stage(kind, 5) * 1e3
        print(f"ncell {ncell} stage_kind {kind} {us:8.2f} us  {prob.dofs / us / 1e3:8.2f} GDOF/s", flush=True)
    prob.close()
