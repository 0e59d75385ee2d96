"""Device time of one fused RHS+stage launch for every BASELINE.json config at its full size
(not the bench contract; feeds the per-config table of profiles/ and DESIGN.md)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import frb200 as FR

o = FR.examples  # the example scripts' initial conditions (host mirror)

G = 5.0 / 3.0


def report(name, prob, bytes16, bytes24, extra=""):
    dofs = prob.dofs
    for kind, nb in ((0, bytes16), (1, bytes24)):
        prob.time_stage(kind, 3)
        ms = prob.time_stage(kind, 20)
        print(f"{name:34s} stage_kind={kind} {ms*1e3:9.1f} us  {dofs/ms/1e6:8.2f} GDOF/s  {dofs*nb/ms/1e6:8.1f} GB/s algorithmic {extra}", flush=True)
    prob.close()


ps = FR.FRPSpace1D(-1.0, 1.0, 100, 2)
report("cfg1 adv1d p2 100 cells", FR.FRAdvectionProblem(np.asfortranarray(np.sin(np.pi * ps.xpg)), (0, 1), ps, 1.0, "period", variant="lowlevel"), 16, 24, "(latency-bound: 2.4 KB)")
ps = FR.FRPSpace1D(0.0, 1.0, 4096, 3)
report("cfg2 euler1d p3 4096 cells", FR.FREulerProblem(o.ic_sod1d(ps, G), (0, 1), ps, G, "dirichlet"), 16, 24, "(latency-bound: 393 KB)")
ps = FR.FRPSpace1D(0.0, 1.0, 8192, 2)
vs = FR.VSpace1D(-5.0, 5.0, 256)
velo, wts = vs.u, vs.weights
report("cfg4 bgk1d p2 8192x256", FR.BGKProblem(o.ic_bgk1d(ps, velo), (0, 1), ps, velo, wts, 1e-2), 16, 24, "(50 MB: L2-resident; 2 launches)")
n = int(os.environ.get("NS_N", "1024"))
ps = FR.FRPSpace2D(0.0, 1.0, n, 0.0, 1.0, n, 3, 1, 1)
mu = FR.ref_vhs_vis(1e-3, 1.0, 0.5)
dt = 0.1 * min(ps.dx, ps.dy) / 3.0
report(f"cfg5 ns2d gks p3 {n}^2", FR.NSCavityProblem(o.ic_cavity(ps, G), (0, 1), ps, 1.0, G, mu, 0.81, dt), 16, 24, "(FP64-compute-bound)")
