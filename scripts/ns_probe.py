"""cfg5 stage timing only (for ncu): 2-D NS cavity, gas-kinetic flux, p3."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import frb200 as FR

o = FR.examples  # the example scripts' initial conditions (host mirror)
G = 5.0 / 3.0
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ps = FR.FRPSpace2D(0.0, 1.0, n, 0.0, 1.0, n, 3, 1, 1)
prob = FR.NSCavityProblem(o.ic_cavity(ps, G), (0, 1), ps, 1.0, G, FR.ref_vhs_vis(1e-3, 1.0, 0.5), 0.81, 0.1 * ps.dx / 3.0)
for kind in (0, 1):
    prob.time_stage(kind, 2)
    print(kind, prob.time_stage(kind, 5) * 1e3, "us")
prob.close()
