import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, frb200 as FR
o = FR.examples
ps = FR.FRPSpace1D(0.0, 1.0, 8192, 2)
vs = FR.VSpace1D(-5.0, 5.0, 256)
velo, wts = vs.u, vs.weights
prob = FR.BGKProblem(o.ic_bgk1d(ps, velo), (0, 1), ps, velo, wts, 1e-2)
for kind in (0, 1):
    prob.time_stage(kind, 2); print(kind, prob.time_stage(kind, 5)*1e3, "us")
prob.close()
