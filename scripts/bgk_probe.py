import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/oracle")
import numpy as np, fr_oracle as o, frb200 as FR
ps = FR.FRPSpace1D(0.0, 1.0, 8192, 2)
velo, wts = o.vspace1d(-5.0, 5.0, 256)
prob = FR.BGKProblem(o.ic_bgk1d(ps, velo), (0, 1), ps, velo, wts, 1e-2)
for kind in (0, 1):
    prob.time_stage(kind, 2); print(kind, prob.time_stage(kind, 5)*1e3, "us")
prob.close()
