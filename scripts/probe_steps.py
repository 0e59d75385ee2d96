"""Wall / device time of long time loops of the small configs (launch-bound): cfg1, cfg2, cfg4."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import frb200 as FR

o = FR.examples  # the example scripts' initial conditions (host mirror)
G = 5.0 / 3.0


def run(name, prob, alg, dt, n, hooks=None):
    itg = FR.init(prob, alg, dt=dt)
    if hooks:
        itg.set_hooks(**hooks)
    prob.step(alg, dt, 20)
    t0 = time.perf_counter(); prob.step(alg, dt, n); el = time.perf_counter() - t0
    ms, launches = prob.last_timing()
    print(f"{name:30s} {n} steps: wall {el*1e3:8.2f} ms, device {ms:8.2f} ms, {ms*1e3/n:7.2f} us/step, {launches} launches", flush=True)
    prob.close()


ps = FR.FRPSpace1D(-1.0, 1.0, 100, 2)
run("cfg1 adv1d ssprk3", FR.FRAdvectionProblem(np.asfortranarray(np.sin(np.pi * ps.xpg)), (0, 1), ps, 1.0, "period", variant="lowlevel"), FR.SSPRK33(), 1e-3, 1000)
ps = FR.FRPSpace1D(0.0, 1.0, 4096, 3)
run("cfg2 euler1d midpoint+limiter", FR.FREulerProblem(o.ic_sod1d(ps, G), (0, 1), ps, G, "dirichlet"), FR.Midpoint(), 0.05 * ps.dx[0], 1000, {"limiter_weights": ps.wp / 2})
ps = FR.FRPSpace1D(0.0, 1.0, 8192, 2)
vs = FR.VSpace1D(-5.0, 5.0, 256)
velo, wts = vs.u, vs.weights
run("cfg4 bgk1d midpoint", FR.BGKProblem(o.ic_bgk1d(ps, velo), (0, 1), ps, velo, wts, 1e-2), FR.Midpoint(), 0.1 * ps.dx[0] / 5.0, 200)
