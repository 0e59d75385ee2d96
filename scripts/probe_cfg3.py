"""Quick device-time probe of the 2-D Euler stage kernels (not the bench contract)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import frb200 as FR

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
kernels = sys.argv[2].split(",") if len(sys.argv) > 2 else ["rc", "march", "generic"]
g = 5.0 / 3.0
ps = FR.FRPSpace2D(0.0, 1.0, n, 0.0, 1.0, n, 3, 1, 1)
t0 = time.time()
rho = 1.0 + 0.1 * np.sin(2 * np.pi * ps.xpg[..., 0])
u0 = np.empty(rho.shape + (4,), order="F")
u0[..., 0] = rho
u0[..., 1] = rho
u0[..., 2] = 0.0
u0[..., 3] = 0.5 * rho / rho / (g - 1) + 0.5 * rho
print("host init %.1fs" % (time.time() - t0), flush=True)
prob = FR.Euler2DProblem(u0, (0.0, 1.0), ps, g)
dofs = prob.dofs
for k in kernels:
    prob.set_kernel(k)
    for kind, nbytes in ((0, 16), (1, 24)):
        prob.time_stage(kind, 3)
        ms = prob.time_stage(kind, 10)
        print(f"{k:8s} stage_kind={kind} {ms:8.3f} ms  {dofs/ms/1e6:8.2f} GDOF/s  {dofs*nbytes/ms/1e6:8.1f} GB/s algorithmic", flush=True)
prob.close()
