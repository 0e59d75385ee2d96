"""Where the time of the host-resident step goes: host ghost fill, frb_step_host at several slab counts, the plain
upload + step + download, f!(du,u) pipelined (cfg3, pinned host buffers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import frb200 as FR
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
ps, u0 = bench.make_ic(FR, n, n)
prob = FR.Euler2DProblem(u0, (0.0, 1.0), ps, bench.GAMMA)
uh, dh = FR.pinned_empty(u0.shape), FR.pinned_empty(u0.shape)
uh[...] = u0
nx = ny = n
alg, dt = FR.SSPRK33(), 1e-5 * 2048 / n


def fill(u):
    u[0] = u[nx]; u[nx + 1] = u[1]
    u[:, 0] = u[:, ny]; u[:, 0, :, :, 2] *= -1
    u[:, ny + 1] = u[:, 1]; u[:, ny + 1, :, :, 2] *= -1


def timeit(f, k=4):
    f()
    t0 = time.perf_counter()
    for _ in range(k):
        f()
    return 1e3 * (time.perf_counter() - t0) / k


print("host ghost fill        %.2f ms" % timeit(lambda: fill(uh)))
for ns in (8, 16, 32, 64, 128):
    print("frb_step_host nslab %3d %.2f ms" % (ns, timeit(lambda: prob.step_host(uh, dh, alg, dt, nslab=ns))))
print("upload only            %.2f ms" % timeit(lambda: prob.upload(uh)))
print("download only          %.2f ms" % timeit(lambda: prob.download(dh)))
print("f_pipelined nslab 32   %.2f ms" % timeit(lambda: prob.f_pipelined(dh, uh, None, 0.0, nslab=32)))


def plain():
    prob.upload(uh); prob.step(alg, dt, 1); prob.download(dh)


print("upload+step+download   %.2f ms" % timeit(plain))
prob.close()
