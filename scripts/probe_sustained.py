"""Does the stage time depend on how long the GPU has been under load?  Back-to-back launches of the 24-B stage of
cfg3 in bursts of 5, 20, 100 and 400 (frb_time_stage), with nvidia-smi clocks / power sampled alongside."""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import frb200 as FR

n = 2048
g = 5.0 / 3.0
ps = FR.FRPSpace2D(0.0, 1.0, n, 0.0, 1.0, n, 3, 1, 1)
rho = 1.0 + 0.1 * np.sin(2 * np.pi * ps.xpg[..., 0])
u0 = np.empty(rho.shape + (4,), order="F")
u0[..., 0] = rho; u0[..., 1] = rho; u0[..., 2] = 0.0; u0[..., 3] = 0.5 / (g - 1) + 0.5 * rho
prob = FR.Euler2DProblem(u0, (0.0, 1.0), ps, g)
lines = []
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap",
                      "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
threading.Thread(target=lambda: [lines.append((time.time(), ln.strip())) for ln in p.stdout], daemon=True).start()
time.sleep(0.3)
for kind in (1, 0):
    prob.time_stage(kind, 3)
    for iters in (5, 20, 100, 400):
        time.sleep(0.5)  # idle between bursts
        t0 = time.time()
        ms = prob.time_stage(kind, iters)
        t1 = time.time()
        s = [ln for t, ln in lines if t0 <= t <= t1 + 0.03]
        print(f"stage_kind {kind} burst {iters:4d}: {ms:.4f} ms per launch; smi samples in the burst: {s[:2]} ... {s[-2:]}", flush=True)
p.terminate()
prob.close()
