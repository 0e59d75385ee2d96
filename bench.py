#!/usr/bin/env python
"""bench.py -- FP64 DOF-updates/s per RK stage of the fused FR residual + explicit stage.

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 2-D Euler
isentropic wave, FRPSpace2D deg 3, 2048 x 2048 elements per GPU (weak scaling: N GPUs hold a
2048 x 2048N mesh split in row slabs), HLL, SSPRK3, fixed dt, ghost fill as
example/euler2d_wave.jl:127-132.

A "step" is one SSPRK3 time step of the resident state = 3 fused RHS+stage kernel launches
(16 + 24 + 24 algorithmic bytes per DOF) plus the per-step ghost fill; `value` is
3 * interior DOFs * K / (device time of the K steps, max over ranks).

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference --steps K --warmup W    # CPU restatement of the reference path

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GAMMA = 5.0 / 3.0
METRIC = "FP64 DOF-updates/s per RK stage (2D Euler p3)"
UNIT = "DOF-updates/s"
BYTES_PER_DOF_STEP = 64.0  # 16 + 24 + 24 over the three SSPRK3 stages


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "dominant_kernel.json")) as fh:
            return json.load(fh).get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_ic(FR, n, ny_local, y_offset_rows, ny_global):
    """isentropic x wave of euler2d_wave.jl:115-120 on this rank's slab (host, NumPy)."""
    import numpy as np

    ps = FR.FRPSpace2D(0.0, 1.0, n, 0.0, ny_global / n, ny_local, 3, 1, 1)  # dx = dy = 1/n
    rho = 1.0 + 0.1 * np.sin(2 * np.pi * ps.xpg[..., 0])
    u0 = np.empty(rho.shape + (4,), order="F")
    u0[..., 0] = rho                     # prim = [rho, 1, 0, lambda=rho]  ->  p = 1/2
    u0[..., 1] = rho
    u0[..., 2] = 0.0
    u0[..., 3] = 0.5 / (GAMMA - 1.0) + 0.5 * rho
    return ps, u0


def cpu_baseline(sample_n=512, evals=3):
    """The C restatement of the reference CPU path (oracle, kind 'port'), all host threads,
    on a bounded sample of the same workload: RHS + one stage axpy per evaluation."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    import frb200 as FR

    ps, u0 = make_ic(FR, sample_n, sample_n, 0, sample_n)
    work = c_oracle.Work2D(sample_n, sample_n, 4)
    du = np.empty_like(u0, order="F")
    c_oracle.rhs_euler2d(u0, ps, GAMMA, work, du)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    for _ in range(evals):
        c_oracle.rhs_euler2d(u0, ps, GAMMA, work, du)
        u1 = u0 + 1e-9 * du  # the stage update OrdinaryDiffEq does outside f!
    dt = time.perf_counter() - t0
    dofs = sample_n * sample_n * 64
    del u1
    return {"value": dofs * evals / dt, "unit": UNIT, "cores": c_oracle.num_threads(), "kind": "port",
            "sample": f"{evals} RHS+stage evaluations of the same workload at {sample_n}x{sample_n} elements "
                      f"(1/{(2048 // sample_n) ** 2} of the per-GPU mesh), C/OpenMP restatement of "
                      "example/shock-vortex.jl:26-118 (Julia is not installed)"}


def run_reference(args):
    """--impl reference: the reference's CPU path (C restatement, all host threads)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    import frb200 as FR

    n = args.ref_n
    ps, u = make_ic(FR, n, n, 0, n)
    dt = 1e-5 * 2048 / n
    for _ in range(args.warmup):
        u = c_oracle.integrate_euler2d(u, ps, GAMMA, dt, 1, "ssprk3", "wave_x")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        u = c_oracle.integrate_euler2d(u, ps, GAMMA, dt, 1, "ssprk3", "wave_x")
    el = time.perf_counter() - t0
    dofs = n * n * 64
    value = 3.0 * dofs * args.steps / el
    cores = c_oracle.num_threads()
    sample = (f"each step = one SSPRK3 step (3 RHS + stage axpys) at {n}x{n} elements, a bounded sample of the "
              "2048x2048 workload; C/OpenMP restatement of the reference CPU path (Julia not installed)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 0, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"2D Euler isentropic wave, FRPSpace2D deg 3, HLL, SSPRK3 (sample {n}x{n})"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "finite": bool(np.isfinite(u).all()),
    }))


def run_ours(args):
    import numpy as np

    # stdout carries exactly one JSON line: everything libraries print while the job runs (NCCL's
    # version banner goes to fd 1) is sent to stderr, the line is written to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import frb200 as FR

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod

        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    n = args.n
    ny_global = n * world
    ps, u0 = make_ic(FR, n, n, rank * n, ny_global)
    ctx = FR.Context(local)
    dt = 1e-5 * 2048 / n
    if world > 1:
        prob = FR.DistributedEuler2D(u0, (0.0, 1.0), ps, GAMMA, dist, ctx=ctx, ghost="wave_x")
    else:
        prob = FR.Euler2DProblem(u0, (0.0, 1.0), ps, GAMMA, ctx=ctx)
        prob.set_hooks(ghost="wave_x")
    alg = FR.SSPRK33()
    dofs = prob.dofs

    def barrier():
        if dist is not None:
            import torch

            dist.barrier()
            torch.cuda.synchronize()

    prob.step(alg, dt, args.warmup)
    prob.set_profiling(True)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    prob.step(alg, dt, args.steps)  # synchronous; CUDA events bracket the K steps on the library stream
    barrier()
    ms, launches = prob.last_timing()
    stage_ms, stage_n = prob.stage_timing()
    clocks = sampler.stop() if rank == 0 else None
    prob.set_profiling(False)
    if dist is not None:
        import torch

        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = 3.0 * dofs * world * args.steps / (ms * 1e-3)

    fin = bool(np.isfinite(prob.download()).all())

    # ---- end to end: the f!(du,u,p,t) call with HOST buffers (pinned), H2D + D2H inside.  Under
    # torchrun every rank evaluates the residual of its own slab (the host array carries the halo
    # rows, as the reference's ghost cells do), all ranks at once: whole-job DOFs / max time.
    uh = FR.pinned_empty(u0.shape)
    dh = FR.pinned_empty(u0.shape)
    uh[...] = u0
    prob.f_pipelined(dh, uh, None, 0.0, nslab=args.e2e_slabs)  # warm-up
    k = max(1, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        prob.f_pipelined(dh, uh, None, 0.0, nslab=args.e2e_slabs)
    el = time.perf_counter() - t0
    if dist is not None:
        import torch

        t = torch.tensor([el], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    nbytes = int(u0.size) * 8
    e2e = {"value": dofs * world * k / el, "unit": UNIT, "h2d_bytes_per_step": nbytes * world,
           "d2h_bytes_per_step": nbytes * world,
           "call": f"frb_rhs_pipelined(prob, u_host, du_host, {args.e2e_slabs}): the f!(du,u,p,t) shape with pinned "
                   "host buffers; upload, fused residual and download overlapped in row slabs"
                   + (f"; {world} ranks, one slab each, concurrently" if world > 1 else ""),
           "ms_per_call": 1e3 * el / k}
    if world == 1:
        # the reference's own user-loop shape (euler2d_wave.jl:125-135): the state lives on the host and is
        # touched between steps, so every SSPRK3 step is upload + 3 fused stages + download
        prob.upload(uh); prob.step(alg, dt, 1); prob.download(dh)
        t0 = time.perf_counter()
        for _ in range(k):
            prob.upload(uh)
            prob.step(alg, dt, 1)
            prob.download(dh)
        el2 = time.perf_counter() - t0
        e2e["user_loop"] = {"value": 3.0 * dofs * k / el2, "unit": UNIT, "ms_per_step": 1e3 * el2 / k,
                            "call": "frb_state_upload + frb_step(SSPRK3, 1 step) + frb_state_download per step "
                                    "(pinned host state, mutated by the user between steps)"}
    FR.pinned_free(uh)
    FR.pinned_free(dh)

    if rank == 0:
        peak, how = measured_peak()
        avg_ms = stage_ms / max(stage_n, 1)
        achieved = dofs * (BYTES_PER_DOF_STEP / 3.0) / (avg_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(), "kernel": "euler2d_rc_kernel<4> (row-chunk layout, 3 launches per SSPRK3 step)", "avg_launch_ms": avg_ms,
                "launches_timed": stage_n, "peak_source": how,
                "algorithmic_bytes_per_launch": dofs * BYTES_PER_DOF_STEP / 3.0}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"2D Euler isentropic wave, FRPSpace2D deg 3, {n}x{n} elements per GPU "
                                   f"({n}x{ny_global} global), HLL, SSPRK3 fixed dt, ghost fill per step",
                       "state_bytes_per_gpu": int(u0.size) * 8, "l2": "inputs (2.15 GB per buffer) larger than L2",
                       "partition": f"row slabs x{world}" if world > 1 else "single GPU"},
            "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "finite": fin,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.cpu_n, args.cpu_evals)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    prob.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=2048, help="elements per side per GPU")
    ap.add_argument("--ref-n", type=int, default=512, help="sample size of the reference arm")
    ap.add_argument("--cpu-n", type=int, default=1024)
    ap.add_argument("--cpu-evals", type=int, default=24)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-slabs", type=int, default=32)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
