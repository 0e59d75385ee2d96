"""bench.py --config {1,2,4,5,f2}: the BASELINE.json configurations other than the headline cfg3, and the
curvilinear path, in the same line shape (`value`, `roofline`, `e2e`, `cpu_baseline`).  The only use of oracle/ here is the
CPU leg (`cpu_baseline`, `--impl reference`), exactly as in bench.py.

  cfg1  1-D advection, deg 2, 100 periodic cells, SSPRK3            (example/advection_lowlevel.jl)
  cfg2  1-D Euler Sod, deg 3, 4096 cells, Midpoint + positivity limiter every step (src/Equation/eq_euler.jl:29-98)
  cfg4  1-D BGK, deg 2, 8192 cells x 256 velocities, Midpoint       (example/bgk_wave.jl)
  cfg5  2-D NS cavity (gas-kinetic flux), deg 3, 1024^2, Euler forward (example/ns_cavity.jl)
  f2    2-D Euler on a sheared (curvilinear) 1024^2 p3 mesh, SSPRK3  (dev/parallelogram.jl)

None of these shards (a stage of cfg1/2/4 is shorter than one NVLink epoch, cfg5/f2 slabs are
measured by scripts/): under torchrun every rank runs an independent replica ("replicas only") and `value` is
the sum.  A "step" is one time step of the scheme the reference's script uses; `value` = stages * DOFs * K / device
time.  `roofline.achieved` = algorithmic bytes (16 B for u' = u + dt L(u), 24 B for a stage that also
reads u_n; + 8 B of metric on f2) / mean device time of a fused stage, measured with CUDA events around every
stage (frb_set_profiling); cfg5 is FP64-compute-bound and reports executed FP64 instructions of the committed
ncu capture against the FP64 pipe peak instead.
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
G = 5.0 / 3.0
UNIT = "DOF-updates/s"


def _peak():
    import bench

    return bench.measured_peak()


def _profile(name):
    try:
        with open(os.path.join(ROOT, "profiles", "config_kernels.json")) as fh:
            return json.load(fh).get(name, {})
    except Exception:
        return {}


PARITY_TESTS = {"1": "test_advection_rhs, test_advection_1000_steps: 1e-12 / 1e-9",
                "2": "test_euler1d_cfg2_full_size_with_limiter: 1000 Midpoint steps at 4096 cells, 1e-9",
                "4": "test_cfg4_full_size_bgk, test_cfg4_full_size_bgk_against_the_long_double_arbiter",
                "5": "test_cfg5_full_size_ns_rhs: 1024^2 RHS, 1e-12",
                "f2": "tests/test_gpu_curv.py: 148 cases over the three stage kernels incl. 1000 steps, 1e-9"}


# ---- problem builders: (prob, alg, dt, stages, stage_bytes [per stage], hooks, host state) ---------------------
def build(FR, cfg, ctx):
    o = FR.examples
    if cfg == "1":
        ps = FR.FRPSpace1D(-1.0, 1.0, 100, 2)
        u0 = o.ic_advection1d(ps)
        prob = FR.FRAdvectionProblem(u0, (0.0, 2.0), ps, 1.0, "period", variant="lowlevel", ctx=ctx)
        return dict(prob=prob, u0=u0, alg=FR.SSPRK33(), dt=0.05 * ps.dx[0] if np.ndim(ps.dx) else 0.001,
                    bytes=[16, 24, 24], ps=ps,
                    workload="1D linear advection, FRPSpace1D deg 2, 100 periodic cells, SSPRK3 fixed dt "
                             "(example/advection_lowlevel.jl; the script's adaptive Tsit5 is out of scope, SURVEY 0.1)",
                    kernel="loop1d_kernel (the whole time loop in one launch of one CTA; per-stage events do not apply)")
    if cfg == "2":
        ps = FR.FRPSpace1D(0.0, 1.0, 4096, 3)
        u0 = o.ic_sod1d(ps, G)
        prob = FR.FREulerProblem(u0, (0.0, 0.15), ps, G, "dirichlet", ctx=ctx)
        prob.set_hooks(limiter_weights=ps.wp / 2.0)
        return dict(prob=prob, u0=u0, alg=FR.Midpoint(), dt=0.05 / 4096, bytes=[16, 24], ps=ps,
                    workload="1D Euler Sod, FRPSpace1D deg 3, 4096 cells, HLL, Midpoint fixed dt, positivity limiter "
                             "before every step (src/Equation/eq_euler.jl:29-98 + src/dissipation.jl:61-123)",
                    kernel="euler1d_kernel + limiter1d_kernel (CUDA-graph replay of step pairs)")
    if cfg == "4":
        ps = FR.FRPSpace1D(0.0, 1.0, 8192, 2)
        vs = FR.VSpace1D(-5.0, 5.0, 256)
        u0 = o.ic_bgk1d(ps, vs.u)
        prob = FR.BGKProblem(u0, (0.0, 1.0), ps, vs.u, vs.weights, 1e-2, ctx=ctx)
        return dict(prob=prob, u0=u0, alg=FR.Midpoint(), dt=0.1 * (1.0 / 8192) / 5.0, bytes=[16, 24], ps=ps, vs=vs,
                    workload="1D BGK, FRPSpace1D deg 2, 8192 cells x 256 velocities, Midpoint fixed dt "
                             "(example/bgk_wave.jl)",
                    kernel="bgk_moments_kernel + bgk1d_pair_kernel<3> (two launches per stage; two adjacent cells per thread)")
    if cfg == "5":
        n = 1024
        ps = FR.FRPSpace2D(0.0, 1.0, n, 0.0, 1.0, n, 3, 1, 1)
        mu = FR.ref_vhs_vis(1e-3, 1.0, 0.5)
        dt = 0.1 * min(ps.dx, ps.dy) / 3.0
        u0 = o.ic_cavity(ps, G)
        prob = FR.NSCavityProblem(u0, (0.0, 1.0), ps, 1.0, G, mu, 0.81, dt, ctx=ctx)
        return dict(prob=prob, u0=u0, alg=FR.Euler(), dt=dt, bytes=[16], ps=ps, mu=mu,
                    workload="2D Navier-Stokes lid-driven cavity (gas-kinetic flux), FRPSpace2D deg 3, 1024x1024 "
                             "elements, Euler forward (example/ns_cavity.jl scaled)",
                    kernel="ns_boundary_kernel + ns_face_kernel + ns_elem_kernel (3 launches per stage)")
    if cfg == "f2":
        n = 1024
        base = FR.PSpace2D(0.0, 1.0, n, 0.0, 1.0, n, 1, 1)
        v = base.vertices.copy()
        v[..., 0] += v[..., 1]  # 45-degree shear (dev/parallelogram.jl:46-57)
        z = np.zeros((n + 2, n + 2))
        ps = FR.FRPSpace2D(FR.PSpace2D(0.0, 1.0, n, 0.0, 1.0, n, z, z, z, z, v), 3)
        x = ps.xpg[..., 0] - ps.xpg[..., 1]
        rho = 1.0 + 0.1 * np.sin(2 * np.pi * x)
        prim = np.stack([rho, np.ones_like(rho), 0.2 * np.ones_like(rho), rho], axis=-1)
        u0 = np.asfortranarray(FR.prim_conserve(prim, G))
        prob = FR.Euler2DCurvProblem(u0, (0.0, 1.0), ps, G, corr="sp", ctx=ctx)
        prob.set_hooks(ghost="periodic")
        return dict(prob=prob, u0=u0, alg=FR.SSPRK33(), dt=2e-6, bytes=[24, 32, 32], ps=ps,
                    workload="2D Euler on a sheared (45 degree) structured quadrilateral mesh, FRPSpace2D(base, deg 3), "
                             "1024x1024 elements, point-wise iJ, HLL in the face frame, SSPRK3, periodic ghost fill "
                             "(dev/parallelogram.jl:80-165 scaled)",
                    kernel="euler2d_curv_fused_kernel<4, unsigned, HLL> (1 launch per stage, DESIGN 4.4)")
    raise SystemExit(f"unknown config {cfg}")


def cpu_leg(cfg, b, budget_s=12.0):
    """The C/OpenMP (or NumPy where no C form exists) restatement of the same RHS on the host cores, bounded."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import c_oracle
    import fr_oracle as o

    c_oracle.set_num_threads(os.cpu_count() or 1)
    ps, u0 = b["ps"], b["u0"]
    kind = "port"
    if cfg == "1":
        f = lambda: o.rhs_advection1d(u0, ps, 1.0, "period", "lowlevel")  # noqa: E731
        what, cores = "NumPy restatement of example/advection_lowlevel.jl:4-47 (300 DOFs: a scalar loop)", 1
    elif cfg == "2":
        f = lambda: c_oracle.rhs_euler1d(u0, ps, G, "dirichlet")  # noqa: E731
        what, cores = "C/OpenMP restatement of frode_euler!", c_oracle.num_threads()
    elif cfg == "4":
        vs = b["vs"]
        f = lambda: c_oracle.rhs_bgk1d(u0, ps.dx, vs.u, vs.weights, ps.ll, ps.lr, ps.dl, ps.dhl, ps.dhr, 1e-2)  # noqa: E731
        what, cores = "C/OpenMP restatement of mol! (example/bgk_wave.jl:69-129)", c_oracle.num_threads()
    elif cfg == "5":
        # 1024^2 of the GKS residual takes minutes on the host: a 256^2 sample of the same arithmetic
        import frb200 as FR

        psn = FR.FRPSpace2D(0.0, 1.0, 256, 0.0, 1.0, 256, 3, 1, 1)
        un = FR.examples.ic_cavity(psn, G)
        f = lambda: c_oracle.rhs_ns2d(un, psn, 1.0, G, b["mu"], 0.81, b["dt"])  # noqa: E731
        f()
        t0 = time.perf_counter()
        n = 0
        while True:
            f(); n += 1
            if time.perf_counter() - t0 > budget_s or n >= 20:
                break
        el = time.perf_counter() - t0
        return {"value": 256 * 256 * 64 * n / el, "unit": UNIT, "cores": c_oracle.num_threads(), "kind": kind,
                "sample": f"{n} RHS evaluations at 256x256 elements (1/16 of the mesh), C/OpenMP restatement of dudt! + "
                          "boundary! (example/ns_cavity.jl:147-344)"}
    else:  # f2: the curvilinear C restatement lives with the test harness
        import importlib.util as iu

        spec = iu.spec_from_file_location("cpu_baseline_curv", os.path.join(ROOT, "tests", "harness", "cpu_baseline_curv.py"))
        m = iu.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m.cpu_baseline(3, 512, 512, 3)
    f()
    t0 = time.perf_counter()
    n = 0
    while True:
        f(); n += 1
        if time.perf_counter() - t0 > budget_s or n >= 200:
            break
    el = time.perf_counter() - t0
    return {"value": b["prob_dofs"] * n / el, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{n} RHS evaluations of the same workload at full size, {what}"}


def run(args):
    import bench
    import frb200 as FR

    cfg = args.config
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    job = bench.Job()
    ctx = FR.Context(local)
    b = build(FR, cfg, ctx)
    b["ctx"] = ctx
    prob, alg, dt = b["prob"], b["alg"], b["dt"]
    dofs = prob.dofs
    b["prob_dofs"] = dofs
    stages = alg.stages
    # small problems: many steps per timed call, so that the call is not one launch latency
    inner = {"1": 2000, "2": 500, "4": 50, "5": 1, "f2": 1}[cfg]
    # cfg5: forward Euler at the script's dt = cfl min(dx, dy) / 3 is beyond its stability limit on a fine p3 mesh
    # (rounding noise grows ~4 x per step, the C oracle does the same), so the timed steps run in
    # groups of `group` steps from the initial state (re-uploaded outside the timed region): the arithmetic is the
    # same whatever the data, and the state stays finite
    group = 4 if cfg == "5" else None
    prob.step(alg, dt, (group or max(args.warmup, 3) * inner))
    prob.upload(b["u0"])
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)  # nvidia-smi start-up; these timed regions are milliseconds long
    job.barrier()
    sampler.mark_begin()
    nsteps = args.steps * inner
    fin = True
    if group:
        ms, launches, done = 0.0, 0, 0
        while done < nsteps:
            k = min(group, nsteps - done)
            prob.step(alg, dt, k)
            m1, l1 = prob.last_timing()
            ms += m1; launches += l1; done += k
            fin = fin and bool(np.isfinite(prob.download()).all())
            prob.upload(b["u0"])
    else:
        prob.step(alg, dt, nsteps)
        ms, launches = prob.last_timing()
        fin = bool(np.isfinite(prob.download()).all())
    sampler.mark_end()
    job.barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = job.reduce_max(ms)
    value = stages * dofs * world * nsteps / (ms * 1e-3)
    # per-stage timing (events around every stage launch; eager launches, so not for the one-launch loop of cfg1)
    prob.upload(b["u0"])
    prob.set_profiling(True)
    prob.step(alg, dt, min(nsteps, group or 10))
    st_ms, st_n = prob.stage_timing()
    prob.set_profiling(False)
    avg = st_ms / max(st_n, 1)
    peak, how = _peak()
    prof = _profile("cfg" + cfg if cfg != "f2" else "f2")
    bpd = sum(b["bytes"]) / len(b["bytes"])
    if cfg == "5":
        # FP64-compute-bound (erfc / exp / pow in the gas-kinetic flux): executed FP64 instructions per stage from
        # the committed ncu capture against the FP64 pipe peak (148 SMs x 64 lanes x SM clock: B300_MICROARCH /
        # B200_PROFILING.md; not in MEASURED_PEAKS.json)
        inst = prof.get("fp64_thread_inst_per_stage")
        peak_f = 148 * 64 * 1.965e9
        ach = inst / (avg * 1e-3) if inst and avg > 0 else None
        roof = {"bound": "fp64", "achieved": ach / 1e12 if ach else None, "peak": peak_f / 1e12,
                "unit": "T FP64 thread-instr/s", "frac": (ach / peak_f) if ach else None,
                "traffic": prof.get("dram_bytes_per_stage"), "traffic_source": prof.get("source"),
                "peak_source": "148 SMs x 64 FP64 lanes x 1.965 GHz (nominal; MEASURED_PEAKS.json has no FP64 figure)",
                "hbm_frac": dofs * 16 / (avg * 1e-3) / 1e9 / peak if avg > 0 else None}
    else:
        ach = dofs * bpd / (avg * 1e-3) / 1e9 if avg > 0 else None
        roof = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if ach else None,
                "traffic": prof.get("dram_bytes_per_stage"), "traffic_source": prof.get("source"), "peak_source": how,
                "algorithmic_bytes_per_launch": dofs * bpd}
        if cfg in ("1", "2"):
            roof["note"] = (f"latency-bound: the state is {int(b['u0'].size) * 8} bytes (L2-resident, a stage is a few "
                            "microseconds of launch latency); the fraction of HBM peak is reported for completeness")
    roof.update({"kernel": b["kernel"], "avg_stage_ms": avg, "stages_timed": int(st_n)})
    # ---- e2e (as in bench.py): per step upload of the state from pinned host memory + one time step of the
    # script's scheme + download of the result; `f_call` = f!(du, u, p, t) with host buffers
    uh, dh = FR.pinned_empty(b["u0"].shape), FR.pinned_empty(b["u0"].shape)
    uh[...] = b["u0"]
    prob.f(dh, uh)
    k = 5 if cfg in ("5", "f2") else 50
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        prob.f(dh, uh)
    el = job.reduce_max(time.perf_counter() - t0)
    nb = int(b["u0"].size) * 8
    f_call = {"value": dofs * world * k / el, "unit": UNIT, "h2d_bytes_per_call": nb * world,
              "d2h_bytes_per_call": nb * world,
              "call": "frb_rhs(prob, u_host, du_host, t): f!(du,u,p,t) with pinned host buffers", "ms_per_call": 1e3 * el / k}

    def one_step():
        prob.upload(uh)
        prob.step(alg, dt, 1)
        prob.download(dh)

    one_step()
    job.barrier()
    t0 = time.perf_counter()
    for _ in range(k):
        one_step()
    el2 = job.reduce_max(time.perf_counter() - t0)
    e2e = {"value": stages * dofs * world * k / el2, "unit": UNIT, "h2d_bytes_per_step": nb * world,
           "d2h_bytes_per_step": nb * world, "ms_per_step": 1e3 * el2 / k,
           "call": "per step: frb_state_upload + frb_step(1 time step of the script's scheme) + frb_state_download, "
                   "pinned host state (the reference's user loop with the state on the host between steps)",
           "f_call": f_call}
    FR.pinned_free(uh); FR.pinned_free(dh)
    out = None
    if rank == 0:
        out = {"metric": f"FP64 DOF-updates/s per RK stage (cfg{cfg})" if cfg != "f2" else
               "FP64 DOF-updates/s per RK stage (curvilinear 2D Euler p3)",
               "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms / args.steps, "time_steps_per_bench_step": nsteps // args.steps if args.steps else 0,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": b["workload"], "state_bytes_per_gpu": nb,
                          "l2": ("inputs larger than L2" if nb > 126e6 else
                                 "state fits L2 (126 MB): the reference's loop is L2-resident by nature; no flush"),
                          "partition": "replicas only" if world > 1 else "single GPU"},
               "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "finite": fin}
        out["parity"] = {"where": "tests/test_gpu_parity.py (" + PARITY_TESTS[cfg] + "): the oracle is test "
                         "infrastructure and is not called from the measured arm", "finite": fin}
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_leg(cfg, b)
    prob.close()
    job.close()
    return out


def run_reference(args):
    """--impl reference --config X: the CPU restatement of that configuration's RHS on all host cores."""
    import frb200 as FR

    class _NoCtx:  # the reference arm builds no device problem: only the host-side arrays of the builder are used
        pass

    cfg = args.config
    # host-side set-up only (spaces, initial data); no device context is opened
    o = FR.examples
    b = {}
    if cfg == "1":
        b["ps"] = FR.FRPSpace1D(-1.0, 1.0, 100, 2); b["u0"] = o.ic_advection1d(b["ps"])
    elif cfg == "2":
        b["ps"] = FR.FRPSpace1D(0.0, 1.0, 4096, 3); b["u0"] = o.ic_sod1d(b["ps"], G)
    elif cfg == "4":
        b["ps"] = FR.FRPSpace1D(0.0, 1.0, 8192, 2); b["vs"] = FR.VSpace1D(-5.0, 5.0, 256)
        b["u0"] = o.ic_bgk1d(b["ps"], b["vs"].u)
    elif cfg == "5":
        b["ps"] = None; b["u0"] = None; b["mu"] = FR.ref_vhs_vis(1e-3, 1.0, 0.5); b["dt"] = 0.1 / 1024 / 3.0
    else:
        b["ps"] = None; b["u0"] = None
    b["prob_dofs"] = int(b["u0"].size) if b["u0"] is not None else 0
    t0 = time.perf_counter()
    c = cpu_leg(cfg, b, budget_s=max(5.0, 2.0 * args.steps))
    el = time.perf_counter() - t0
    print(json.dumps({"impl": "reference", "metric": f"FP64 DOF-updates/s per RK stage (cfg{cfg})", "value": c["value"],
                      "unit": UNIT, "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": 1e3 * el / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": f"cfg{cfg} RHS"},
                      "cpu_baseline": c,
                      "e2e": {"value": c["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
