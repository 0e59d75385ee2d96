"""bench.py's reference arm runs without a GPU: one JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-n", "32"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "DOF-updates/s"
    assert d["metric"].startswith("FP64 DOF-updates/s per RK stage")
    for k in ("value", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"] > 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-n", "32"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_curv_probe_record_has_the_roofline_object():
    """scripts/probe_curv.py prints bench.py's roofline object for the curvilinear path (pure function, no GPU)."""
    import importlib.util as u
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = u.spec_from_file_location("probe_curv", os.path.join(root, "scripts", "probe_curv.py"))
    m = u.module_from_spec(spec)
    spec.loader.exec_module(m)
    rec = m.line("w", 67108864, 16, 0.5, 2.7e9)
    roof = rec["roofline"]
    assert rec["stage_bytes_per_dof"] == 24 and abs(rec["gdof_per_s"] - 134.22) < 0.01
    assert roof["bound"] == "hbm" and roof["unit"] == "GB/s" and roof["traffic"] == 2.7e9
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-4 and 3000 < roof["achieved"] < 3400
    assert "oracle" not in open(os.path.join(root, "scripts", "probe_curv.py")).read().replace("only tests/ may execute the oracle", "")
    spec = u.spec_from_file_location("cpu_baseline_curv", os.path.join(root, "tests", "harness", "cpu_baseline_curv.py"))
    c = u.module_from_spec(spec)
    spec.loader.exec_module(c)
    cpu = c.cpu_baseline(2, 48, 32, 1)  # the C/OpenMP restatement of the same residual, timed (tiny sample here)
    assert cpu["kind"] == "port" and cpu["unit"] == "DOF-updates/s" and cpu["value"] > 0 and cpu["cores"] >= 1


def test_reference_arm_of_the_other_configurations():
    """bench.py --impl reference --config {1,2,4}: the same line shape for the other BASELINE configurations."""
    for cfg in ("1", "2", "4"):
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", cfg,
                            "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        d = json.loads([ln for ln in r.stdout.splitlines() if ln.strip()][-1])
        assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
        assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "DOF-updates/s"


def test_clock_sampler_counts_only_the_samples_inside_the_timed_window():
    """bench.ClockSampler: nvidia-smi lines are stamped when read; only those between mark_begin and mark_end count,
    with the warm-up samples under load as the stated fallback when the window caught none."""
    import importlib.util as u
    import time

    spec = u.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    b = u.module_from_spec(spec)
    spec.loader.exec_module(b)

    class FakeProc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    def line(sm, power, cap):
        return f"2026/10/17 20:00:00.000, 0, {sm}, 1965, {power}, 0x4, Not Active, Not Active, Not Active, {cap}"

    t = time.time()
    s = b.ClockSampler(0)
    s.proc = FakeProc()
    s.lines = [(t - 1.0, line(345, 120.0, "Not Active")), (t - 0.5, line(1965, 600.0, "Not Active")),
               (t + 0.01, line(1900, 650.0, "Active")), (t + 0.03, line(1850, 660.0, "Active")),
               (t + 1.0, line(345, 110.0, "Not Active"))]
    s.t0, s.t1 = t, t + 0.05
    r = s.stop()
    assert r["samples"] == 2 and r["sm_mhz"] in (1900.0, 1850.0) and r["reasons"] == ["sw_power_cap"]
    assert r["window"] == "timed region" and r["sm_max_mhz"] == 1965.0
    s2 = b.ClockSampler(0)
    s2.proc = FakeProc()
    s2.lines = s.lines[:2]
    s2.t0, s2.t1 = t, t + 0.05
    r2 = s2.stop()
    assert r2["samples"] == 1 and r2["sm_mhz"] == 1965.0 and r2["window"].startswith("warm-up")
