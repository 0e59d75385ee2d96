"""CPU checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/frb200.h declares, matches the ctypes signatures, does not link the oracle, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "frb200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(frb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(FR):
    lib = FR.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in frb200.h but not exported by libfrb200.so"


def test_ctypes_signatures_cover_the_header(FR):
    assert sorted(FR.SIGNATURES) == declared_symbols()


def test_library_is_native_sm100a_and_free_of_the_oracle(FR):
    nm = subprocess.run(["nm", "-D", FR.LIB_PATH], capture_output=True, text=True).stdout
    assert "fro_" not in nm
    ldd = subprocess.run(["ldd", FR.LIB_PATH], capture_output=True, text=True).stdout
    assert "fr_oracle" not in ldd and "torch" not in ldd
    sass = subprocess.run(["cuobjdump", "-lelf", FR.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_march_kernel_uses_tma_and_fp64_pipe(FR):
    """SASS evidence (B200_PROFILING.md): UTMALDG = cp.async.bulk.tensor, DFMA = FP64 FMA."""
    out = subprocess.run(["cuobjdump", "-sass", FR.LIB_PATH], capture_output=True, text=True).stdout
    seg = out[out.index("euler2d_march_kernel"):]
    assert "UTMALDG" in seg and "DFMA" in seg and "SYNCS" in seg


def test_no_gpu_fails_loudly(FR):
    from conftest import HAS_GPU

    if HAS_GPU:
        pytest.skip("a GPU is present")
    with pytest.raises(FR.FRBError) as ei:
        FR.Context()
    assert "no CPU fallback" in str(ei.value)
    ps = FR.FRPSpace1D(0.0, 1.0, 8, 2)
    import numpy as np

    with pytest.raises(FR.FRBError):
        FR.FREulerProblem(np.ones((8, 3, 3), order="F"), (0, 1), ps, 5 / 3, "period")


def test_argument_errors_do_not_need_a_gpu(FR):
    lib = FR.lib()
    out = ctypes.c_void_p()
    assert lib.frb_ctx_create(0, None) == -1  # FRB_ERR_ARG
    assert b"NULL" in lib.frb_last_error(None)
    assert lib.frb_step(None, 0, 0.1, 1) == -1
    assert lib.frb_state_len(None) == 0
    assert lib.frb_prob_destroy(None) == 0
    # the entry points added for the curvilinear path, the kinetic model and the column slabs
    assert lib.frb_euler2d_curv_create(None, 4, 4, None, None, None, None, None, 0, 1.4, ctypes.byref(out)) == -1
    assert b"NULL" in lib.frb_last_error(None)
    assert lib.frb_euler2d_curv_set_vertices(None, None, None) == -3  # FRB_ERR_STATE: not a curvilinear problem
    assert lib.frb_bgk1d_set_model(None, 1, 1.0) == -1
    assert lib.frb_halo_export(None, None) == -1


def test_curv_kernels_are_in_the_library(FR):
    """The curvilinear path is CUDA in libfrb200.so (face + element kernel), not host code."""
    out = subprocess.run(["cuobjdump", "-sass", FR.LIB_PATH], capture_output=True, text=True).stdout
    for name in ("euler2d_curv_face_kernel", "euler2d_curv_elem_kernel", "ghost_cyl_theta_kernel",
                 "halo_push_cols_kernel"):
        assert name in out, name
    seg = out[out.index("euler2d_curv_elem_kernel"):]
    assert "DFMA" in seg[:400000] and "BAR.SYNC" in seg[:400000]


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fluxreconstruction.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "fr_oracle" not in txt and "c_oracle" not in txt and "fro_" not in txt, f


def test_julia_shim_binds_every_entry_point():
    """julia/FRB200.jl (the host language the north star names; Julia is not in the image, so the file cannot
    be executed here) has a ccall for every function include/frb200.h declares, with as many argument types as
    the ctypes twin that the tests exercise."""
    import re as _re

    src = open(os.path.join(ROOT, "julia", "FRB200.jl")).read()
    for name in declared_symbols():
        assert f"(:{name}, lib)" in src, f"{name} has no ccall in julia/FRB200.jl"
    assert src.count("(") == src.count(")") and src.count("[") == src.count("]")

    def split_top(t):
        out, d, cur = [], 0, ""
        for ch in t:
            d += ch in "([{"
            d -= ch in ")]}"
            if ch == "," and d == 0:
                out.append(cur.strip())
                cur = ""
            else:
                cur += ch
        return out + ([cur.strip()] if cur.strip() else [])

    import frb200 as FR

    for m in _re.finditer(r"ccall\(\(:(frb_[a-z0-9_]+), lib\),", src):
        i = j = m.start() + len("ccall")
        d = 0
        while True:  # the matching parenthesis of ccall(
            d += src[j] == "("
            d -= src[j] == ")"
            if d == 0:
                break
            j += 1
        parts = split_top(src[i + 1: j])  # (:name, lib), Ret, (types...), args...
        want = len(FR.SIGNATURES[m.group(1)][1])
        assert len(split_top(parts[2][1:-1])) == want and len(parts) - 3 == want, m.group(1)
    blocks = len(_re.findall(r"^\s*(?:function|struct|mutable struct|module)\b", src, flags=_re.M))
    blocks += len(_re.findall(r"=\s*function\s*\(", src))
    assert blocks == len(_re.findall(r"^\s*end\b", src, flags=_re.M))


def test_only_test_infrastructure_touches_the_oracle():
    """oracle/ is test infrastructure: besides tests/, only __graft_entry__ (build + smoke) and the CPU legs of
    bench.py may import it -- no script, no package module."""
    offenders = []
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in (".git", "oracle", "tests", "gpurun_out", "__pycache__", "lib")]
        for f in files:
            if not f.endswith((".py", ".sh", ".jl", ".cu", ".cuh", ".h")):
                continue
            path = os.path.join(dirpath, f)
            if os.path.relpath(path, ROOT) in ("bench.py", "bench_configs.py", "__graft_entry__.py"):  # bench_configs:
                # the --config arms of bench.py (CPU legs only, checked below)
                continue
            txt = open(path).read()
            if "fr_oracle" in txt or "c_oracle" in txt or 'os.path.join(ROOT, "oracle")' in txt:
                offenders.append(os.path.relpath(path, ROOT))
    assert offenders == []
    # and inside the bench the oracle is only ever reached from the CPU legs
    import ast

    for name, allowed in (("bench.py", {"_oracle", "cpu_baseline", "run_reference"}),
                          ("bench_configs.py", {"cpu_leg", "run_reference"})):
        tree = ast.parse(open(os.path.join(ROOT, name)).read())
        for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
            uses = any(isinstance(n, ast.Name) and n.id in ("c_oracle", "fr_oracle") for n in ast.walk(fn)) or any(
                isinstance(n, (ast.Import, ast.ImportFrom)) and any("oracle" in a.name for a in n.names)
                for n in ast.walk(fn))
            assert not uses or fn.name in allowed, (name, fn.name)


def test_argument_checks_of_the_host_mirror_come_before_the_device(FR):
    """Wrong array shapes are ValueErrors of the host mirror (the reference: BoundsError / DimensionMismatch) and
    are raised before a device context is opened, so they can be checked here."""
    import numpy as np

    g = 5.0 / 3.0
    ps1 = FR.FRPSpace1D(0.0, 1.0, 8, 2)
    with pytest.raises(ValueError):
        FR.FREulerProblem(np.ones((8, 4, 3), order="F"), (0, 1), ps1, g, "period")  # nsp = 3 expected
    with pytest.raises(ValueError):
        FR.FRAdvectionProblem(np.ones((8, 2), order="F"), (0, 1), ps1, 1.0, "period")
    vs = FR.VSpace1D(-5.0, 5.0, 16)
    with pytest.raises(ValueError):
        FR.BGKProblem(np.ones((8, 12, 3), order="F"), (0, 1), ps1, vs.u, vs.weights)  # 12 velocities vs 16
    ps2 = FR.FRPSpace2D(0.0, 1.0, 6, 0.0, 1.0, 4, 2, 1, 1)
    with pytest.raises(ValueError):
        FR.Euler2DProblem(np.ones((6, 4, 3, 3, 4), order="F"), (0, 1), ps2, g)  # ghost ring missing
    with pytest.raises(ValueError):
        FR.NSCavityProblem(np.ones((4, 3, 3, 4, 6), order="F"), (0, 1), ps2, 1.0, g, 1e-3, 0.81, 1e-4)
    u = np.ones((8, 6, 3, 3, 4), order="F")
    with pytest.raises(ValueError):
        FR.Euler2DCurvProblem(u, (0, 1), ps2, g, n1=np.zeros((6, 4, 2)), n2=np.zeros((6, 5, 2)))  # n1 is [nx+1, ny, 2]
    with pytest.raises(ValueError):
        FR.Euler2DCurvProblem(u, (0, 1), ps2, g, corr="nodal")
    with pytest.raises(ValueError):
        FR.Euler2DCurvProblem(u, (0, 1), FR.FRPSpace2D(0.0, 1.0, 8, 0.0, 1.0, 6, 2, 0, 0), g)  # no ghost ring
