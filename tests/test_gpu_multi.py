"""Slab-parallel (multi-GPU) parity: N ranks == single-domain oracle, N = 2, 4, 8.

A box with fewer GPUs than a case needs SKIPS it -- unless FRB_REQUIRE_GPUS=N is set (the multi-GPU gate:
`FRB_REQUIRE_GPUS=8 python -m pytest tests/test_gpu_multi.py -m gpu`), in which case every case needing <= N
GPUs must run and a missing GPU is a FAILURE, not a skip."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRE = int(os.environ.get("FRB_REQUIRE_GPUS", "0") or 0)


def _ngpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except OSError:
        return 0


def _need(world):
    have = _ngpu()
    if have >= world:
        return
    if REQUIRE >= world:
        pytest.fail(f"FRB_REQUIRE_GPUS={REQUIRE}: this case needs {world} GPUs, the box has {have}")
    pytest.skip(f"needs {world} GPUs (box has {have}; set FRB_REQUIRE_GPUS to turn the skip into a failure)")


def _torchrun(world, port, script, *args, timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist", script),
           *[str(a) for a in args]]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("scheme,kernel,ghost", [("ssprk3", "auto", "wave_x"), ("midpoint", "generic", "wave_y"),
                                                 ("euler", "auto", "wave_x"), ("ssprk3", "march", "wave_x"),
                                                 ("midpoint", "rc", "wave_y")])
def test_two_rank_slabs_match_oracle(scheme, kernel, ghost):
    _need(2)
    _torchrun(2, 29517, "check_dist.py", 64, 96, 20, scheme, kernel, ghost)


@pytest.mark.parametrize("world", [4, 8])
@pytest.mark.parametrize("scheme,kernel,ghost", [("ssprk3", "auto", "wave_x"), ("midpoint", "march", "wave_y")])
def test_n_rank_slabs_match_oracle(world, scheme, kernel, ghost):
    """4 and 8 slabs of 96 rows (24 / 12 rows each: every slab is a few segments of the row-chunk kernel)"""
    _need(world)
    _torchrun(world, 29521, "check_dist.py", 64, 96, 20, scheme, kernel, ghost)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_n_rank_slabs_uneven_rows_and_resident_rhs(world):
    """100 rows do not divide by 8 (slabs of 13 and 12 rows); f!(du,u) of the resident slabs afterwards"""
    _need(world)
    _torchrun(world, 29523, "check_dist.py", 91, 100, 10, "ssprk3", "auto", "wave_x", "rhs")


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kernel,hooks", [("auto", "limiter"), ("generic", "limiter"), ("auto", "filter")])
def test_n_rank_slabs_with_step_hooks(world, kernel, hooks):
    """positive_limiter before / modal filter after every step rewrite rows the neighbours hold halo copies of:
    the rows are re-sent (frb_api.cu halo_republish) and N ranks still equal the single domain"""
    _need(world)
    _torchrun(world, 29525, "check_dist.py", 64, 96, 5, "ssprk3", kernel, "wave_x", hooks)


@pytest.mark.parametrize("nyg", [2, 4, 6, 10])
def test_two_rank_slabs_of_one_two_three_rows(nyg):
    """tiny slabs: 1, 2, 3 and 5 owned rows per rank -- the row-chunk kernel's boundary rows are one-row segments of
    their own; with one owned row the same CTA consumes both halo rows and raises both flags"""
    _need(2)
    _torchrun(2, 29531, "check_dist.py", 64, nyg, 6, "ssprk3", "auto", "wave_x")
    _torchrun(2, 29531, "check_dist.py", 31, nyg, 4, "midpoint", "rc", "wave_y")


@pytest.mark.parametrize("nyg", [8, 16])
def test_eight_rank_slabs_of_one_and_two_rows(nyg):
    """eight slabs of one / two rows: every rank's only rows are boundary rows of both neighbours"""
    _need(8)
    _torchrun(8, 29533, "check_dist.py", 64, nyg, 6, "ssprk3", "auto", "wave_x")


@pytest.mark.parametrize("world", [2, 4])
def test_n_rank_slabs_with_the_state_on_the_host_between_steps(world):
    """frb_step_host on a connected problem: upload, halo rows re-sent behind a neighbour barrier, step, download"""
    _need(world)
    _torchrun(world, 29529, "check_dist.py", 64, 96, 4, "ssprk3", "auto", "wave_x", "step_host")


@pytest.mark.parametrize("nx,ny,deg", [(33, 20, 2), (16, 12, 3)])
def test_two_rank_cavity_slabs_match_oracle(nx, ny, deg):
    """cfg5 split in column slabs: RHS and 30 Euler steps of 2 ranks == the single-domain C oracle."""
    _need(2)
    _torchrun(2, 29519, "check_dist_ns.py", nx, ny, deg, 30)


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kernel,scheme,deg", [("auto", "ssprk3", 2), ("generic", "midpoint", 3),
                                               ("curv_march", "euler", 1)])
def test_n_rank_curvilinear_slabs_match_oracle(world, kernel, scheme, deg):
    """Row slabs of a sheared mesh (SURVEY 8f-2 on 8e's partition): every stage kernel of the curvilinear path,
    interior boundaries exchanged per stage, the periodic seam per step; 33 x 16 elements, 20 steps, 1e-10."""
    _need(world)
    _torchrun(world, 29541, "check_dist_curv.py", 33, 16, 20, scheme, deg, kernel)


@pytest.mark.parametrize("world", [4, 8])
def test_n_rank_cavity_slabs_match_oracle(world):
    _need(world)
    _torchrun(world, 29527, "check_dist_ns.py", 33, 20, 2, 30)
