"""Slab-parallel (multi-GPU) parity: N ranks == single-domain oracle.  Needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except OSError:
        return 0


@pytest.mark.parametrize("scheme,kernel,ghost", [("ssprk3", "auto", "wave_x"), ("midpoint", "generic", "wave_y"),
                                                 ("euler", "auto", "wave_x"), ("ssprk3", "march", "wave_x"),
                                                 ("midpoint", "rc", "wave_y")])
def test_two_rank_slabs_match_oracle(scheme, kernel, ghost):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist", "check_dist.py"), "64", "96", "20",
           scheme, kernel, ghost]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("nx,ny,deg", [(33, 20, 2), (16, 12, 3)])
def test_two_rank_cavity_slabs_match_oracle(nx, ny, deg):
    """cfg5 split in column slabs: RHS and 30 Euler steps of 2 ranks == the single-domain C oracle."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29519", os.path.join(ROOT, "tests", "dist", "check_dist_ns.py"), str(nx), str(ny),
           str(deg), "30"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
