"""Builds the CPU test harness of the curvilinear element routines (tests only; see curv_host.cu)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libcurv_host.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def build(force=False):
    src = os.path.join(HERE, "curv_host.cu")
    hdr = os.path.join(HERE, "..", "..", "fluxreconstruction.jl_b200", "csrc", "frb_euler2d_curv_elem.cuh")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(
        [NVCC, "-O2", "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-ffp-contract=off", "-gencode",
         "arch=compute_100a,code=sm_100a", "-cudart", "static", src, "-o", OUT],
        check=True, capture_output=True, text=True,
    )
    return OUT


if __name__ == "__main__":
    print(build(force=True))
