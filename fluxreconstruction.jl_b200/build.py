"""Builds libfrb200.so in-tree with nvcc for sm_100a (no torch, no JIT cache).

    python fluxreconstruction.jl_b200/build.py [--force]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libfrb200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v", "--expt-relaxed-constexpr",
    # nvcc's default host compiler lookup honours PATH; the image's CC override lacks libgomp but is fine here
]

# the latency-bound 1-D kernels keep the reference's literal operation order (see the file header)
PER_FILE_FLAGS = {"frb_kernels_1d.cu": ["-fmad=false"], "frb_rk.cu": ["-fmad=false"]}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "frb200.h"))
    objs, jobs = [], []
    for src in sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(LIBDIR, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        extra = PER_FILE_FLAGS.get(os.path.basename(s), [])
        r = subprocess.run([NVCC, *FLAGS, *extra, "-c", s, "-o", o], capture_output=True, text=True)
        return s, r

    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(f"== {os.path.basename(s)}\n{r.stdout}{r.stderr}\n")
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {s}")
            with open(os.path.join(LIBDIR, os.path.basename(s)[:-3] + ".ptxas.log"), "w") as fh:
                fh.write(r.stderr)
    if force or jobs or _stale(LIB, objs):
        r = subprocess.run(
            [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"],
            capture_output=True, text=True,
        )
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
