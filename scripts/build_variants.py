"""Kernel experiments: build libfrb200 variants that differ only in -D flags of one source.

    python scripts/build_variants.py frb_euler2d_march.cu v1="-DMARCH_OPT=1" v2="-DMARCH_OPT=2" ...

Each variant lands in fluxreconstruction.jl_b200/lib/variants/libfrb200_<name>.so (git-ignored,
travels to the GPU box); select it at run time with FRB200_LIB=<path>.
"""
import concurrent.futures as cf
import importlib.util as u
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = u.spec_from_file_location("frb200_build", os.path.join(ROOT, "fluxreconstruction.jl_b200", "build.py"))
B = u.module_from_spec(spec)
spec.loader.exec_module(B)


def main():
    src = sys.argv[1]
    variants = dict(a.split("=", 1) for a in sys.argv[2:])
    B.build_lib()
    vdir = os.path.join(B.LIBDIR, "variants")
    os.makedirs(vdir, exist_ok=True)
    others = [os.path.join(B.LIBDIR, f[:-3] + ".o") for f in B.sources() if f != src]

    def one(item):
        name, flags = item
        o = os.path.join(vdir, f"{src[:-3]}_{name}.o")
        r = subprocess.run([B.NVCC, *B.FLAGS, *B.PER_FILE_FLAGS.get(src, []), *flags.split(), "-c",
                            os.path.join(B.CSRC, src), "-o", o], capture_output=True, text=True)
        if r.returncode:
            return name, r.stderr
        with open(os.path.join(vdir, f"{src[:-3]}_{name}.ptxas.log"), "w") as fh:
            fh.write(r.stderr)
        lib = os.path.join(vdir, f"libfrb200_{name}.so")
        r = subprocess.run([B.NVCC, "-shared", "-o", lib, o, *others, "-gencode", "arch=compute_100a,code=sm_100a",
                            "-cudart", "static"], capture_output=True, text=True)
        return name, r.stderr if r.returncode else "ok " + lib

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for name, msg in ex.map(one, variants.items()):
            print(name, msg)


if __name__ == "__main__":
    main()
