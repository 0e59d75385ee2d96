"""profiles/sass_summary.txt: what the built libfrb200.so contains, per kernel -- architecture, registers, spills
(from the ptxas logs of the build) and counts of the SASS mnemonics that show how bytes move and where the
arithmetic runs (UBLKCP = cp.async.bulk, UTMALDG = TMA tensor load, SYNCS = mbarrier, DFMA/DMUL/DADD = FP64 pipe,
MUFU = special-function seeds, SHFL, LDG/STG/LDS/STS, ATOM/RED, no HMMA/DMMA/UTCMMA: nothing here is GEMM-shaped).

    python scripts/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fluxreconstruction.jl_b200", "lib")
SO = os.path.join(LIB, "libfrb200.so")
MN = ["UBLKCP", "UTMALDG", "UTMAPF", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "LDG", "STG", "LDS", "STS",
      "ATOM", "RED", "BAR", "HMMA", "DMMA", "UTCMMA", "BRA"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def short(d):
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    return re.sub(r"\(.*$", "", d)


def main():
    regs = {}
    for log in sorted(glob.glob(os.path.join(LIB, "*.ptxas.log"))):
        cur = None
        for ln in open(log):
            m = re.search(r"Compiling entry function '([^']+)'", ln)
            if m:
                cur = m.group(1)
                regs[cur] = {"file": os.path.basename(log).replace(".ptxas.log", ".cu")}
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and cur:
                regs[cur].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers", ln)
            if m and cur:
                regs[cur]["regs"] = int(m.group(1))
                m2 = re.search(r"(\d+) bytes smem", ln)
                regs[cur]["smem"] = int(m2.group(1)) if m2 else 0
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, cur, total = {}, None, collections.Counter()
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            counts[cur]["_instr"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["_instr"] += 1
            if op in MN:
                counts[cur][op] += 1
                total[op] += 1
    dm = demangle(list(counts))
    print(f"libfrb200.so: {len(counts)} kernels, arch {', '.join(arch)}; totals: "
          + ", ".join(f"{k} {total[k]}" for k in MN if total[k]))
    print("tensor-core mnemonics (HMMA/DMMA/UTCMMA):", total["HMMA"] + total["DMMA"] + total["UTCMMA"],
          "-- by design: contractions are <= 4 wide per point (DESIGN section 4)")
    print()
    hdr = f"{'kernel':72s} {'file':24s} {'regs':>4s} {'spill':>7s} {'smem':>6s} {'instr':>6s} " + " ".join(f"{k:>6s}" for k in MN[:16])
    print(hdr)
    for k in sorted(counts, key=lambda x: (regs.get(x, {}).get("file", ""), dm[x])):
        r = regs.get(k, {})
        c = counts[k]
        print(f"{short(dm[k])[:72]:72s} {r.get('file', '?')[:24]:24s} {r.get('regs', 0):4d} "
              f"{r.get('spill_st', 0):3d}/{r.get('spill_ld', 0):<3d} {r.get('smem', 0):6d} {c['_instr']:6d} "
              + " ".join(f"{c[m]:6d}" for m in MN[:16]))


if __name__ == "__main__":
    main()
