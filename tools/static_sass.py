"""Static evidence from the built library (no GPU needed): registers / stack / shared memory per kernel
(`cuobjdump -res-usage`) and the SASS mnemonics that show which kernels use the bulk-copy (TMA) engine, mbarriers
and 16-byte global accesses (`cuobjdump -sass`).  Writes a markdown table.

    python tools/static_sass.py > profiles/r02_static_sass.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "setintersectionprojection.jl_b200", "libsipb200.so")
MNEMONICS = ["UBLKCP", "SYNCS", "LDG.E.128", "STG.E.128", "LDG.E.64", "REDUX", "SHFL", "ATOMS", "ATOMG", "RED.E", "DFMA", "DADD", "F2F"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True, check=True).stdout.splitlines()
    return dict(zip(names, out))


def short(d):
    d = re.sub(r"^void ", "", d)
    d = re.sub(r"\(.*$", "", d)               # drop the argument list
    d = d.replace("(anonymous namespace)::", "")
    if d.startswith("cub::"):                 # library code (histogram set's radix sort): keep the kernel name and key type
        key = "u64 keys" if "policy_hub<unsigned long long" in d else "u32 keys"
        d = re.sub(r"CUB_\d+_SM_\d+::", "", d.split("<")[0]) + " (" + key + ")"
    return d


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            usage[cur] = tuple(int(v) for v in m.groups())
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    ninstr = collections.Counter()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if not m:
            continue
        ninstr[cur] += 1
        ins = m.group(1)
        for mn in MNEMONICS:
            if re.search(r"(^|\s)" + re.escape(mn) + r"(\.|\s|$)", ins) or (("." in mn) and mn in ins):
                counts[cur][mn] += 1
    names = sorted(usage)
    dm = demangle(names)
    print("# Static resource usage and SASS mnemonics of `libsipb200.so` (sm_100a)\n")
    print("`python tools/static_sass.py` on the library built by `__graft_entry__.build()` (nvcc 12.9, `-O3 -fmad=false "
          "-lineinfo`).  No GPU involved: this is what the compiler emitted, not a measurement.\n")
    tot = collections.Counter()
    for f in names:
        for mn, c in counts[f].items():
            tot[mn] += c
    print("Whole library: %d kernels, %d SASS instructions; " % (len(names), sum(ninstr.values()))
          + ", ".join("`%s` %d" % (mn, tot[mn]) for mn in MNEMONICS) + ".\n")
    print("`UBLKCP` = `cp.async.bulk` (the TMA engine without a tensor map), `SYNCS` = mbarrier arrive / try_wait.  "
          "Kernels with a non-zero STACK spill.\n")
    print("| kernel (template arguments kept) | regs | stack B | static smem B | SASS instr | UBLKCP | SYNCS | LDG.128 | STG.128 |")
    print("|---|---|---|---|---|---|---|---|---|")
    for f in sorted(names, key=lambda f: short(dm[f])):
        r, s, sh = usage[f]
        c = counts[f]
        print("| `%s` | %d | %d | %d | %d | %d | %d | %d | %d |" % (short(dm[f]), r, s, sh, ninstr[f], c["UBLKCP"], c["SYNCS"],
                                                               c["LDG.E.128"], c["STG.E.128"]))
    spill = [short(dm[f]) for f in names if usage[f][1] > 0]
    print("\nKernels with stack use: %d of %d%s" % (len(spill), len(names), (" — " + ", ".join("`%s`" % s for s in spill)) if spill else ""))


if __name__ == "__main__":
    sys.exit(main())
