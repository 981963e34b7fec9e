"""Static evidence from the built library: per kernel, the SASS mnemonics that identify the Blackwell paths
(B200_PROFILING.md "What proves a Blackwell-native kernel"), plus registers / spills from cuobjdump's resource usage.
    python scripts/sass_summary.py > profiles/r02_sass_summary.md        (no GPU needed)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "lattice_net_b200", "liblattice_b200.so")
KEYS = [("UTC*MMA (tcgen05.mma)", r"\bUTC[A-Z]*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"),
        ("UTMALDG (TMA tensor)", r"\bUTMALDG"), ("LDGSTS (cp.async)", r"\bLDGSTS"), ("SYNCS (mbarrier)", r"\bSYNCS"),
        ("REDG (no-return atomics)", r"\bREDG?\."), ("REDG .F32x4 (16-byte reductions)", r"\bREDG?\.[A-Za-z0-9.]*F32x4"),
        ("ATOMG (CAS / add with return)", r"\bATOMG?\."), ("MATCH (match.any)", r"\bMATCH\."), ("HMMA (legacy mma.sync)", r"\bHMMA"), ("FFMA", r"\bFFMA")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return [re.sub(r"\(.*", "", re.sub(r"^void ", "", n)) for n in out]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.a-z]*)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["_total"] += 1
        for label, pat in KEYS:
            if re.match(pat, op):
                counts[cur][label] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            usage[fn] = (int(m.group(1)), int(m.group(2)))
            fn = None
    names = list(counts)
    pretty = dict(zip(names, demangle(names)))
    print("# SASS summary of `lattice_net_b200/liblattice_b200.so` (cuobjdump, sm_100a)\n")
    print("Counts of the instructions that identify each path; `regs` / static `smem` from `cuobjdump -res-usage`.\n")
    cols = [k for k, _ in KEYS]
    print("| kernel | SASS instrs | regs | " + " | ".join(cols) + " |")
    print("|---|---:|---:|" + "---:|" * len(cols))
    for fn in sorted(names, key=lambda f: pretty[f]):
        c = counts[fn]
        regs = usage.get(fn, ("", ""))[0]
        print(f"| `{pretty[fn][:70]}` | {c['_total']} | {regs} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + " |")


if __name__ == "__main__":
    main()
