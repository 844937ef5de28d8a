#!/usr/bin/env python
"""Static census of the shipped library: per kernel, the SASS mnemonics that prove (or disprove) the tcgen05 / TMEM / bulk-copy path,
local-memory traffic and the ptxas resource lines.  No GPU needed.

    python tools/sass_census.py [path/to/libi2sdf_b200.so] > profiles/<round>_sass_census.txt
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "LDL", "STL", "MUFU", "HMMA", "REDG", "ATOM"]


def demangle(n):
    try:
        d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except OSError:
        d = n
    return re.sub(r"\(.*", "", d).replace("void ", "")


def main():
    so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "i2sdf_b200", "libi2sdf_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    cur, stats = None, collections.OrderedDict()
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            stats[cur] = collections.Counter()
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln) if cur else None
        if m:
            stats[cur][m.group(1).split(".")[0]] += 1
            stats[cur]["_total"] += 1
    res = {}
    for f in glob.glob(os.path.join(ROOT, "i2sdf_b200", "csrc", "build", "*.ptxas.txt")):
        name = None
        for ln in open(f):
            m = re.search(r"Compiling entry function '(\S+)' for 'sm_100a'", ln)
            if m:
                name = m.group(1)
                res[name] = {}
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
            if m and name:
                res[name].update(stack=int(m.group(1)), spill_st=int(m.group(2)), spill_ld=int(m.group(3)))
            m = re.search(r"Used (\d+) registers", ln)
            if m and name:
                res[name]["regs"] = int(m.group(1))
    print(f"SASS census of {os.path.relpath(so, ROOT)} (cuobjdump -sass; static instruction counts per kernel) + ptxas -v resources")
    print("UTCHMMA = tcgen05.mma (kind::f16), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (1-D bulk copy),")
    print("UTMALDG = tensor-map TMA load, SYNCS = mbarrier ops, LDL / STL = local memory, HMMA = legacy mma.sync (must be 0), REDG / ATOM = global reductions / atomics\n")
    print(f"{'kernel':48s} {'instr':>6s} {'regs':>4s} {'stack':>5s} {'spill':>7s} " + " ".join(f"{k:>7s}" for k in KEYS))
    for n, c in stats.items():
        r = res.get(n, {})
        spill = f"{r.get('spill_st', 0)}/{r.get('spill_ld', 0)}"
        print(f"{demangle(n)[:48]:48s} {c['_total']:6d} {r.get('regs', 0):4d} {r.get('stack', 0):5d} {spill:>7s} " + " ".join(f"{c[k]:7d}" for k in KEYS))
    tot = collections.Counter()
    for c in stats.values():
        tot.update(c)
    print("\nlibrary totals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS))


if __name__ == "__main__":
    main()
