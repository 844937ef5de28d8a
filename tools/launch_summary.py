"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share per kernel over the LAST `frac` of launches."""
import sys, csv, collections, re
path = sys.argv[1]; nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) ]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
L = []
for r in rows[1:]:
    if len(r) <= vi: continue
    v = float(r[vi].replace(",", "")); u = r[ui]
    v = v / 1000 if u in ("ns", "nsecond") else v * (1000 if u in ("ms", "msecond") else 1)
    L.append((re.sub(r"\(.*", "", r[ki])[:60], v))
# one full step = the launches between two consecutive sampler_init_kernel launches (a cyclic shift of a step: the tail of one
# step and the head of the next); falls back to the last 1/nsteps of the list
marks = [i for i, (k, _) in enumerate(L) if "sampler_init_kernel" in k]
if len(marks) >= 2:
    L = L[marks[-2]:marks[-1]]
else:
    L = L[-(len(L) // nsteps):]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in L:
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v for _, v in L)
print(f"{len(L)} launches, {tot/1000:.3f} ms serialised")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v:9.1f} us {100*v/tot:5.1f}%  x{c:3d}  {k}")
