"""Top stall sites from `ncu -i X.ncu-rep --page source --csv [--kernel-name ...]` (first kernel in the csv).
usage: python tools/ncu_src.py file.csv [N]   prints the N (default 40) SASS instructions with the most stall samples, with their
dominant stall reasons and executed counts, plus the sampled share of a few instruction classes."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
# several kernels may be concatenated: take the first block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    body.append(r)
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
print(f"{len(body)} SASS instructions, {tot} samples")
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))
for i in order[:N]:
    r = body[i]
    n = int(r[col["# Samples"]] or 0)
    st = sorted(((int(r[col[h]] or 0), h[6:]) for h in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100.0 * n / tot:5.2f}%  exec {r[col['Instructions Executed']]:>9s}  {r[col['Source']][:70]:70s} " + " ".join(f"{h}:{v}" for v, h in st if v))
classes = {"MUFU": 0, "SYNCS": 0, "UTCHMMA": 0, "LDTM": 0, "STS": 0, "FENCE": 0, "MEMBAR": 0, "LDG": 0, "UTCBAR": 0, "BAR": 0, "WARPSYNC": 0, "F2FP": 0, "HADD2": 0}
ex = dict(classes)
for r in body:
    for k in classes:
        if k in r[col["Source"]].split("(")[0].upper().split()[0:3].__str__():
            classes[k] += int(r[col["# Samples"]] or 0)
            ex[k] += int(r[col["Instructions Executed"]] or 0)
print({k: f"{100.0 * v / tot:.1f}% of samples, {ex[k]} executed" for k, v in classes.items() if v})
