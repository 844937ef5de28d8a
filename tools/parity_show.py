"""Print a parity report written by tools/parity_report.py as a compact table."""
import json
import sys

r = json.load(open(sys.argv[1]))
for v, rv in r["variants"].items():
    for c, e in rv.items():
        print(v, c)
        print("   out", {k: (f"{x:.2e}" if isinstance(x, float) else x) for k, x in e["outputs"].items()})
        if "sampler" in e:
            print("   smp", {k: (f"{x:.4g}" if isinstance(x, float) else x) for k, x in e["sampler"].items()})
        if "round_decisions_identical_inputs" in e:
            print("   dec", {k: (f"{x:.3g}" if isinstance(x, float) else x) for k, x in e["round_decisions_identical_inputs"].items()})
for k, v in r.items():
    if k not in ("variants",):
        print(k, v if not isinstance(v, dict) else json.dumps(v)[:2000])
