"""Print the headline fields of bench.py JSON lines (one file per argument)."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f).read().strip().split("\n"):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if "unavailable" in d:
            print(f, d)
            continue
        e2e = d.get("e2e", {})
        print(f"{f}: {d.get('impl', 'ours')} n_gpus={d.get('n_gpus')} {d['value']:.4g} {d['unit']}  {d['ms_per_step']:.3f} ms/step  e2e {e2e.get('value', 0):.4g}  "
              f"roofline {d.get('roofline', {}).get('frac')}  clocks {d.get('clocks')}")
        if "kernel_ms_per_step" in d:
            print("    kernels:", {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
        if "render" in d:
            print("    render:", round(d["render"]["ms_per_step"], 3), "ms")
