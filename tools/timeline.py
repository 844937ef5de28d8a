"""Hand-off timeline of the sdf-only tensor-core chain (development probe: I2SDF_DEBUG_TIMELINE selects the instrumented
instantiation tc_sdf8_kernel<true>): clock64 stamps of CTA 0's second tile.  Per op: when the epilogue warps start waiting for the
accumulator / see it / publish their first and last chunk, and when the MMA warp sees each A chunk, issues each k step and commits."""
import os
import sys

os.environ["I2SDF_DEBUG_TIMELINE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs  # noqa: E402
from i2sdf_b200.network import I2SDFNetwork  # noqa: E402

torch.manual_seed(0)
m = I2SDFNetwork(configs.model_conf("synthetic")).cuda().eval()
core = m._ready_core()
MAIN = len(sys.argv) > 1 and sys.argv[1] == "main"
if MAIN:        # the full main pass (21 ops for synthetic.yml): usage  python tools/timeline.py main
    R = 1024
    g = torch.Generator().manual_seed(1)
    o = torch.zeros(R, 3, device="cuda")
    o[:, 2] = -1.5
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=1).cuda()
    z = torch.sort(torch.rand(R, 98, generator=g) * 6.0, dim=1).values.cuda()
    with torch.no_grad():
        m.density.beta.fill_(0.01)
    for _ in range(3):
        core.render(o, d, torch.ones(R, device="cuda"), z, m.density.beta.detach())
    NOPS = 21
else:
    pts = (torch.rand(131072, 3, device="cuda") - 0.5) * 3
    for _ in range(3):
        core.sdf_forward(pts)
    NOPS = 8
torch.cuda.synchronize()
raw = core._ws[:65536].view(torch.int64).cpu()
ep = raw[:NOPS * 16 * 4].reshape(NOPS, 16, 4)
issue = raw[2048:2048 + NOPS * 32].reshape(NOPS, 32)
seen = raw[4096:4096 + NOPS * 32].reshape(NOPS, 32)
commit = raw[6144:6144 + NOPS]
t0 = int(ep[0, :, 1].min())
print("clocks relative to 'accumulator of the op complete' (first epilogue warp sees d_full)")
print("op | acc complete (+period) | epilogue: wait starts min..max | first chunk published min..max | last chunk published min..max || MMA of the NEXT op: "
      "first A chunk seen | k steps issued: first, last | commit")
prev = None
for op in range(NOPS):
    base = int(ep[op, :, 1].min())
    w0, w2, w3 = ep[op, :, 0] - base, ep[op, :, 2] - base, ep[op, :, 3] - base
    line = f"{op}  | {base - t0:7d} (+{(base - prev) if prev is not None else 0:5d}) | {int(w0.min()):6d}..{int(w0.max()):6d} | {int(w2.min()):5d}..{int(w2.max()):5d} | {int(w3.min()):5d}..{int(w3.max()):5d}"
    if op + 1 < NOPS:
        n = 16
        line += f" || {int(seen[op + 1, 0]) - base:6d} | {int(issue[op + 1, 0]) - base:6d} {int(issue[op + 1, n - 1]) - base:6d} | {int(commit[op + 1]) - base:6d}"
    print(line)
    prev = base
for op in ((2, 5, 10, 14, 17) if MAIN else (2, 5)):
    b0 = int(ep[op - 1, :, 1].min())
    print(f"MMA warp, op {op}: per k step, clocks since the previous op's accumulator completed: A chunk seen (even steps) / issued; step length")
    last = None
    for k in range(16):
        i = int(issue[op, k]) - b0
        s = f"{int(seen[op, k]) - b0:6d}" if k % 2 == 0 else "      "
        print(f"   ks {k:2d}: seen {s}  issued {i:6d}  (+{(i - last) if last is not None else 0:4d})")
        last = i
