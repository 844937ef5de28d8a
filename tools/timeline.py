"""Hand-off timeline of the sdf-only tensor-core chain (development probe, I2SDF_DEBUG_TIMELINE): clock64 stamps of CTA 0's second
tile: per op, when each epilogue warp starts waiting for the accumulator / sees it / has published its last item, and when the MMA
thread sees the first / last chunk and issues its last commit.  Prints clocks relative to the op's 'accumulator complete'."""
import os, sys
os.environ["I2SDF_DEBUG_TIMELINE"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs
from i2sdf_b200.network import I2SDFNetwork
torch.manual_seed(0)
m = I2SDFNetwork(configs.model_conf("synthetic")).cuda().eval()
core = m._ready_core()
pts = (torch.rand(131072, 3, device="cuda") - 0.5) * 3
for _ in range(3):
    core.sdf_forward(pts)
torch.cuda.synchronize()
raw = core._ws[:16384].view(torch.int64).cpu()
tl = raw[:1024].reshape(-1, 4)[: 8 * 20].reshape(8, 20, 4)
t0 = int(tl[0, 0, 1])
print("op | acc complete (abs) | warps: wait_start min/max, seen max, last publish min/max | mma: first chunk, last chunk, last commit   [clocks, relative to acc complete of the op]")
prev = None
for op in range(8):
    w = tl[op, :16]
    seen = w[:, 1]
    base = int(seen.min())
    ws, pub = w[:, 0] - base, w[:, 2] - base
    mma = tl[op, 16, :3] - base
    dur = (base - prev) if prev is not None else 0
    print(f"{op}  | {base - t0:7d} (+{dur:5d}) | wait_start {int(ws.min()):6d}..{int(ws.max()):6d}  seen ..{int((seen - base).max()):4d}  publish {int(pub.min()):6d}..{int(pub.max()):6d} | mma first {int(mma[0]):6d} last {int(mma[1]):6d} commit {int(mma[2]):6d}")
    prev = base

ks = raw[1024:1024 + 80].reshape(16, 5)
b0 = int(ks[0, 0])
print("MMA thread, op 2, per k step [clocks from the step's start]: wait a_ready | wait full (weights) | issue MMAs | commit ; start rel. to op start")
for k in range(16):
    r = ks[k]
    print(f"ks {k:2d}: start {int(r[0]) - b0:6d} | a_ready {int(r[1] - r[0]):5d} | full {int(r[2] - r[1]):5d} | mma issue {int(r[3] - r[2]):5d} | commit {int(r[4] - r[3]):5d}")
