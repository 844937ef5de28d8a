"""Hand-off / item timeline of the backward chain kernel (development probe: I2SDF_DEBUG_TIMELINE selects tc_bwd8_kernel<true>):
clock64 stamps of CTA 0's second tile.  Per op: when the epilogue warps start waiting for the accumulator / see it / publish their first
item / finish their last one, when the MMA warp issues its first and last k step and commits; and for one epilogue warp every item's
phases: operands (TMEM + slot segments) there, values computed, published, slot stores issued."""
import ctypes as C
import os
import sys

os.environ["I2SDF_DEBUG_TIMELINE"] = "1"
import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from i2sdf_b200 import _lib, configs  # noqa: E402
from i2sdf_b200 import synthetic as syn  # noqa: E402
from i2sdf_b200.network import I2SDFLoss, I2SDFNetwork  # noqa: E402

conf = configs.model_conf("synthetic")
conf["use_normal"] = True
torch.manual_seed(0)
m = I2SDFNetwork(conf)
with torch.no_grad():
    m.density.beta.fill_(0.01)
m = m.cuda().train()
R = 1024
inp = {k: v.cuda() for k, v in syn.synthetic_rays(R, seed=1, train_layout=True).items()}
gt = {k: v.cuda() for k, v in bench.make_train_gt(R, 7).items()}
loss_fn = I2SDFLoss(**configs.LOSS_SYNTHETIC)
for _ in range(3):
    out = m(inp)
    loss = loss_fn(out, gt, 0)["loss"]
    m.zero_grad(set_to_none=True)
    loss.backward()
torch.cuda.synchronize()
lib = _lib.load()
buf = (C.c_int64 * 8192)()
n = lib.i2sdf_debug_bwd_timeline(buf, 8192)
assert n == 8192, "the probe did not run (I2SDF_DEBUG_TIMELINE must be set before the library launches the kernel)"
raw = torch.tensor(list(buf), dtype=torch.int64)
NOPS = 20
KIND = ["T0", "T1", "T2", "T3s", "T4", "T5", "T6", "T7L", "CR3", "CR2", "CR1", "FA", "PF", "P7", "P6", "P5", "P4", "P3s", "P2", "P1"]     # synthetic.yml: 8 + 3 + 1 + 1 + 7
ep = raw[:NOPS * 16 * 4].reshape(NOPS, 16, 4)
issue = raw[2048:2048 + NOPS * 32].reshape(NOPS, 32)
commit = raw[6144:6144 + NOPS]
items = raw[7168:7168 + NOPS * 40].reshape(NOPS, 8, 5)
print("clocks; per op relative to 'accumulator of the op complete' (first epilogue warp sees d_full)")
print("op kind | acc complete (+period) | epilogue wait starts min..max | first item published min..max | last item done min..max || next op's MMA: first k step, last k step, commit")
prev = None
for op in range(NOPS):
    base = int(ep[op, :, 1].min())
    w0, w2, w3 = ep[op, :, 0] - base, ep[op, :, 2] - base, ep[op, :, 3] - base
    line = f"{op:2d} {KIND[op]:4s} | {base - int(ep[0, :, 1].min()):8d} (+{(base - prev) if prev is not None else 0:6d}) | {int(w0.min()):7d}..{int(w0.max()):7d} | {int(w2.min()):6d}..{int(w2.max()):6d} | {int(w3.min()):6d}..{int(w3.max()):6d}"
    if op + 1 < NOPS:
        ks = int((issue[op + 1] != 0).sum())
        if ks:
            line += f" || {int(issue[op + 1, 0]) - base:7d} {int(issue[op + 1, ks - 1]) - base:7d} | {int(commit[op + 1]) - base:7d}  ({ks} k steps)"
    print(line)
    prev = base
print("\nepilogue warp 0 (rows 0-31, column group 0), per item: clocks from item start to [operands there | values computed | published | slot stores issued], then gap to the next item's start")
for op in range(NOPS):
    it = items[op]
    if int(it[0, 0]) == 0:
        continue
    segs = []
    for i in range(8):
        s = it[i]
        gap = int(it[i + 1, 0] - s[4]) if i < 7 else 0
        segs.append(f"{int(s[1] - s[0]):5d} {int(s[2] - s[0]):5d} {int(s[3] - s[0]):5d} {int(s[4] - s[0]):5d} (+{gap:4d})")
    tot = int(it[7, 4] - it[0, 0])
    print(f"{op:2d} {KIND[op]:4s} total {tot:6d} | " + " | ".join(segs[:4]))
    print(f"{'':17s} | " + " | ".join(segs[4:]))
