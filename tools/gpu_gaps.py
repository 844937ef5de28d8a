"""Where is the GPU idle inside a training step?  torch.profiler (CUPTI) trace of a few steady-state steps: kernels in stream order with the idle
gap in front of each one, gaps summed per 'kernel that follows', and the step's busy / idle split."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from i2sdf_b200 import configs  # noqa: E402
from i2sdf_b200.network import I2SDFLoss, I2SDFNetwork  # noqa: E402
from i2sdf_b200.optim import Adam  # noqa: E402
from i2sdf_b200.parallel import GradBucket  # noqa: E402
from i2sdf_b200.synthetic import make_train_gt, synthetic_rays  # noqa: E402

conf = configs.model_conf("synthetic")
conf["use_normal"] = True
torch.manual_seed(0)
m = I2SDFNetwork(conf)
with torch.no_grad():
    m.density.beta.fill_(0.01)
m = m.cuda().train()
R = 1024
inp = {k: v.cuda() for k, v in synthetic_rays(R, seed=1, train_layout=True).items()}
gt = {k: v.cuda() for k, v in make_train_gt(R, 7).items()}
loss_fn = I2SDFLoss(**configs.LOSS_SYNTHETIC)
opt = Adam(m.parameters(), lr=5e-4, eps=1e-15)
bucket = GradBucket(m.parameters())


def step():
    out = m(inp)
    loss = loss_fn(out, gt, 0)["loss"]
    bucket.zero()
    loss.backward()
    bucket.allreduce()
    opt.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
N = 6
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ev.sort(key=lambda e: e.time_range.start)
ks = [(e.time_range.start, e.time_range.end, e.name) for e in ev if "Memcpy" not in e.name and "Memset" not in e.name or True]
t0, t1 = ks[0][0], ks[-1][1]
busy = sum(b - a for a, b, _ in ks)
print(f"{N} steps: wall {1e-3 * (t1 - t0) / N:.3f} ms per step, kernels busy {1e-3 * busy / N:.3f} ms, idle {1e-3 * ((t1 - t0) - busy) / N:.3f} ms")
gaps = {}
prev_end, prev_name = None, None
for a, b, name in ks:
    if prev_end is not None and a > prev_end:
        key = (prev_name[:48], name[:48])
        g = gaps.setdefault(key, [0.0, 0])
        g[0] += a - prev_end
        g[1] += 1
    prev_end, prev_name = max(prev_end or b, b), name
print("largest idle gaps (us per step)  [previous kernel -> next kernel]")
for (pn, nn), (tot, cnt) in sorted(gaps.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"  {tot / N:8.1f} us  x{cnt / N:4.1f}   {pn:48s} -> {nn}")
