"""Kernel table of one training step (torch.profiler / CUPTI): every kernel the step launches - the library's and
PyTorch's - with launches per step and device time per step, plus the idle time of the stream between kernels."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
from i2sdf_b200 import configs
from i2sdf_b200.network import I2SDFNetwork, I2SDFLoss
from i2sdf_b200 import synthetic as orc   # (neutral input generator: perf tools do not touch oracle/)
import bench
name = os.environ.get("CONFIG", "synthetic")
conf = configs.model_conf(name); conf["use_normal"] = True
torch.manual_seed(0)
m = I2SDFNetwork(conf)
with torch.no_grad():
    m.density.beta.fill_(0.01)
m = m.cuda().train()
R = 1024
light = name != "synthetic"
inp = {k: v.cuda() for k, v in orc.synthetic_rays(R, seed=1, train_layout=True).items()}
gt = {k: v.cuda() for k, v in bench.make_train_gt(R, 7, light).items()}
loss_fn = I2SDFLoss(**bench.loss_weights(name)[0])
from i2sdf_b200.optim import Adam
opt = Adam(m.parameters(), lr=5e-4, eps=1e-15)
def step():
    out = m(inp)
    loss = loss_fn(out, gt, 0)["loss"]
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
N = 10
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(N): step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tab = {}
for e in evs:
    t = tab.setdefault(e.name[:90], [0, 0.0])
    t[0] += 1; t[1] += e.time_range.elapsed_us()
busy = sum(v[1] for v in tab.values())
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"{len(evs)/N:.0f} device activities per step; busy {busy/N/1e3:.3f} ms, span {span/N/1e3:.3f} ms per step")
ours = ("tc_", "sampler_", "wgrad_", "planes_", "composite", "rays_kernel", "pack_kernel", "light_", "gemm_", "i2sdf", "sigmoid_adjoint", "colsum", "sum_kernel", "relu_copy", "embed_kernel")
tot_ours = sum(v[1] for k, v in tab.items() if any(o in k for o in ours))
print(f"library kernels {tot_ours/N/1e3:.3f} ms/step, everything else (PyTorch ops, memcpy/memset) {(busy-tot_ours)/N/1e3:.3f} ms/step")
for k, v in sorted(tab.items(), key=lambda kv: -kv[1][1])[:70]:
    print(f"{v[1]/N:9.1f} us  x{v[0]/N:5.1f}  {k}")

# idle time of the stream, by (kernel before the gap -> kernel after the gap)
gaps = {}
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 2:
        t = gaps.setdefault((a.name[:48], b.name[:48]), [0, 0.0])
        t[0] += 1; t[1] += g
print(f"idle gaps > 2 us: {sum(v[1] for v in gaps.values())/N/1e3:.3f} ms per step")
for k, v in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1]/N:9.1f} us  x{v[0]/N:5.1f}  {k[0]}  ->  {k[1]}")
