"""One line per run: training-step time (CUDA events, L2 not flushed) and the library's per-kind kernel times (core.profile) on the
1024-ray C2 batch.  Meant for env-switch sweeps:  for v in 0 1 2 8; do I2SDF_BWD_PREFETCH=$v python tools/step_times.py; done"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs
from i2sdf_b200.network import I2SDFNetwork, I2SDFLoss
from i2sdf_b200 import synthetic as syn
from i2sdf_b200.optim import Adam
import bench
name = os.environ.get("CONFIG", "synthetic")
light = name != "synthetic"
conf = configs.model_conf(name); conf["use_normal"] = True
torch.manual_seed(0)
m = I2SDFNetwork(conf)
with torch.no_grad():
    m.density.beta.fill_(0.01)
m = m.cuda().train()
R = int(os.environ.get("R", 1024))
inp = {k: v.cuda() for k, v in syn.synthetic_rays(R, seed=1, train_layout=True).items()}
gt = {k: v.cuda() for k, v in bench.make_train_gt(R, 7, light).items()}
loss_fn = I2SDFLoss(**bench.loss_weights(name)[0])
opt = Adam(m.parameters(), lr=float(os.environ.get("LR", 0.0)), eps=1e-15)     # lr 0: the weights (and the sampler's round count) stay put
def step():
    out = m(inp)
    loss = loss_fn(out, gt, 0)["loss"]
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
core = m._ready_core()
N = int(os.environ.get("IT", 30))
core.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N): step()
e1.record()
torch.cuda.synchronize()
prof = core.profile_read()
core.profile(False)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("I2SDF_"))
print(json.dumps({"env": tag, "ms_per_step": round(e0.elapsed_time(e1) / N, 4),
                  "kernel_ms": {k: round(v["ms"] / N, 4) for k, v in prof.items()},
                  "launches": {k: v["launches"] / N for k, v in prof.items()}}))
