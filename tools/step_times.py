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
import time
host = {"forward": 0.0, "loss": 0.0, "backward": 0.0, "opt": 0.0}
def step():
    t0 = time.perf_counter()
    out = m(inp)
    t1 = time.perf_counter()
    loss = loss_fn(out, gt, 0)["loss"]
    opt.zero_grad(set_to_none=True)
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    opt.step()
    t4 = time.perf_counter()
    host["forward"] += t1 - t0; host["loss"] += t2 - t1; host["backward"] += t3 - t2; host["opt"] += t4 - t3
for _ in range(5): step()
torch.cuda.synchronize()
core = m._ready_core()
N = int(os.environ.get("IT", 30))
REP = int(os.environ.get("REP", 3))
PROF = os.environ.get("PROF", "1") == "1"       # PROF=0: no per-kind event brackets (they add ~40 event records per step)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("I2SDF_"))
for rep in range(REP):
    for k in host: host[k] = 0.0
    if PROF: core.profile(True)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    evs[0].record()
    for i in range(N):
        step()
        evs[i + 1].record()
    torch.cuda.synchronize()
    prof = core.profile_read() if PROF else {}
    if PROF: core.profile(False)
    per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(N))
    print(json.dumps({"env": tag, "ms_per_step": round(evs[0].elapsed_time(evs[N]) / N, 4), "median": round(per[N // 2], 4), "min": round(per[0], 4),
                      "host_ms": {k: round(1e3 * v / N, 3) for k, v in host.items()},
                      "kernel_ms": {k: round(v["ms"] / N, 4) for k, v in prof.items() if v["launches"]},
                      "launches": sum(v["launches"] for v in prof.values()) / N if PROF else None}))
