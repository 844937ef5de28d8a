"""cProfile of the host side of training steps (where does the Python / dispatch time go?)."""
import cProfile, pstats, io, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs
from i2sdf_b200.network import I2SDFNetwork, I2SDFLoss
from i2sdf_b200 import synthetic as orc   # (neutral input generator: perf tools do not touch oracle/)
import bench
conf = configs.model_conf("synthetic"); conf["use_normal"] = True
torch.manual_seed(0)
m = I2SDFNetwork(conf)
with torch.no_grad():
    m.density.beta.fill_(0.01)
m = m.cuda().train()
R = 1024
inp = {k: v.cuda() for k, v in orc.synthetic_rays(R, seed=1, train_layout=True).items()}
gt = {k: v.cuda() for k, v in bench.make_train_gt(R, 7).items()}
loss_fn = I2SDFLoss(**configs.LOSS_SYNTHETIC)
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
N = 30
t0 = time.perf_counter()
for _ in range(N): step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue time per step {1e3*(t1-t0)/N:.2f} ms ; incl. final drain {1e3*(t2-t0)/N:.2f} ms")
pr = cProfile.Profile(); pr.enable()
for _ in range(N): step()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(25); print(s.getvalue()[:6000])
