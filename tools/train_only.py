"""Runs a few training steps on a 1024-ray batch (target for ncu captures of the backward kernels)."""
import os, sys, torch
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
R = int(os.environ.get("R", 1024))
inp = {k: v.cuda() for k, v in orc.synthetic_rays(R, seed=1, train_layout=True).items()}
gt = {k: v.cuda() for k, v in bench.make_train_gt(R, 7).items()}
loss_fn = I2SDFLoss(**configs.LOSS_SYNTHETIC)
for _ in range(int(os.environ.get("IT", 2))):
    out = m(inp)
    loss = loss_fn(out, gt, 0)["loss"]
    m.zero_grad(set_to_none=True)
    loss.backward()
torch.cuda.synchronize()
print("done")
