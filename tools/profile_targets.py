"""ncu target: ONE training step (config via CONFIG, default synthetic_light_mask so the light-head kernels appear) followed by
ONE eval render of the same 1024 rays.  Used for the `ncu --set full` captures and the launch lists under profiles/."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs
from i2sdf_b200.network import I2SDFNetwork, I2SDFLoss
from i2sdf_b200 import synthetic as orc   # (neutral input generator: perf tools do not touch oracle/)
import bench
name = os.environ.get("CONFIG", "synthetic_light_mask")
light = name != "synthetic"
conf = configs.model_conf(name); conf["use_normal"] = True
torch.manual_seed(0)
m = I2SDFNetwork(conf)
with torch.no_grad():
    m.density.beta.fill_(0.01)
m = m.cuda().train()
R = int(os.environ.get("R", 1024))
inp = {k: v.cuda() for k, v in orc.synthetic_rays(R, seed=1, train_layout=True).items()}
gt = {k: v.cuda() for k, v in bench.make_train_gt(R, 7, light).items()}
loss_fn = I2SDFLoss(**bench.loss_weights(name)[0])
from i2sdf_b200.optim import Adam
opt = Adam(m.parameters(), lr=5e-4, eps=1e-15)
for _ in range(int(os.environ.get("IT", 1))):
    out = m(inp)
    loss = loss_fn(out, gt, 0)["loss"]
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
if os.environ.get("EVAL", "1") == "1":
    m.eval()
    ev = {k: v.cuda() for k, v in orc.synthetic_rays(R, seed=1).items()}
    m(ev)
    torch.cuda.synchronize()
print("done")
