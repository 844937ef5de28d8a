"""Runs the tensor-core main-pass kernel a few times on a 1024-ray batch (target for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs
from i2sdf_b200.network import I2SDFNetwork
from i2sdf_b200 import synthetic as orc   # (neutral input generator: perf tools do not touch oracle/)
torch.manual_seed(0)
m = I2SDFNetwork(configs.model_conf("synthetic"))
with torch.no_grad():
    m.density.beta.fill_(0.01)
m = m.cuda().eval()
core = m._ready_core()
inp = {k: v.cuda() for k, v in orc.synthetic_rays(int(os.environ.get("R", 1024)), seed=1).items()}
o, d, dn = core.rays(inp["uv"], inp["pose"], inp["intrinsics"])
z, _ = core.sample(o, d, m.density.beta.detach())
for _ in range(4):
    core.render(o, d, dn, z, m.density.beta.detach())
torch.cuda.synchronize()
print("done", core.uses_tensor_cores_main)
