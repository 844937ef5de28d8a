"""Runs the tensor-core SDF kernel a few times on a 1024x128-point batch (target for ncu captures)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs
from i2sdf_b200.network import I2SDFNetwork
torch.manual_seed(0)
m = I2SDFNetwork(configs.model_conf("synthetic")).cuda().eval()
core = m._ready_core()
pts = (torch.rand(int(os.environ.get("M", 131072)), 3, device="cuda") - 0.5) * 3
for _ in range(int(os.environ.get("IT", 4))):
    core.sdf_forward(pts)
torch.cuda.synchronize()
print("done", core.uses_tensor_cores)
