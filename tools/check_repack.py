"""Does the training loop re-pack the weights after every optimizer step?  (p._version must move under fused Adam.)"""
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
R = 256
inp = {k: v.cuda() for k, v in orc.synthetic_rays(R, seed=1, train_layout=True).items()}
gt = {k: v.cuda() for k, v in bench.make_train_gt(R, 7).items()}
loss_fn = I2SDFLoss(**configs.LOSS_SYNTHETIC)
opt = torch.optim.Adam(m.parameters(), lr=5e-4, eps=1e-15, fused=True)
n_pack = [0]
orig = m.pack_weights
def counted():
    n_pack[0] += 1
    return orig()
m.pack_weights = counted
p0 = m.implicit_network.lin3.weight_v
for it in range(4):
    v0 = p0._version
    out = m(inp)
    loss = loss_fn(out, gt, 0)["loss"]
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
    print(f"step {it}: loss {loss.item():.6f} packs so far {n_pack[0]} version {v0} -> {p0._version}")
# the packed weights must equal a fresh module with the same state dict
m.eval()
inp_e = {k: v.cuda() for k, v in orc.synthetic_rays(R, seed=1).items()}
a = m(inp_e)["rgb_values"].clone()
m2 = I2SDFNetwork(configs.model_conf("synthetic")); m2.load_state_dict(m.state_dict()); m2 = m2.cuda().eval()
b = m2(inp_e)["rgb_values"]
print("eval after training vs fresh module with same state:", float((a - b).abs().max()), "packs", n_pack[0])
