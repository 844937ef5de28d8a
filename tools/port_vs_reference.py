"""How fast is the oracle PORT that bench.py's CPU arms time, relative to the UNMODIFIED reference?  (Build container only: needs
/root/reference; the GPU box has no reference tree, which is why the arms time the port.)  Same weights, rays, thread count:
eval render and one training step (forward + I2SDFLoss + backward incl. the double backward) of 128 rays.
Result recorded in bench.py: PORT_VS_REFERENCE and DESIGN.md §7.

    python tools/port_vs_reference.py
"""
import sys, time, torch, warnings
warnings.filterwarnings("ignore")
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_shim, i2sdf_oracle as orc
from i2sdf_b200 import configs
import contextlib, io
torch.set_num_threads(8)
net, _ = ref_shim.load()
conf = ref_shim.load_conf("synthetic.yml")
conf.model.use_normal = True
torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = net.I2SDFNetwork(conf.model)
with torch.no_grad():
    m.density.beta.fill_(0.01)
R = 128
# ---- eval render
m.eval()
inp = orc.synthetic_rays(R, seed=1)
P = {k: v.detach().clone() for k, v in m.state_dict().items()}
spec = orc.spec_from_model_conf(configs.model_conf("synthetic"), use_normal=False)
def t(fn, n=3):
    fn(); ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)
tr = t(lambda: m({k: v.clone() for k, v in inp.items()}))
with torch.no_grad():
    to = t(lambda: orc.render(spec, P, inp, training=False))
print(f"eval render {R} rays: reference {tr*1e3:.0f} ms = {R*97/tr:.0f} rs/s | oracle port {to*1e3:.0f} ms = {R*97/to:.0f} rs/s")
# ---- training step (forward + loss + backward)
m.train()
inp_t = orc.synthetic_rays(R, seed=1, train_layout=True)
g = torch.Generator().manual_seed(7)
gt = {"rgb": torch.rand(R, 3, generator=g), "depth": torch.rand(R, generator=g) * 2 + 0.5, "depth_mask": torch.ones(R, dtype=torch.bool),
      "normal": torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1), "normal_mask": torch.ones(R, dtype=torch.bool)}
loss_fn = net.I2SDFLoss(**configs.LOSS_SYNTHETIC)
def ref_step():
    m.zero_grad(set_to_none=True)
    out = m({k: v.clone() for k, v in inp_t.items()})
    loss_fn(out, gt, 0)["loss"].backward()
spec_t = orc.spec_from_model_conf(configs.model_conf("synthetic"), use_normal=True)
keys = ("eikonal_weight", "smooth_weight", "depth_weight", "normal_weight", "bubble_weight", "light_mask_weight")
lw = {k: v for k, v in configs.LOSS_SYNTHETIC.items() if k in keys}
def orc_step():
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    tape = {"jitter": torch.rand(R, 128), "u_final": torch.rand(R, 64), "extra_perm": lambda n: torch.randperm(n)[:32],
            "eik_idx": torch.randint(98, (R,)), "eik_uniform": torch.empty(R, 3).uniform_(-3, 3), "nbr_uniform": torch.empty(R, 3).uniform_(-0.005, 0.005)}
    out = orc.render(spec_t, Pg, inp_t, training=True, tape=tape)
    orc.recon_loss(out, gt, smooth_active=False, **lw).backward()
tr = t(ref_step, 2); to = t(orc_step, 2)
print(f"train step {R} rays: reference {tr*1e3:.0f} ms = {R*97/tr:.0f} rs/s | oracle port {to*1e3:.0f} ms = {R*97/to:.0f} rs/s")
