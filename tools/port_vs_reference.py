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
def pair(fa, fb, n=7):
    """Interleaved repetitions (the container's vCPUs are shared: back-to-back blocks drift by 20-30 %) -> (min, median) of each."""
    fa(); fb()
    ta, tb = [], []
    for _ in range(n):
        t0 = time.perf_counter(); fa(); ta.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); fb(); tb.append(time.perf_counter() - t0)
    ta.sort(); tb.sort()
    return (ta[0], ta[n // 2]), (tb[0], tb[n // 2])
def orc_eval():
    with torch.no_grad():
        orc.render(spec, P, inp, training=False)
(tr, trm), (to, tom) = pair(lambda: m({k: v.clone() for k, v in inp.items()}), orc_eval)
print(f"eval render {R} rays: reference min {tr*1e3:.0f} / median {trm*1e3:.0f} ms | oracle port min {to*1e3:.0f} / median {tom*1e3:.0f} ms | port speed = {tr/to:.2f}x (min), {trm/tom:.2f}x (median) of the reference")
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
(tr, trm), (to, tom) = pair(ref_step, orc_step)
print(f"train step {R} rays: reference min {tr*1e3:.0f} / median {trm*1e3:.0f} ms | oracle port min {to*1e3:.0f} / median {tom*1e3:.0f} ms | port speed = {tr/to:.2f}x (min), {trm/tom:.2f}x (median) of the reference")
