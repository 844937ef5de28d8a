"""Error of the sdf-only kernels (tensor-core chain, fp32 SIMT cross-check) against a float64 evaluation of the same network.

Prints max / rms absolute error of sdf over random points and over points on rays near the surface (where the sampler's
decisions are taken), plus the fp32 CPU oracle's own error against float64 as the yardstick."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_util import Case  # noqa: E402
from oracle import i2sdf_oracle as orc  # noqa: E402
from parity_report import make_model  # noqa: E402


def sdf64(m, x):
    Ws, bs = m.effective_weights()
    n_sdf = m.implicit_network.num_layers - 1
    W = [w.detach().double() for w in Ws[:n_sdf]]
    b = [t.detach().double() for t in bs[:n_sdf]]
    x = x.double()
    mx = m.implicit_network.multires
    e = torch.cat([x] + [f(x * 2.0 ** k) for k in range(mx) for f in (torch.sin, torch.cos)], -1)
    h = e
    skip = m.implicit_network.skip_in[0] if m.implicit_network.skip_in else -1
    for l in range(n_sdf):
        if l == skip:
            h = torch.cat([h, e], -1) / 2 ** 0.5
        a = h @ W[l].T + b[l]
        if l < n_sdf - 1:
            h = torch.nn.functional.softplus(a, beta=100)
    return a[:, 0]


def main():
    kappas = [float(k) for k in os.environ.get("KAPPAS", "").split(",") if k] or [None]
    g = torch.Generator().manual_seed(3)
    inp = orc.synthetic_rays(1024, seed=1)
    o, d, _ = orc.flatten_rays(inp["uv"], inp["pose"], inp["intrinsics"])
    t = torch.linspace(0.0, 6.0, 128)
    sets = {"uniform cube [-1.5,1.5]^3": (torch.rand(131072, 3, generator=g) - 0.5) * 3.0,
            "ray points (1024 rays x 128)": (o[:, None, :] + t[None, :, None] * d[:, None, :]).reshape(-1, 3)}
    for case_name, noise in (("eval_synthetic_sharp", 0.0), ("eval_light_sharp", 0.0), ("eval_synthetic_sharp", 0.3)):
        c = Case(case_name)
        if noise:       # a stand-in for trained weights: every weight direction perturbed by 30 %, gains by 10 %
            gn = torch.Generator().manual_seed(7)
            for k in list(c.params):
                if k.endswith("weight_v"):
                    c.params[k] = c.params[k] * (1.0 + noise * torch.randn(c.params[k].shape, generator=gn))
                elif k.endswith("weight_g"):
                    c.params[k] = c.params[k] * (1.0 + 0.1 * torch.randn(c.params[k].shape, generator=gn))
        print(f"==== {case_name}  weight noise {noise}")
        ms = make_model(c, {"I2SDF_SIMT": "1"})
        layers = orc.layer_params(c.params, "implicit_network", c.spec.n_sdf_layers)
        for name, pts in sets.items():
            ref = sdf64(ms, pts.cuda())
            near = ref.abs() < 0.05
            print(f"  {name}: max|sdf| {float(ref.abs().max()):.3f}, near surface (|sdf| < 0.05): {int(near.sum())} points")

            def row(k, v):
                e = v - ref
                print(f"     {k:34s} all: max {float(e.abs().max()):.3e} rms {float(e.pow(2).mean().sqrt()):.3e} mean {float(e.mean()):+.3e} | near surface: "
                      f"max {float(e[near].abs().max()):.3e} rms {float(e[near].pow(2).mean().sqrt()):.3e} mean {float(e[near].mean()):+.3e}")
            row("fp32 SIMT kernel", ms._core_obj.sdf_forward(pts.cuda())[0].double())
            with torch.no_grad():
                row("fp32 CPU oracle", orc.sdf_mlp(c.spec, layers, pts)[0][:, 0].double().cuda())
            for kp in kappas:
                mt = make_model(c, {} if kp is None else {"I2SDF_KAPPA": repr(kp)})
                row(f"tensor-core chain kappa={'default' if kp is None else kp}", mt._core_obj.sdf_forward(pts.cuda())[0].double())


if __name__ == "__main__":
    main()
