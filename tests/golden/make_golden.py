"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

For every case the script (1) builds the reference I2SDFNetwork from the reference yaml with the reference's
own constructor and RNG seed, (2) runs reference forward (and, for the training case, I2SDFLoss + backward),
(3) runs oracle/i2sdf_oracle.py on the same weights / rays / RNG tape, (4) asserts they agree, and (5) stores
inputs, tape, reference outputs and the oracle's per-round sampler trace.  The GPU box has no reference tree:
tests read only these files.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

from oracle import ref_shim  # noqa: E402
from oracle import i2sdf_oracle as orc  # noqa: E402
from i2sdf_b200 import configs  # noqa: E402

torch.set_num_threads(8)


def build_ref(conf_name, use_normal, beta, perturb):
    net, _ = ref_shim.load()
    conf = ref_shim.load_conf(conf_name + ".yml")
    conf.model.use_normal = use_normal
    torch.manual_seed(0)
    m = net.I2SDFNetwork(conf.model)
    g = torch.Generator().manual_seed(1234)
    with torch.no_grad():
        m.density.beta.fill_(beta)
        if perturb > 0:   # exercise the positional-encoding columns the geometric init zeroes
            imp = m.implicit_network
            for l in (0, *imp.skip_in):
                v = getattr(imp, f"lin{l}").weight_v
                cols = slice(3, None) if l == 0 else slice(v.shape[1] - 36, None)
                v[:, cols] += perturb * (2 ** 0.5 / v.shape[0] ** 0.5) * torch.randn(v[:, cols].shape, generator=g)
            # move the colour / light heads away from their small default init
            for mod in [m.rendering_network] + ([m.light_network] if m.use_light else []):
                for p in mod.parameters():
                    p.mul_(1.5)
    return net, conf, m


def params_of(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def hook_intermediates(m, store):
    gz = m.ray_sampler.get_z_vals
    go = m.implicit_network.get_outputs

    def get_z_vals(*a, **k):
        z, ze = gz(*a, **k)
        store["z_all"], store["z_eik"] = z.detach().clone(), ze.detach().clone()
        return z, ze

    def get_outputs(x, returns_grad=True):
        s, f, g = go(x, returns_grad)
        store["sdf"] = s.detach().clone()
        store["feat"] = f.detach().clone()
        store["feat_head"] = f.detach()[:FEAT_HEAD].clone()   # only a slice of feat is stored (size)
        store["grad"] = g.detach().clone() if g is not None else None
        return s, f, g

    m.ray_sampler.get_z_vals = get_z_vals
    m.implicit_network.get_outputs = get_outputs


FEAT_HEAD = 2 * 97      # points whose 256 features are stored in the fixture


def relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-12))


def np_dict(d, prefix=""):
    out = {}
    for k, v in d.items():
        if v is None:
            continue
        if isinstance(v, torch.Tensor):
            out[prefix + k] = v.detach().cpu().numpy()
        else:
            out[prefix + k] = np.asarray(v)
    return out


def save_weights(name, P):
    np.savez_compressed(os.path.join(HERE, f"weights_{name}.npz"), **np_dict(P))


def sampler_trace_arrays(trace):
    out = {"n_rounds": trace["n_rounds"], "n_final": trace["n_final"], "extra_idx": trace["extra_idx"],
           "beta_final": trace["beta_final"]}
    for i, r in enumerate(trace["rounds"]):
        for k in ("z", "sdf", "d_star", "beta", "cdf", "inds", "samples", "perm", "z_merged"):
            if k in r:
                out[f"round{i}_{k}"] = r[k]
        out[f"round{i}_upsample"] = int(r["upsample"])
    return out


def eval_case(case, conf_name, weights_name, beta, R, perturb, seed):
    net, conf, m = build_ref(conf_name, False, beta, perturb)
    m.eval()
    P = params_of(m)
    spec = orc.spec_from_model_conf(configs.model_conf(conf_name), use_normal=False)
    inp = orc.synthetic_rays(R, seed=seed)
    store = {}
    hook_intermediates(m, store)
    ref_out = m({k: v.clone() for k, v in inp.items()})
    ref_out = {k: v.detach() for k, v in ref_out.items()}
    trace = {}
    with torch.no_grad():
        o_out = orc.render(spec, P, inp, training=False, trace=trace)
    print(f"[{case}] rounds={trace['n_rounds']} n_final={trace['n_final']}")
    assert torch.equal(trace["z"], store["z_all"]), "oracle sampler z differs from the reference"
    for k in ref_out:
        e = relerr(o_out[k], ref_out[k])
        print(f"   {k:14s} oracle-vs-reference rel err {e:.2e}")
        assert e < 2e-5, (k, e)
    for k in ("sdf", "feat", "grad"):
        e = relerr(trace[k], store[k])
        print(f"   {k:14s} oracle-vs-reference rel err {e:.2e}")
        assert e < 2e-5, (k, e)
    arrays = {}
    arrays.update(np_dict(inp, "in_"))
    arrays.update(np_dict(ref_out, "ref_"))
    arrays.update(np_dict({k: v for k, v in store.items() if k != "feat"}, "ref_mid_"))
    arrays.update(np_dict(sampler_trace_arrays(trace), "trace_"))
    arrays["ref_mid_rgb"] = trace["rgb"].numpy()       # oracle rgb per sample (checked through rgb_values)
    arrays["meta_conf"] = np.asarray(conf_name)
    arrays["meta_weights"] = np.asarray(weights_name)
    arrays["meta_beta"] = np.asarray(beta, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, f"{case}.npz"), **arrays)
    return P


def train_case(case, conf_name, weights_name, beta, R, perturb, seed, n_cloud=24):
    net, conf, m = build_ref(conf_name, True, beta, perturb)
    m.train()
    P = params_of(m)
    spec = orc.spec_from_model_conf(configs.model_conf(conf_name), use_normal=True)
    inp = orc.synthetic_rays(R, seed=seed, train_layout=True)
    g = torch.Generator().manual_seed(99)
    inp["pointcloud"] = (torch.rand(n_cloud, 3, generator=g) - 0.5) * 1.2
    gt = {
        "rgb": torch.rand(R, 3, generator=g),
        "depth": torch.rand(R, generator=g) * 2 + 0.5,
        "depth_mask": torch.rand(R, generator=g) > 0.2,
        "normal": torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1),
        "normal_mask": torch.rand(R, generator=g) > 0.2,
    }
    if m.use_light:
        gt["light_mask"] = (torch.rand(R, generator=g) > 0.5).float()
    loss_conf = dict(configs.LOSS_SYNTHETIC_LIGHT_MASK if m.use_light else configs.LOSS_SYNTHETIC)
    loss_fn = net.I2SDFLoss(**loss_conf)
    step = 200000
    # Normal supervision only on rays that carry weight: for an empty ray normal_values = normalize(sum w n) divides by
    # |sum w n| ~ 1e-10, so its loss gradient is amplified rounding noise in the reference itself (real scenes are
    # closed rooms: every supervised pixel hits a surface).  Pre-pass with the same RNG state to find those rays.
    torch.manual_seed(4242)
    np.random.seed(7)
    pre = m({k: v.clone() for k, v in inp.items()})
    gt["normal_mask"] = gt["normal_mask"] & (pre["weight_sum"][:, 0].detach() > 0.5)
    del pre
    assert gt["normal_mask"].sum() >= 4
    # ---- reference run under a known RNG state
    store = {}
    hook_intermediates(m, store)
    torch.manual_seed(4242)
    np.random.seed(7)
    ref_out = m({k: v.clone() for k, v in inp.items()})
    ref_loss = loss_fn(ref_out, gt, step)["loss"]
    ref_loss.backward()
    ref_grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    # ---- replay the same RNG stream into a tape (order: ray_sampler.py:39,190,223,233; network/__init__.py:178,186,198)
    torch.manual_seed(4242)
    np.random.seed(7)
    tape = {"jitter": torch.rand(R, spec.n_samples_eval), "u_final": torch.rand(R, spec.n_samples)}
    o, d, _ = orc.flatten_rays(inp["uv"], inp["pose"], inp["intrinsics"])
    layers = orc.layer_params(P, "implicit_network", spec.n_sdf_layers)
    tr0 = {}
    with torch.no_grad():
        orc.sample_z(spec, lambda p: orc.sdf_mlp(spec, layers, p)[0][:, :1], o, d,
                     P["density.beta"].abs() + spec.beta_min, True,
                     dict(tape, extra_perm=torch.arange(spec.n_samples_extra)), tr0)
    tape["extra_perm"] = torch.randperm(tr0["n_final"])[:spec.n_samples_extra]
    tape["eik_idx"] = torch.randint(spec.n_samples + 2 + spec.n_samples_extra, (R,))
    tape["eik_uniform"] = torch.empty(R, 3).uniform_(-spec.bounding_sphere, spec.bounding_sphere)
    tape["nbr_uniform"] = torch.empty(R, 3).uniform_(-0.005, 0.005)
    tape["bubble_cam_idx"] = torch.tensor(np.random.randint(0, R))
    # ---- oracle run with autograd
    Pg = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    trace = {}
    o_out = orc.render(spec, Pg, inp, training=True, tape=tape, trace=trace)
    o_loss = orc.recon_loss(o_out, gt, **{k: v for k, v in loss_conf.items()
                                          if k in ("eikonal_weight", "smooth_weight", "depth_weight", "normal_weight",
                                                   "bubble_weight", "light_mask_weight")})
    o_loss.backward()
    print(f"[{case}] rounds={trace['n_rounds']} n_final={trace['n_final']} loss ref={ref_loss.item():.6f} oracle={o_loss.item():.6f}")
    assert torch.equal(trace["z"], store["z_all"]), "oracle sampler z differs from the reference (train)"
    for k in ref_out:
        e = relerr(o_out[k].detach(), ref_out[k].detach())
        print(f"   {k:14s} oracle-vs-reference rel err {e:.2e}")
        assert e < 5e-5, (k, e)
    assert abs(o_loss.item() - ref_loss.item()) < 1e-5 * abs(ref_loss.item()) + 1e-7
    worst = 0.0
    for k, gr in ref_grads.items():
        e = relerr(Pg[k].grad, gr)
        worst = max(worst, e)
        assert e < 2e-3, (k, e)
    print(f"   param grads    oracle-vs-reference worst rel err {worst:.2e}")
    arrays = {}
    arrays.update(np_dict(inp, "in_"))
    arrays.update(np_dict(gt, "gt_"))
    arrays.update(np_dict(tape, "tape_"))
    arrays.update(np_dict({k: v.detach() for k, v in ref_out.items()}, "ref_"))
    arrays.update(np_dict({k: v for k, v in store.items() if k != "feat"}, "ref_mid_"))
    arrays.update(np_dict(sampler_trace_arrays(trace), "trace_"))
    arrays["ref_loss"] = np.asarray(ref_loss.item(), dtype=np.float64)
    for k, gr in ref_grads.items():
        if gr.numel() <= 4096:
            arrays["refgrad_full_" + k] = gr.numpy()
        else:       # digest of the big matrices: norm, sum, strided subsample
            flat = gr.flatten()
            arrays["refgrad_norm_" + k] = np.asarray(flat.norm().item())
            arrays["refgrad_sum_" + k] = np.asarray(flat.double().sum().item())
            arrays["refgrad_sub_" + k] = flat[::97].numpy()
    arrays["meta_conf"] = np.asarray(conf_name)
    arrays["meta_weights"] = np.asarray(weights_name)
    arrays["meta_beta"] = np.asarray(beta, dtype=np.float32)
    arrays["meta_step"] = np.asarray(step)
    np.savez_compressed(os.path.join(HERE, f"{case}.npz"), **arrays)
    return P


def uniform_case(case, conf_name, weights_name, beta, R, perturb, seed, n_uniform=64):
    """BASELINE.json configs[0] (C1): the reference forward with its OWN `UniformSampler(scene_bounding_sphere, near, 64)`
    (model/network/ray_sampler.py:15-43) in place of the error-bounded sampler, eval mode: z = linspace(0, 6, 64), 63
    composited points per ray.  I2SDFNetwork.forward expects (z_vals, z_samples_eik) from its sampler (network/__init__.py:96),
    UniformSampler returns z_vals only, so the stand-in pairs it with z[:, :1] (unused in eval mode)."""
    import importlib
    net, conf, m = build_ref(conf_name, False, beta, perturb)
    m.eval()
    P = params_of(m)
    rs = importlib.import_module("model.network.ray_sampler")
    uni = rs.UniformSampler(conf.model.scene_bounding_sphere, conf.model.ray_sampler.near, n_uniform)

    class _Pair:
        def get_z_vals(self, ray_dirs, cam_loc, model):
            z = uni.get_z_vals(ray_dirs, cam_loc, model)
            return z, z[:, :1]

    m.ray_sampler = _Pair()
    spec = orc.spec_from_model_conf(configs.model_conf(conf_name), use_normal=False)
    inp = orc.synthetic_rays(R, seed=seed)
    store = {}
    hook_intermediates(m, store)
    ref_out = {k: v.detach() for k, v in m({k: v.clone() for k, v in inp.items()}).items()}
    z = store["z_all"]
    assert z.shape == (R, n_uniform) and torch.equal(z[0], torch.linspace(0., 1., n_uniform) * spec.far)
    with torch.no_grad():
        o_out = orc.render(spec, P, inp, training=False, z_override=z)
    for k in ref_out:
        e = relerr(o_out[k], ref_out[k])
        print(f"   [{case}] {k:14s} oracle-vs-reference rel err {e:.2e}")
        assert e < 2e-5, (k, e)
    arrays = {}
    arrays.update(np_dict(inp, "in_"))
    arrays.update(np_dict(ref_out, "ref_"))
    arrays.update(np_dict({"z_all": z, "sdf": store["sdf"], "grad": store["grad"]}, "ref_mid_"))
    arrays["meta_conf"] = np.asarray(conf_name)
    arrays["meta_weights"] = np.asarray(weights_name)
    arrays["meta_beta"] = np.asarray(beta, dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, f"{case}.npz"), **arrays)
    return P


LOSS_ALL_KW = dict(eikonal_weight=0.1, smooth_weight=0.01, mask_weight=0.2, depth_weight=0.1, normal_weight=0.05, angular_weight=0.05,
                   bubble_weight=0.5, light_mask_weight=0.5, smooth_iter=10)


def loss_case(case="loss_all_terms", R=57, seed=11):
    """The reference's I2SDFLoss (model/network/__init__.py:289-406) with EVERY term switched on (the training fixtures only
    exercise the terms of the shipped yamls) on random model outputs: all ten returned values at two steps (before / after
    smooth_iter) and d loss / d output from the reference's autograd.  Includes the edges the CUDA kernel special-cases:
    |grad_theta| = 0, weight_sum / light_mask outside the BCE clip range."""
    net, _ = ref_shim.load()
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.rand(*s, generator=g)          # noqa: E731
    out = {"rgb_values": rnd(R, 3), "depth_values": rnd(R) * 3, "weight_sum": rnd(R, 1) * 1.2 - 0.1,
           "normal_values": torch.nn.functional.normalize(rnd(R, 3) - 0.5, dim=1), "grad_theta": (rnd(2 * R, 3) - 0.5) * 3,
           "diff_norm": rnd(R), "surface_sdf": rnd(23, 1) - 0.5, "light_mask": rnd(R, 1) * 1.2 - 0.1}
    out["grad_theta"][3] = 0.0
    gt = {"rgb": rnd(R, 1, 3), "depth": rnd(R, 1) * 3, "depth_mask": rnd(R, 1) > 0.3,
          "normal": torch.nn.functional.normalize(rnd(R, 3) - 0.5, dim=1), "normal_mask": rnd(R) > 0.5,
          "mask": (rnd(R, 1) > 0.5).float(), "light_mask": (rnd(R, 1) > 0.8).float()}
    fn = net.I2SDFLoss(**LOSS_ALL_KW)
    arrays = {}
    arrays.update(np_dict(out, "out_"))
    arrays.update(np_dict(gt, "gt_"))
    for step in (5, 100):
        leaves = {k: v.clone().requires_grad_(True) for k, v in out.items()}
        res = fn(leaves, gt, step)
        res["loss"].backward()
        arrays.update(np_dict({k: v.detach() for k, v in res.items()}, f"ref{step}_"))
        arrays.update(np_dict({k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}, f"refgrad{step}_"))
        o_loss = orc.recon_loss(out, gt, smooth_active=step > LOSS_ALL_KW["smooth_iter"],
                                **{k: v for k, v in LOSS_ALL_KW.items() if k != "smooth_iter"})
        print(f"   [{case}] step {step}: reference loss {res['loss'].item():.7f} oracle {o_loss.item():.7f}")
        assert abs(o_loss.item() - res["loss"].item()) < 1e-6 * abs(res["loss"].item())
    np.savez_compressed(os.path.join(HERE, f"{case}.npz"), **arrays)


if __name__ == "__main__":
    if "--only-loss" in sys.argv:             # adds the all-terms loss fixture without rewriting the others
        loss_case()
        sys.exit(0)
    if "--only-uniform" in sys.argv:          # adds the C1 fixture without rewriting the others (same weights: asserted)
        Pu = uniform_case("eval_uniform_c1", "synthetic", "synthetic", beta=0.05, R=48, perturb=0.06, seed=6)
        w = np.load(os.path.join(HERE, "weights_synthetic.npz"))
        assert all(np.array_equal(w[k], Pu[k].numpy()) for k in w.files)
        sys.exit(0)
    P = eval_case("eval_synthetic_sharp", "synthetic", "synthetic", beta=0.01, R=48, perturb=0.06, seed=1)
    Pw = {k: v for k, v in P.items() if k != "density.beta"}
    save_weights("synthetic", Pw)
    P2 = eval_case("eval_synthetic_soft", "synthetic", "synthetic", beta=0.1, R=48, perturb=0.06, seed=2)
    assert all(torch.equal(P2[k], Pw[k]) for k in Pw)
    P3 = train_case("train_synthetic", "synthetic", "synthetic", beta=0.02, R=32, perturb=0.06, seed=3)
    assert all(torch.equal(P3[k], Pw[k]) for k in Pw)
    P4 = eval_case("eval_light_sharp", "synthetic_light_mask", "light", beta=0.01, R=48, perturb=0.06, seed=4)
    Pl = {k: v for k, v in P4.items() if k != "density.beta"}
    save_weights("light", Pl)
    P5 = train_case("train_light", "synthetic_light_mask", "light", beta=0.05, R=32, perturb=0.06, seed=5)
    assert all(torch.equal(P5[k], Pl[k]) for k in Pl)
    Pu = uniform_case("eval_uniform_c1", "synthetic", "synthetic", beta=0.05, R=48, perturb=0.06, seed=6)
    assert all(torch.equal(Pu[k], Pw[k]) for k in Pw)
    loss_case()
    print("golden fixtures written to", HERE)
