"""End-to-end parity of the CUDA path against the CPU oracle / the reference fixtures, per precision variant.

Writes ONE JSON document (default profiles/parity_r02.json) with, for every variant (tensor-core path as shipped, the fp32 SIMT
cross-check path, ...) and every case (the three eval fixtures made from the unmodified reference + a 1024-ray W-sharp batch
rendered by the oracle on this box's CPU):
  * per-output relative error  max|a-b| / max|b|  (the north_star metric) and the PSNR of the render,
  * sampler: round count, share of z's within 1e-3 / 1e-5 of the reference, rays with any z off by > 1e-3,
  * per-round decisions on IDENTICAL float inputs (fixtures only): mismatching searchsorted indices, mismatching betas.
The oracle is the checker here (tools/ is test infrastructure, like tests/).

Usage:  python tools/parity_report.py [--out profiles/parity_r02.json] [--rays 1024]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from golden_util import EVAL_CASES, Case, relerr  # noqa: E402
from oracle import i2sdf_oracle as orc  # noqa: E402

VARIANTS = [("tensor_core", {}), ("fp32_simt", {"I2SDF_SIMT": "1"})]


def make_model(case, env):
    from i2sdf_b200.network import I2SDFNetwork
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        conf = dict(case.model_conf)
        conf["use_normal"] = False
        m = I2SDFNetwork(conf)
        m.load_state_dict(dict(case.params), strict=True)
        m = m.cuda().eval()
        m._ready_core()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return m


def compare_outputs(out, ref):
    res = {}
    hit = ref["weight_sum"][:, 0] > 1e-2
    for k, v in ref.items():
        a = out[k].detach().cpu()
        if k == "normal_map":
            a, v = a[hit], v[hit]
        res[k] = relerr(a, v)
    mse = ((out["rgb_values"].cpu() - ref["rgb_values"]) ** 2).mean()
    res["psnr_db"] = float(-10.0 * torch.log10(mse.clamp(min=1e-20)))
    # how many rays carry the error: rays whose rgb differs by more than 1e-4 of the tensor's max
    d = (out["rgb_values"].cpu() - ref["rgb_values"]).abs().max(-1)[0] / ref["rgb_values"].abs().max()
    res["rays_rgb_over_1e-4"] = int((d > 1e-4).sum())
    res["rays"] = int(d.numel())
    return res


def sampler_stats(core, model, o, d, z_ref, n_rounds_ref, sdf_min_ref):
    z, _, info = core.sample(o.cuda(), d.cuda(), model.density.beta.detach(), None, want_info=True)
    zc = z.cpu()
    hit = sdf_min_ref < 0
    dz = (zc - z_ref).abs()
    return {"rounds": int(info[0]), "rounds_ref": int(n_rounds_ref), "n_final": int(info[1]),
            "z_within_1e-3": float((dz < 1e-3).float().mean()), "z_within_1e-5": float((dz < 1e-5).float().mean()),
            "z_within_1e-3_hit": float((dz < 1e-3)[hit].float().mean()) if hit.any() else None,
            "rays_any_z_over_1e-3": int((dz.max(-1)[0] > 1e-3).sum()), "rays_any_z_over_1e-5": int((dz.max(-1)[0] > 1e-5).sum()),
            "z_max_abs_err": float(dz.max()), "rays": int(zc.shape[0]), "rays_hit": int(hit.sum())}


def round_decisions(core, model, c):
    """Each sampler round fed with the reference's own (z, sdf, beta_in): integer decisions on identical float inputs."""
    nr = int(c.trace["n_rounds"])
    beta_param = model.density.beta.detach()
    z0 = c.trace["round0_z"]
    dists = z0[:, 1:] - z0[:, :-1]
    beta_in = torch.sqrt((1.0 / (4.0 * torch.log(torch.tensor(c.spec.eps + 1.0)))) * (dists ** 2.0).sum(-1))
    hit = (c.mid["sdf"].reshape(c.mid["z_all"].shape[0], -1).min(-1)[0] < 0)
    st = dict(bad_inds=0, n_inds=0, bad_beta=0, n_beta=0, bad_inds_empty_rays=0, n_inds_empty_rays=0, max_cdf_err=0.0, max_sample_err_same_bin=0.0,
              max_sample_err_any=0.0)
    for i in range(nr):
        z, sdf = c.trace[f"round{i}_z"], c.trace[f"round{i}_sdf"]
        up = bool(int(c.trace[f"round{i}_upsample"]))
        out = core.sampler_round_debug(z.cuda(), sdf.cuda(), beta_param, beta_in.cuda(), up)
        ref_beta = c.trace[f"round{i}_beta"]
        beta = out["beta"].cpu()
        ok = (beta - ref_beta).abs() <= 1e-5 * ref_beta
        st["bad_beta"] += int((~ok).sum())
        st["n_beta"] += beta.numel()
        same = out["inds"].cpu().long() == c.trace[f"round{i}_inds"]
        good, empty = ok & hit, ok & ~hit
        st["bad_inds"] += int((~same[good]).sum())
        st["n_inds"] += int(same[good].numel())
        st["bad_inds_empty_rays"] += int((~same[empty]).sum())
        st["n_inds_empty_rays"] += int(same[empty].numel())
        cdf_err = (out["cdf"].cpu() - c.trace[f"round{i}_cdf"]).abs().max(-1)[0]
        se = (out["samples"].cpu() - c.trace[f"round{i}_samples"]).abs()
        if good.any():
            st["max_cdf_err"] = max(st["max_cdf_err"], float(cdf_err[good].max()))
            st["max_sample_err_same_bin"] = max(st["max_sample_err_same_bin"], float(se[good][same[good]].max()))
            st["max_sample_err_any"] = max(st["max_sample_err_any"], float(se[good].max()))
        beta_in = ref_beta
    return st


class _Libm64:
    """torch.exp / torch.expm1 of fp32 tensors evaluated in float64 and rounded once: another libm of IEEE quality (<= 0.5 ulp,
    where torch's Sleef kernels guarantee 1 ulp) - what a bit-different but equally valid exp does to the reference's own path."""

    def __enter__(self):
        self.exp, self.expm1 = torch.exp, torch.expm1
        torch.exp = lambda x: self.exp(x.double()).float() if x.dtype == torch.float32 else self.exp(x)
        torch.expm1 = lambda x: self.expm1(x.double()).float() if x.dtype == torch.float32 else self.expm1(x)

    def __exit__(self, *a):
        torch.exp, torch.expm1 = self.exp, self.expm1


class _Sdf64:
    """The oracle's MLPs evaluated in float64 and rounded to fp32: the most accurate sdf any fp32 implementation could return."""

    def __enter__(self):
        self.orig = orc.sdf_mlp

        def wrapped(spec, layers, x, want_grad=False):
            out, g = self.orig(spec, [(W.double(), b.double()) for W, b in layers], x.double(), want_grad)
            return out.float(), (None if g is None else g.float())
        orc.sdf_mlp = wrapped

    def __exit__(self, *a):
        orc.sdf_mlp = self.orig


def self_sensitivity(spec, params, inp, ref, trace_ref):
    """The reference path against ITSELF under perturbations at the last-bit level: the floor no implementation can beat."""
    import contextlib
    res = {}
    for name, ctxs in (("libm_exp_correctly_rounded", [_Libm64]), ("mlp_in_float64", [_Sdf64]), ("both", [_Libm64, _Sdf64])):
        tr = {}
        with contextlib.ExitStack() as st:
            for c in ctxs:
                st.enter_context(c())
            with torch.no_grad():
                out = orc.render(spec, params, inp, training=False, trace=tr)
        e = dict(outputs=compare_outputs({k: out[k] for k in ref}, ref))
        dz = (tr["z"] - trace_ref["z"]).abs()
        e["sampler"] = {"rounds": int(tr["n_rounds"]), "rounds_ref": int(trace_ref["n_rounds"]), "z_within_1e-3": float((dz < 1e-3).float().mean()),
                        "z_within_1e-5": float((dz < 1e-5).float().mean()), "rays_any_z_over_1e-3": int((dz.max(-1)[0] > 1e-3).sum()),
                        "rays_any_z_over_1e-5": int((dz.max(-1)[0] > 1e-5).sum()), "z_max_abs_err": float(dz.max()), "rays": int(dz.shape[0])}
        res[name] = e
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "parity_r02.json"))
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--variants", default=",".join(v[0] for v in VARIANTS))
    args = ap.parse_args()
    want = args.variants.split(",")
    report = dict(metric="max|a-b|/max|b| per tensor (fp32); normal_map on rays with weight_sum > 1e-2", gpu=torch.cuda.get_device_name(0),
                  torch=torch.__version__, variants={})
    cases = {n: Case(n) for n in EVAL_CASES}
    # the BASELINE batch: 1024 rays, W-sharp weights (all 5 sampler rounds), oracle render on this box's CPU
    big = cases["eval_synthetic_sharp"]
    inp_big = orc.synthetic_rays(args.rays, seed=1)
    t0 = time.time()
    trace_big = {}
    with torch.no_grad():
        ref_big = orc.render(big.spec, big.params, inp_big, training=False, trace=trace_big)
    report["oracle_seconds_big"] = time.time() - t0
    o_b, d_b, _ = orc.flatten_rays(inp_big["uv"], inp_big["pose"], inp_big["intrinsics"])
    keys = ("rgb_values", "depth_values", "weight_sum", "normal_map")
    report["reference_self_sensitivity"] = {
        "what": "the CPU oracle (= the reference's arithmetic) compared with itself when exp/expm1 come from another IEEE-quality libm and / or the "
                "MLPs are evaluated in float64: last-bit perturbations of the same algorithm on the same rays and weights",
        f"w_sharp_{args.rays}_rays": self_sensitivity(big.spec, big.params, inp_big, {k: ref_big[k] for k in keys}, trace_big)}
    print("reference self-sensitivity", json.dumps(report["reference_self_sensitivity"], indent=1))
    for vname, env in VARIANTS:
        if vname not in want:
            continue
        rv = {}
        for name, c in cases.items():
            m = make_model(c, env)
            core = m._ready_core()
            out = m({k: v.cuda() for k, v in c.inputs.items()})
            e = dict(outputs=compare_outputs(out, c.ref))
            o, d, _ = orc.flatten_rays(c.inputs["uv"], c.inputs["pose"], c.inputs["intrinsics"])
            sdf_min = c.mid["sdf"].reshape(c.mid["z_all"].shape[0], -1).min(-1)[0]
            e["sampler"] = sampler_stats(core, m, o, d, c.mid["z_all"], c.trace["n_rounds"], sdf_min)
            e["round_decisions_identical_inputs"] = round_decisions(core, m, c)
            e["tensor_cores"] = dict(sampler=core.uses_tensor_cores, main=core.uses_tensor_cores_main)
            rv[name] = e
        m = make_model(big, env)
        core = m._ready_core()
        out = m({k: v.cuda() for k, v in inp_big.items()})
        e = dict(outputs=compare_outputs(out, {k: ref_big[k] for k in ("rgb_values", "depth_values", "weight_sum", "normal_map")}))
        zr = trace_big.get("z")
        if zr is not None:
            sdf_min = trace_big["sdf"].reshape(zr.shape[0], -1).min(-1)[0]
            e["sampler"] = sampler_stats(core, m, o_b, d_b, zr, trace_big.get("n_rounds", -1), sdf_min)
        rv[f"w_sharp_{args.rays}_rays_vs_oracle"] = e
        report["variants"][vname] = rv
        print(vname, json.dumps(rv, indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(report, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
