"""CPU oracle for the I2-SDF per-ray hot path.  TEST INFRASTRUCTURE — NOT THE PRODUCT.

A functional, single-file fp32 restatement (torch CPU tensors, no nn.Module, no autograd.grad for the
spatial gradient) of the reference algorithm, each function citing the reference file:line it follows
(paths relative to jingsenzhu/i2-sdf).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` leg may import this module; the product path (i2sdf_b200/) never does.

Pinning: tests/golden/*.npz were produced by tests/golden/make_golden.py, which runs the UNMODIFIED
reference (imported read-only through oracle/ref_shim.py) and this oracle on identical weights/rays and
records both; tests/test_oracle_golden.py re-checks the oracle against those reference outputs on every run, and
tests/test_oracle_vs_reference_live.py runs the reference LIVE beside the oracle wherever the reference tree exists
(fresh seeds, ragged ray counts, rotated cameras with skew, stage-wise rays / embeddings / density / error bound).
The reference itself ships no tests or golden vectors, so "reference outputs generated
here" is the strongest pin available.

Design notes
* torch CPU ops are used on purpose: the reference's arithmetic *is* ATen CPU kernels (cumsum accumulates
  fp32 inputs in double, searchsorted(right=True), unstable sort, nan_to_num), so the same primitives make the
  restatement bit-comparable with the reference on the sampler's discrete decisions.
* grad_x sdf is an explicit reverse sweep built from differentiable ops, so autograd through it yields the
  second-order parameter gradients the eikonal / normal losses need (reference: autograd.grad(create_graph)).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------
# Network / sampler description  (reference: config/synthetic.yml:30-74, config/synthetic_light_mask.yml)
# --------------------------------------------------------------------------------------------------
@dataclass
class PathSpec:
    sdf_dims: List[int]            # e.g. [39,256,...,256,257]   (mlp.py:29-36)
    skip_in: Tuple[int, ...]       # e.g. (4,)
    multires_x: int                # 6  -> 39-wide embedding
    color_dims: List[int]          # e.g. [283,256,256,256,256,3] (mlp.py:176-187)
    multires_d: int                # 4  -> 27-wide embedding
    light_dims: Optional[List[int]] = None   # [256,128,1] or None (network/__init__.py:29-32)
    feature_size: int = 256
    bounding_sphere: float = 3.0
    near: float = 0.0
    n_samples: int = 64
    n_samples_eval: int = 128
    n_samples_extra: int = 32
    eps: float = 0.1
    beta_iters: int = 10
    max_total_iters: int = 5
    add_tiny: float = 1.0e-6
    beta_min: float = 1.0e-4
    use_normal: bool = False
    detach_light_feature: bool = True

    @property
    def far(self) -> float:        # ray_sampler.py:17  (far = 2 * scene_bounding_sphere)
        return 2.0 * self.bounding_sphere

    @property
    def n_sdf_layers(self) -> int:
        return len(self.sdf_dims) - 1

    @property
    def n_color_layers(self) -> int:
        return len(self.color_dims) - 1

    def sdf_layer_shape(self, l: int) -> Tuple[int, int]:
        """(out, in) of lin{l}; the layer feeding a skip connection is narrower (mlp.py:44-50)."""
        out = self.sdf_dims[l + 1]
        if (l + 1) in self.skip_in:
            out -= self.sdf_dims[0]
        return out, self.sdf_dims[l]


def spec_from_model_conf(conf: dict, use_normal: Optional[bool] = None) -> PathSpec:
    """conf = the `model:` node of a reference yaml, as a plain dict (network/__init__.py:20-47)."""
    imp, ren, smp = conf["implicit_network"], conf["rendering_network"], conf["ray_sampler"]
    fvs = conf["feature_vector_size"]
    ex = 3 + 6 * imp["multires"]
    ed = 3 + 6 * ren["multires"]
    sdf_dims = [ex] + list(imp["dims"]) + [imp["d_out"] + fvs]
    color_dims = [ren["d_in"] + fvs + (ed - 3)] + list(ren["dims"]) + [ren["d_out"]]
    light = conf.get("light_network")
    return PathSpec(
        sdf_dims=sdf_dims, skip_in=tuple(imp.get("skip_in", ())), multires_x=imp["multires"],
        color_dims=color_dims, multires_d=ren["multires"],
        light_dims=([fvs] + list(light["dims"]) + [1]) if light else None,
        feature_size=fvs, bounding_sphere=float(conf.get("scene_bounding_sphere", 1.0)),
        near=float(smp["near"]), n_samples=smp["N_samples"], n_samples_eval=smp["N_samples_eval"],
        n_samples_extra=smp["N_samples_extra"], eps=float(smp["eps"]), beta_iters=smp["beta_iters"],
        max_total_iters=smp["max_total_iters"], add_tiny=float(smp.get("add_tiny", 0.0)),
        beta_min=float(conf["density"].get("beta_min", 1e-4)),
        use_normal=bool(conf.get("use_normal", False)) if use_normal is None else use_normal,
        detach_light_feature=bool(conf.get("detach_light_feature", True)),
    )


# --------------------------------------------------------------------------------------------------
# A.1 rays   (utils/rend_util.py:92-147, network/__init__.py:86-93)
# --------------------------------------------------------------------------------------------------
def camera_rays(uv: Tensor, pose: Tensor, intr: Tensor) -> Tuple[Tensor, Tensor]:
    """uv [B,P,2], pose [B,4,4], intr [B,4,4] -> un-normalised dirs [B,P,3], cam origin [B,3]."""
    fx, fy = intr[:, 0, 0, None], intr[:, 1, 1, None]
    cx, cy, sk = intr[:, 0, 2, None], intr[:, 1, 2, None], intr[:, 0, 1, None]
    u, v = uv[..., 0], uv[..., 1]
    one = torch.ones_like(u)
    xl = (u - cx + cy * sk / fy - sk * v / fy) / fx * one           # rend_util.py:143
    yl = (v - cy) / fy * one                                        # rend_util.py:144
    pc = torch.stack((xl, yl, one, one), dim=-1)                    # [B,P,4]
    world = torch.bmm(pose, pc.permute(0, 2, 1)).permute(0, 2, 1)[..., :3]   # rend_util.py:116
    cam = pose[:, :3, 3]
    return world - cam[:, None, :], cam


def flatten_rays(uv, pose, intr):
    """-> o [R,3], d [R,3] (unit), dnorm [R]   (network/__init__.py:86-93)."""
    dirs, cam = camera_rays(uv, pose, intr)
    B, P, _ = dirs.shape
    o = cam.unsqueeze(1).repeat(1, P, 1).reshape(-1, 3)
    dirs = dirs.reshape(-1, 3)
    dnorm = torch.linalg.vector_norm(dirs, dim=1)
    return o, F.normalize(dirs, dim=1), dnorm


# --------------------------------------------------------------------------------------------------
# A.2 embedding + MLPs   (embedder.py:6-38,138-152; mlp.py:10-229)
# --------------------------------------------------------------------------------------------------
def posenc(x: Tensor, n_freq: int) -> Tensor:
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^{L-1}x), cos(2^{L-1}x)]   (embedder.py:18-38)."""
    bands = 2.0 ** torch.linspace(0.0, n_freq - 1, n_freq)
    parts = [x]
    for f in bands:
        parts += [torch.sin(x * f), torch.cos(x * f)]
    return torch.cat(parts, -1)


def posenc_jacobian_T_apply(x: Tensor, r: Tensor, n_freq: int) -> Tensor:
    """J(x)^T r for the embedding above; r [M, 3+6L] -> [M,3].  (what autograd does at mlp.py:134-140)"""
    bands = 2.0 ** torch.linspace(0.0, n_freq - 1, n_freq)
    out = r[:, 0:3]
    c = 3
    for f in bands:
        out = out + r[:, c:c + 3] * (torch.cos(x * f) * f) - r[:, c + 3:c + 6] * (torch.sin(x * f) * f)
        c += 6
    return out


def weight_norm_weight(g: Tensor, v: Tensor) -> Tensor:
    """nn.utils.weight_norm(dim=0):  W = g * v / ||v||_row   (mlp.py:71-72, 200-201)."""
    return torch._weight_norm(v, g, 0)


def layer_params(P: Dict[str, Tensor], prefix: str, n_layers: int):
    """[(W [out,in], b [out])] from reference-named parameters  (SURVEY §8(b))."""
    out = []
    for l in range(n_layers):
        k = f"{prefix}.lin{l}"
        if f"{k}.weight_g" in P:
            W = weight_norm_weight(P[f"{k}.weight_g"], P[f"{k}.weight_v"])
        else:
            W = P[f"{k}.weight"]
        out.append((W, P[f"{k}.bias"]))
    return out


def softplus100(a: Tensor) -> Tensor:
    return F.softplus(a, beta=100.0)                                # nn.Softplus(beta=100), mlp.py:76


def sdf_mlp(spec: PathSpec, layers, x: Tensor, want_grad: bool = False):
    """ImplicitNetwork.forward (mlp.py:84-105) -> out [M, 1+F]; optionally grad_x out[:,0] [M,3]
    by an explicit reverse sweep (replaces autograd.grad at mlp.py:107-143)."""
    L = len(layers)
    e = posenc(x, spec.multires_x)
    h = e
    sig = []                       # sigmoid(100 a_l): derivative of softplus100
    inv_sqrt2 = 1.0 / math.sqrt(2.0)
    for l, (W, b) in enumerate(layers):
        if l in spec.skip_in:
            h = torch.cat([h, e], 1) / math.sqrt(2.0)               # mlp.py:94-95
        a = F.linear(h, W, b)
        if l < L - 1:
            h = softplus100(a)
            if want_grad:
                sig.append(torch.sigmoid(100.0 * a))
        else:
            h = a
    if not want_grad:
        return h, None
    # reverse sweep for d out[:,0] / d x
    r = layers[L - 1][0][0:1, :].expand(x.shape[0], -1)             # adjoint of input of last layer
    r_embed = torch.zeros_like(e)
    for l in range(L - 1, 0, -1):
        if l in spec.skip_in:                                       # input was cat[h, e]/sqrt2
            r = r * inv_sqrt2
            k = r.shape[1] - e.shape[1]
            r_embed = r_embed + r[:, k:]
            r = r[:, :k]
        r = (r * sig[l - 1]) @ layers[l - 1][0]                     # through softplus then lin{l-1}
    if 0 in spec.skip_in:
        raise NotImplementedError("skip at layer 0 is not used by any shipped config")
    r_embed = r_embed + r
    return h, posenc_jacobian_T_apply(x, r_embed, spec.multires_x)


def color_mlp(spec: PathSpec, layers, view_dirs: Tensor, feat: Tensor) -> Tensor:
    """RenderingNetwork.forward, mode 'nerf' (mlp.py:208-229)."""
    h = torch.cat([posenc(view_dirs, spec.multires_d), feat], -1)
    L = len(layers)
    for l, (W, b) in enumerate(layers):
        h = F.linear(h, W, b)
        if l < L - 1:
            h = torch.relu(h)
    return torch.sigmoid(h)


def light_mlp(layers, feat: Tensor) -> Tensor:
    """light-mask head: ImplicitNetwork(dims [F,128,1], no PE, sigmoid out)  (network/__init__.py:32)."""
    h = feat
    L = len(layers)
    for l, (W, b) in enumerate(layers):
        h = F.linear(h, W, b)
        if l < L - 1:
            h = softplus100(h)
    return torch.sigmoid(h)


def laplace_density(sdf: Tensor, beta) -> Tensor:
    """density.py:21-26."""
    return (1.0 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))


# --------------------------------------------------------------------------------------------------
# A.3 error-bounded sampler   (ray_sampler.py:67-251)
# --------------------------------------------------------------------------------------------------
def _error_bound(beta, sdf2d, dists, d_star):
    """ray_sampler.py:243-251; beta scalar tensor or [R,1]."""
    sigma = laplace_density(sdf2d, beta)
    sfe = torch.cat([torch.zeros(dists.shape[0], 1), dists * sigma[:, :-1]], dim=-1)
    integral = torch.cumsum(sfe, dim=-1)
    eps_i = torch.exp(-d_star / beta) * (dists ** 2.0) / (4 * beta ** 2)
    E = torch.cumsum(eps_i, dim=-1)
    bound = (torch.clamp(torch.exp(E), max=1.0e6) - 1.0) * torch.exp(-integral[:, :-1])
    return bound.max(-1)[0]


def d_star_bound(z: Tensor, sdf2d: Tensor) -> Tuple[Tensor, Tensor]:
    """Theorem-1 bound per section  (ray_sampler.py:98-114) -> (dists, d_star), both [R,n-1]."""
    dists = z[:, 1:] - z[:, :-1]
    a, b, c = dists, sdf2d[:, :-1].abs(), sdf2d[:, 1:].abs()
    c1 = a.pow(2) + b.pow(2) <= c.pow(2)
    c2 = a.pow(2) + c.pow(2) <= b.pow(2)
    s = (a + b + c) / 2.0
    area = s * (s - a) * (s - b) * (s - c)
    tri = ~c1 & ~c2 & (b + c - a > 0)
    c1 = c1 & ~c2
    d_star = c1 * b + c2 * c + torch.nan_to_num((2.0 * torch.sqrt(area)) / a) * tri
    d_star = (sdf2d[:, 1:].sign() * sdf2d[:, :-1].sign() == 1) * d_star
    return dists, d_star


def inverse_cdf(cdf: Tensor, bins: Tensor, u: Tensor):
    """ray_sampler.py:193-207 -> (samples [R,Ns], inds [R,Ns] int64)."""
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = c1 - c0
    small = denom < 1e-5
    denom = small + ~small * denom
    t = (u - c0) / denom
    return b0 + t * (b1 - b0), inds


def sample_z(spec: PathSpec, sdf_fn, o: Tensor, d: Tensor, beta0: Tensor, training: bool,
             tape: Optional[dict] = None, trace: Optional[dict] = None):
    """ErrorBoundSampler.get_z_vals  (ray_sampler.py:67-241).

    sdf_fn(points [M,3]) -> sdf [M,1]  (no-grad SDF, mlp.py:145-151).
    tape (training only): 'jitter' [R,N_eval], 'u_final' [R,N], 'extra_perm' callable n->LongTensor[N_extra]
    or LongTensor, 'eik_idx' [R] int64.   Eval uses the deterministic linspace variants.
    Returns z [R, N+2+N_extra], z_eik [R,1] (None if no eik_idx tape in eval)."""
    R = o.shape[0]
    tape = tape or {}
    near = spec.near * torch.ones(R, 1)
    far = spec.far * torch.ones(R, 1)
    t = torch.linspace(0.0, 1.0, steps=spec.n_samples_eval)
    z = near * (1.0 - t) + far * t                                   # ray_sampler.py:30-31
    if training:                                                     # stratified jitter :33-41
        mids = 0.5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * tape["jitter"]
    samples, perm = z, None
    dists = z[:, 1:] - z[:, :-1]
    bound = (1.0 / (4.0 * torch.log(torch.tensor(spec.eps + 1.0)))) * (dists ** 2.0).sum(-1)
    beta = torch.sqrt(bound)                                         # :75-77
    it, not_conv = 0, True
    sdf = None
    rounds = []
    while not_conv and it < spec.max_total_iters:
        pts = (o.unsqueeze(1) + samples.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
        with torch.no_grad():
            new_sdf = sdf_fn(pts)
        if perm is not None:                                         # :90-93
            merged = torch.cat([sdf.reshape(-1, z.shape[1] - samples.shape[1]),
                                new_sdf.reshape(-1, samples.shape[1])], -1)
            sdf = torch.gather(merged, 1, perm).reshape(-1, 1)
        else:
            sdf = new_sdf
        s2 = sdf.reshape(z.shape)
        dists, d_star = d_star_bound(z, s2)
        # beta line search :118-132
        err = _error_bound(beta0, s2, dists, d_star)
        ok = err <= spec.eps
        beta = beta * ~ok + beta0 * ok
        lo, hi = beta0.unsqueeze(0).repeat(R), beta
        for _ in range(spec.beta_iters):
            mid = (lo + hi) / 2.0
            err = _error_bound(mid.unsqueeze(-1), s2, dists, d_star)
            ok = err <= spec.eps
            hi = hi * ~ok + mid * ok
            lo = lo * ok + mid * ~ok
        beta = hi
        # weights with per-ray beta :139-147
        sigma = laplace_density(s2, beta.unsqueeze(-1))
        dists_inf = torch.cat([dists, torch.full([R, 1], 1e10)], -1)
        fe = dists_inf * sigma
        sfe = torch.cat([torch.zeros(R, 1), fe[:, :-1]], dim=-1)
        alpha = 1 - torch.exp(-fe)
        T = torch.exp(-torch.cumsum(sfe, dim=-1))
        w = alpha * T
        it += 1
        not_conv = bool(beta.max() > beta0)                          # batch-global, :151
        upsample = not_conv and it < spec.max_total_iters
        if upsample:                                                 # :153-171
            n_new = spec.n_samples_eval
            b = beta.unsqueeze(-1)
            eps_i = torch.exp(-d_star / b) * (dists_inf[:, :-1] ** 2.0) / (4 * b ** 2)
            E = torch.cumsum(eps_i, dim=-1)
            pdf = (torch.clamp(torch.exp(E), max=1.0e6) - 1.0) * T[:, :-1] + spec.add_tiny
        else:                                                        # :173-183
            n_new = spec.n_samples
            pdf = w[..., :-1] + 1e-5
        pdf = pdf / torch.sum(pdf, -1, keepdim=True)
        cdf = torch.cumsum(pdf, -1)
        cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
        if upsample or not training:                                 # :187-191
            u = torch.linspace(0.0, 1.0, steps=n_new).unsqueeze(0).repeat(R, 1)
        else:
            u = tape["u_final"]
        u = u.contiguous()
        samples, inds = inverse_cdf(cdf, z, u)
        rec = dict(z=z, sdf=s2, d_star=d_star, beta=beta, cdf=cdf, inds=inds, samples=samples,
                   weights=w, upsample=upsample)
        if upsample:                                                 # :211-212
            z, perm = torch.sort(torch.cat([z, samples], -1), -1)
            rec["perm"] = perm
            rec["z_merged"] = z
        rounds.append(rec)
    # post :215-234
    n = z.shape[1]
    if spec.n_samples_extra > 0:
        if training:
            ep = tape["extra_perm"]
            idx = ep(n) if callable(ep) else ep
        else:
            idx = torch.linspace(0, n - 1, spec.n_samples_extra).long()
        extra = torch.cat([near, far, z[:, idx]], -1)
    else:
        idx = None
        extra = torch.cat([near, far], -1)
    z_out, _ = torch.sort(torch.cat([samples, extra], -1), -1)
    z_eik = None
    if "eik_idx" in tape:
        z_eik = torch.gather(z_out, 1, tape["eik_idx"].unsqueeze(-1))
    if trace is not None:
        trace.update(rounds=rounds, n_rounds=it, n_final=n, extra_idx=idx, beta_final=beta)
    return z_out, z_eik


# --------------------------------------------------------------------------------------------------
# A.4 compositing   (network/__init__.py:223-240, 118-125)
# --------------------------------------------------------------------------------------------------
def composite_weights(z: Tensor, z_max: Tensor, sdf: Tensor, beta) -> Tuple[Tensor, Tensor]:
    sigma = laplace_density(sdf, beta).reshape(-1, z.shape[1])
    dists = torch.cat([z[:, 1:] - z[:, :-1], z_max.unsqueeze(-1) - z[:, -1:]], -1)
    fe = dists * sigma
    sfe = torch.cat([torch.zeros(dists.shape[0], 1), fe], dim=-1)
    alpha = 1 - torch.exp(-fe)
    T = torch.exp(-torch.cumsum(sfe, dim=-1))
    return alpha * T[:, :-1], T[:, -1]


# --------------------------------------------------------------------------------------------------
# I2SDFNetwork.forward   (network/__init__.py:80-221)
# --------------------------------------------------------------------------------------------------
def render(spec: PathSpec, P: Dict[str, Tensor], inputs: Dict[str, Tensor], training: bool,
           tape: Optional[dict] = None, predict_only: bool = False, trace: Optional[dict] = None,
           z_override: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """P: reference-named parameter dict (…lin{l}.weight_g/_v/bias, density.beta)."""
    tape = tape or {}
    sdf_layers = layer_params(P, "implicit_network", spec.n_sdf_layers)
    col_layers = layer_params(P, "rendering_network", spec.n_color_layers)
    o, d, dnorm = flatten_rays(inputs["uv"], inputs["pose"], inputs["intrinsics"])
    R = o.shape[0]
    beta_param = P["density.beta"].abs() + spec.beta_min              # density.py:28-30

    def sdf_only(p):
        return sdf_mlp(spec, sdf_layers, p)[0][:, :1]

    if z_override is None:
        z_all, z_eik = sample_z(spec, sdf_only, o, d, beta_param.detach(), training, tape, trace)
    else:
        z_all, z_eik = z_override, tape.get("z_eik")
    z_max, z = z_all[:, -1], z_all[:, :-1]                            # :99-100
    N = z.shape[1]
    pts = (o.unsqueeze(1) + z.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
    dirs = d.unsqueeze(1).repeat(1, N, 1).reshape(-1, 3)
    want_grad = spec.use_normal or (not training)                     # :109
    out, grad = sdf_mlp(spec, sdf_layers, pts, want_grad=want_grad)
    sdf, feat = out[:, :1], out[:, 1:]
    rgb = color_mlp(spec, col_layers, dirs, feat).reshape(-1, N, 3)
    w, _bg_T = composite_weights(z, z_max, sdf, beta_param)
    res = {
        "rgb_values": torch.sum(w.unsqueeze(-1) * rgb, 1),
        "depth_values": torch.sum(w * z, 1) / torch.clamp(dnorm, min=1e-6),
        "weight_sum": torch.sum(w, -1, keepdim=True),
    }
    if spec.light_dims is not None:                                   # :162-170
        lf = torch.relu(feat)
        if spec.detach_light_feature:
            lf = lf.detach()
        lm = light_mlp(layer_params(P, "light_network", len(spec.light_dims) - 1), lf)
        res["light_mask"] = torch.sum(w.unsqueeze(-1).detach() * lm.reshape(-1, N, 1), 1)
    if trace is not None:
        trace.update(z=z_all, sdf=sdf, feat=feat, grad=grad, rgb=rgb, weights=w, o=o, d=d)
    if predict_only:
        return res
    if training:                                                      # :175-209
        eik_u, nbr_u = tape["eik_uniform"], tape["nbr_uniform"]       # uniform_(-r,r), uniform_(-.005,.005)
        near_pts = (o.unsqueeze(1) + z_eik.unsqueeze(2) * d.unsqueeze(1)).reshape(-1, 3)
        all_pts = torch.cat([eik_u, near_pts, near_pts + nbr_u], 0)
        _, g = sdf_mlp(spec, sdf_layers, all_pts, want_grad=True)
        res["grad_theta"] = g[:2 * R]
        nrm = F.normalize(g[R:], dim=1, eps=1e-6)
        res["diff_norm"] = torch.norm(nrm[:R] - nrm[R:], dim=1)
        if "pointcloud" in inputs:                                    # :196-201
            sp = torch.cat([inputs["pointcloud"], o[tape["bubble_cam_idx"]].unsqueeze(0)], 0)
            res["surface_sdf"] = sdf_only(sp)[:-1]
        if spec.use_normal:
            n = F.normalize(grad, dim=-1).reshape(-1, N, 3)
            res["normal_values"] = F.normalize(torch.sum(w.unsqueeze(-1).detach() * n, 1), dim=-1)
    else:                                                             # :212-219
        n = F.normalize(grad.detach(), dim=-1).reshape(-1, N, 3)
        res["normal_map"] = F.normalize(torch.sum(w.unsqueeze(-1) * n, 1), dim=-1)
    return res


# --------------------------------------------------------------------------------------------------
# I2SDFLoss.forward restated as a plain function   (network/__init__.py:289-406)
# --------------------------------------------------------------------------------------------------
def recon_loss(out: Dict[str, Tensor], gt: Dict[str, Tensor], *, eikonal_weight=0.1, smooth_weight=0.0,
               depth_weight=0.1, normal_weight=0.05, angular_weight=0.05, bubble_weight=0.0,
               light_mask_weight=0.0, mask_weight=0.0, smooth_active=True) -> Tensor:
    loss = F.l1_loss(out["rgb_values"], gt["rgb"].reshape(-1, 3))
    if "grad_theta" in out:
        loss = loss + eikonal_weight * ((out["grad_theta"].norm(2, dim=1) - 1) ** 2).mean()
    if smooth_active and smooth_weight > 0 and "diff_norm" in out:
        loss = loss + smooth_weight * out["diff_norm"].mean()
    if "mask" in gt and mask_weight > 0:
        loss = loss + mask_weight * F.binary_cross_entropy(out["weight_sum"].clip(1e-3, 1 - 1e-3), gt["mask"])
    if "depth" in gt and depth_weight > 0:
        m = gt["depth_mask"].flatten()
        loss = loss + depth_weight * F.mse_loss(out["depth_values"][m], gt["depth"].flatten()[m])
    if "normal" in gt and (normal_weight > 0 or angular_weight > 0):
        m = gt["normal_mask"].flatten()
        l1 = torch.abs(1 - torch.sum(out["normal_values"][m] * gt["normal"].reshape(-1, 3)[m], dim=-1)).mean()
        loss = loss + (normal_weight + angular_weight) * l1           # the "angular" term re-uses the L1 loss (:368-371)
    if "surface_sdf" in out and bubble_weight > 0:
        loss = loss + bubble_weight * out["surface_sdf"].abs().mean()
    if "light_mask" in out and light_mask_weight > 0:
        loss = loss + light_mask_weight * F.binary_cross_entropy(
            out["light_mask"].reshape(-1, 1).clip(1e-3, 1 - 1e-3), gt["light_mask"].reshape(-1, 1))
    return loss


# --------------------------------------------------------------------------------------------------
# helpers shared by tests / bench (synthetic inputs of SURVEY §8(d))
# --------------------------------------------------------------------------------------------------
def synthetic_rays(R: int, seed: int = 1, train_layout: bool = False) -> Dict[str, Tensor]:
    """pose = I with t=(0,0,-1.5); fx=fy=300, cx=160, cy=120; uv ~ U([0,320]x[0,240])."""
    g = torch.Generator().manual_seed(seed)
    uv = torch.rand(R, 2, generator=g) * torch.tensor([320.0, 240.0])
    pose = torch.eye(4)
    pose[2, 3] = -1.5
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 300.0
    K[0, 2], K[1, 2] = 160.0, 120.0
    if train_layout:
        return {"uv": uv.reshape(R, 1, 2), "pose": pose.repeat(R, 1, 1), "intrinsics": K.repeat(R, 1, 1)}
    return {"uv": uv.reshape(1, R, 2), "pose": pose[None], "intrinsics": K[None]}


def init_params(spec: PathSpec, seed: int = 0, beta: float = 0.1, bias: float = 0.6,
                perturb: float = 0.0) -> Dict[str, Tensor]:
    """Reference-named parameters with the reference's geometric initialisation statistics
    (mlp.py:55-69) drawn from our own generator (NOT the reference's RNG stream; fixtures store the
    reference-initialised weights explicitly).  `perturb`>0 adds noise to the embedding columns that the
    geometric init zeroes so the positional encoding is exercised."""
    g = torch.Generator().manual_seed(seed)
    P: Dict[str, Tensor] = {}
    L = spec.n_sdf_layers
    d0 = spec.sdf_dims[0]
    for l in range(L):
        out, inn = spec.sdf_layer_shape(l)
        if l == L - 1:
            W = torch.randn(out, inn, generator=g) * 1e-4 + math.sqrt(math.pi) / math.sqrt(inn)
            b = torch.full((out,), -bias)
        else:
            W = torch.randn(out, inn, generator=g) * (math.sqrt(2) / math.sqrt(out))
            b = torch.zeros(out)
            if l == 0:
                W[:, 3:] = perturb * torch.randn(out, inn - 3, generator=g) * (math.sqrt(2) / math.sqrt(out))
            elif l in spec.skip_in:
                W[:, -(d0 - 3):] = perturb * torch.randn(out, d0 - 3, generator=g) * (math.sqrt(2) / math.sqrt(out))
        k = f"implicit_network.lin{l}"
        P[f"{k}.weight_v"] = W
        P[f"{k}.weight_g"] = W.norm(dim=1, keepdim=True)
        P[f"{k}.bias"] = b
    for name, dims in (("rendering_network", spec.color_dims), ("light_network", spec.light_dims)):
        if dims is None:
            continue
        for l in range(len(dims) - 1):
            bound = 1.0 / math.sqrt(dims[l])
            W = (torch.rand(dims[l + 1], dims[l], generator=g) * 2 - 1) * bound
            k = f"{name}.lin{l}"
            P[f"{k}.weight_v"] = W
            P[f"{k}.weight_g"] = W.norm(dim=1, keepdim=True)
            P[f"{k}.bias"] = (torch.rand(dims[l + 1], generator=g) * 2 - 1) * bound
    P["density.beta"] = torch.tensor(beta)
    return P
