"""Drop-in mirror of the reference's `model.network` surface, backed by the sm_100a kernels.

    from i2sdf_b200.network import I2SDFNetwork, I2SDFLoss        # instead of `from model.network import ...`

Same constructor node (`conf.model`), same sub-module / parameter names and shapes (so reference checkpoints load
with strict=True), same `forward(input, predict_only=False) -> dict` keys and shapes
(reference: model/network/__init__.py:19-221, mlp.py:10-229, density.py:5-30, ray_sampler.py:46-65).
The modules below only HOLD parameters and describe the network; all arithmetic of the per-ray path runs in
libi2sdf_b200.so through i2sdf_b200.core.RenderCore.  There is no PyTorch/CPU fallback: on a non-CUDA device
forward raises.
"""
import math
import weakref

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ._lib import I2SDFError
from .core import RenderCore


# Packed device copies of the weights must follow every parameter update.  Tensor version counters catch most in-place
# writes (load_state_dict, foreach/for-loop optimizers) but NOT the fused optimizers (torch._fused_adam_ leaves _version
# unchanged), so (a) a training-mode forward always re-packs from the effective weights it computes anyway, and (b) any
# optimizer step anywhere bumps this epoch, which is part of the key an eval-mode forward checks.
_OPT_EPOCH = [0]


def _bump_opt_epoch(*_args, **_kwargs):
    _OPT_EPOCH[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook
    _reg_post_hook(_bump_opt_epoch)
except Exception:                                   # very old torch: training forwards still re-pack every step
    pass


def _get(conf, key, default=None):
    if isinstance(conf, dict):
        return conf.get(key, default)
    return getattr(conf, key, default)


def _plain(conf):
    """CfgNode / dict / namespace -> plain nested dict."""
    if isinstance(conf, dict):
        return {k: _plain(v) for k, v in conf.items()}
    if isinstance(conf, (list, tuple)):
        return [_plain(v) for v in conf]
    if hasattr(conf, "__dict__") and not isinstance(conf, (int, float, str, bool)):
        return {k: _plain(v) for k, v in vars(conf).items() if not k.startswith("_")}
    return conf


def _embed_width(multires, d_in=3):
    return d_in + 2 * multires * d_in


def _make_linear(n_in, n_out, init=None, *, geo_bias=0.0, n_embed=0, use_weight_norm=True):
    """nn.Linear + the reference's geometric initialisation variants (mlp.py:52-72), drawn in the same order
    from the global RNG so that a seed reproduces the reference's weights."""
    lin = nn.Linear(n_in, n_out)
    if init == "sdf_out":
        nn.init.normal_(lin.weight, mean=np.sqrt(np.pi) / np.sqrt(n_in), std=0.0001)
        nn.init.constant_(lin.bias, -geo_bias)
    elif init == "first":
        nn.init.constant_(lin.bias, 0.0)
        nn.init.constant_(lin.weight[:, 3:], 0.0)
        nn.init.normal_(lin.weight[:, :3], 0.0, np.sqrt(2) / np.sqrt(n_out))
    elif init == "skip":
        nn.init.constant_(lin.bias, 0.0)
        nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(n_out))
        nn.init.constant_(lin.weight[:, -(n_embed - 3):], 0.0)
    elif init == "hidden":
        nn.init.constant_(lin.bias, 0.0)
        nn.init.normal_(lin.weight, 0.0, np.sqrt(2) / np.sqrt(n_out))
    if use_weight_norm:
        lin = nn.utils.weight_norm(lin)       # old-style API on purpose: checkpoint keys weight_g / weight_v
    return lin


def _effective_weight(lin):
    if hasattr(lin, "weight_g"):
        return torch._weight_norm(lin.weight_v, lin.weight_g, 0)
    return lin.weight


class _Stack(nn.Module):
    """Parameter holder for a chain of lin{l} layers."""

    def layers(self):
        return [getattr(self, f"lin{l}") for l in range(self.num_layers - 1)]

    def effective(self):
        ls = self.layers()
        return [_effective_weight(l) for l in ls], [l.bias for l in ls]

    def get_param_groups(self, lr):
        return [{"params": self.parameters(), "lr": lr}]


class ImplicitNetwork(_Stack):
    """SDF network (reference: mlp.py:10-151).  Holds lin0..lin{L-1}; arithmetic runs in the CUDA core."""

    def __init__(self, feature_vector_size, sdf_bounding_sphere, d_in, d_out, dims, geometric_init=True, bias=1.0,
                 skip_in=(), weight_norm=True, embed_type=None, sphere_scale=1.0, output_activation=None, **kwargs):
        super().__init__()
        if sdf_bounding_sphere and sdf_bounding_sphere > 0.0:
            raise I2SDFError("sdf_bounding_sphere > 0 (sphere clamp) is not used by I2SDFNetwork and not supported")
        self.sdf_bounding_sphere = sdf_bounding_sphere
        self.sphere_scale = sphere_scale
        dims = [d_in] + list(dims) + [d_out + feature_vector_size]
        self.embed_type = embed_type
        self.multires = kwargs.get("multires", 0)
        if embed_type:
            if embed_type != "positional":
                raise I2SDFError(f"embed_type {embed_type!r} is not supported (shipped configs use 'positional')")
            dims[0] = _embed_width(self.multires, d_in)
        print(f"[INFO] Implicit network dims: {dims}")
        self.dims = dims
        self.num_layers = len(dims)
        self.skip_in = tuple(skip_in)
        self.weight_norm = weight_norm
        self.output_activation_name = output_activation
        last = self.num_layers - 2
        for l in range(self.num_layers - 1):
            n_out = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            kind = None
            if geometric_init:
                if l == last:
                    kind = "sdf_out"
                elif embed_type and l == 0:
                    kind = "first"
                elif embed_type and l in self.skip_in:
                    kind = "skip"
                else:
                    kind = "hidden"
            setattr(self, f"lin{l}", _make_linear(dims[l], n_out, kind, geo_bias=bias, n_embed=dims[0],
                                                  use_weight_norm=weight_norm))
        self._owner = None

    # --- standalone callable surface used by the reference's eval / plotting code
    def _core(self):
        owner = self._owner() if self._owner is not None else None
        if owner is None or owner.implicit_network is not self:
            raise I2SDFError("this ImplicitNetwork is a parameter holder; call it through I2SDFNetwork.implicit_network")
        return owner._ready_core()

    def forward(self, input):
        """x [M,3] -> [M, 1+F]  (mlp.py:84-105); used by mesh extraction as implicit_network(pts)[:,0]."""
        sdf, feat, _ = self._core().sdf_forward(input, want_feat=True)
        return torch.cat([sdf[:, None], feat], 1)

    def get_sdf_vals(self, x):
        return self._core().sdf_forward(x)[0][:, None]

    def gradient(self, x):
        owner = self._owner()
        if owner is not None and torch.is_grad_enabled() and owner.training:
            from .autograd import sdf_with_grad
            return sdf_with_grad(owner, x)[2]
        return self._core().sdf_forward(x, want_grad=True)[2]

    def feature(self, x):
        return self._core().sdf_forward(x, want_feat=True)[1]

    def get_outputs(self, x, returns_grad=True):
        sdf, feat, grad = self._core().sdf_forward(x, want_feat=True, want_grad=returns_grad)
        return sdf[:, None], feat, grad


class RenderingNetwork(_Stack):
    """Radiance network (reference: mlp.py:159-229), mode 'nerf'."""

    def __init__(self, feature_vector_size, mode, d_in, d_out, dims, weight_norm=True, embed_type=None,
                 embed_point=None, output_activation="sigmoid", **kwargs):
        super().__init__()
        if mode != "nerf":
            raise I2SDFError("rendering_network.mode 'idr' is not configured by any shipped yaml and not supported")
        if output_activation != "sigmoid":
            raise I2SDFError("rendering_network output_activation must be 'sigmoid'")
        self.mode = mode
        self.d_out = d_out
        dims = [d_in + feature_vector_size] + list(dims) + [d_out]
        self.multires = kwargs.get("multires", 0)
        if embed_type:
            if embed_type != "positional":
                raise I2SDFError(f"embed_type {embed_type!r} is not supported")
            dims[0] += _embed_width(self.multires, 3) - 3
        print(f"[INFO] Rendering network dims: {dims}")
        self.dims = dims
        self.num_layers = len(dims)
        self.weight_norm = weight_norm
        for l in range(self.num_layers - 1):
            setattr(self, f"lin{l}", _make_linear(dims[l], dims[l + 1], None, use_weight_norm=weight_norm))

    def forward(self, points, normals, view_dirs, feature_vectors):
        raise I2SDFError("RenderingNetwork is evaluated inside the fused render kernel; call I2SDFNetwork.forward")


class LaplaceDensity(nn.Module):
    """alpha * Laplace(0, beta).cdf(-sdf)  (reference: density.py:5-30).  `beta` stays an nn.Parameter."""

    def __init__(self, params_init={}, beta_min=0.0001):
        super().__init__()
        for p in params_init:
            setattr(self, p, nn.Parameter(torch.tensor(params_init[p])))
        self.beta_min = torch.tensor(beta_min)

    def get_beta(self):
        return self.beta.abs() + self.beta_min.to(self.beta.device)

    def density_func(self, sdf, beta=None):
        # API-compat only (not on the hot path, which evaluates the density inside the kernels)
        if beta is None:
            beta = self.get_beta()
        return (1 / beta) * (0.5 + 0.5 * sdf.sign() * torch.expm1(-sdf.abs() / beta))

    def forward(self, sdf, beta=None):
        return self.density_func(sdf, beta=beta)


class ErrorBoundSampler:
    """Configuration holder for the error-bounded sampler (reference: ray_sampler.py:46-65)."""

    def __init__(self, scene_bounding_sphere, near, N_samples, N_samples_eval, N_samples_extra, eps, beta_iters,
                 max_total_iters, inverse_sphere_bg=False, N_samples_inverse_sphere=0, add_tiny=0.0):
        if inverse_sphere_bg:
            raise I2SDFError("inverse-sphere background is not enabled by any shipped config and not supported")
        self.near, self.far = near, 2.0 * scene_bounding_sphere
        self.N_samples, self.N_samples_eval, self.N_samples_extra = N_samples, N_samples_eval, N_samples_extra
        self.eps, self.beta_iters, self.max_total_iters = eps, beta_iters, max_total_iters
        self.scene_bounding_sphere = scene_bounding_sphere
        self.add_tiny = add_tiny
        self.inverse_sphere_bg = False

    def get_z_vals(self, ray_dirs, cam_loc, model):
        """-> (z_vals [R,98], z_samples_eik [R,1])   (ray_sampler.py:67-241)."""
        core = model._ready_core()
        tape = model._draw_sampler_tape(ray_dirs.shape[0], ray_dirs.device) if model.training else None
        if tape is None:
            tape = {"eik_idx": torch.randint(core.n_out, (ray_dirs.shape[0],), device=ray_dirs.device)}
        z, z_eik = core.sample(cam_loc.contiguous().float(), ray_dirs.contiguous().float(),
                               model.density.beta.detach(), tape)
        return z, z_eik[:, None]


class I2SDFNetwork(nn.Module):
    """Reference: model/network/__init__.py:19-221."""

    def __init__(self, conf):
        super().__init__()
        self._model_conf = _plain(conf)
        mc = self._model_conf
        self.feature_vector_size = mc["feature_vector_size"]
        self.scene_bounding_sphere = mc.get("scene_bounding_sphere", 1.0)
        self.implicit_network = ImplicitNetwork(self.feature_vector_size, 0.0, **mc["implicit_network"])
        self.rendering_network = RenderingNetwork(self.feature_vector_size, **mc["rendering_network"])
        self.use_light = "light_network" in mc
        if self.use_light:
            self.light_network = ImplicitNetwork(0, 0, d_in=self.feature_vector_size, d_out=1, geometric_init=False,
                                                 embed_type=None, output_activation="sigmoid", **mc["light_network"])
        self.density = LaplaceDensity(**mc["density"])
        self.use_bg = "bg_network" in mc
        if self.use_bg:
            raise I2SDFError("model.bg_network (inverse-sphere background) is out of scope: no shipped config enables it")
        print("[INFO] BG Network Disabled")
        self.ray_sampler = ErrorBoundSampler(self.scene_bounding_sphere, inverse_sphere_bg=False, **mc["ray_sampler"])
        self.use_normal = mc.get("use_normal", False)
        self.detach_light_feature = mc.get("detach_light_feature", True)
        if not self.detach_light_feature:
            raise I2SDFError("detach_light_feature=False is not supported (the reference default is True)")
        self.implicit_network._owner = weakref.ref(self)
        self._core_obj = None
        self._packed_key = None
        self._tape_override = None       # test hook: dict of RNG tapes / reference z's (see autograd.forward_train)
        # rays of one batch sharded over several GPUs: a torch.distributed group here makes the sampler's convergence test
        # batch-global again (parallel.use_global_convergence); None = per-call test, as a single-GPU reference run has it
        self.convergence_group = None          # training forwards: batch-global sampler convergence over this group (parallel.use_global_convergence)
        self.convergence_group_eval = None     # eval forwards: explicit opt-in only (rank-0-only validation must not enter a collective)

    def get_param_groups(self, lr):
        return [{"params": self.parameters(), "lr": lr}]

    # ------------------------------------------------------------------ device state
    def _stacks(self):
        s = [self.implicit_network, self.rendering_network]
        if self.use_light:
            s.append(self.light_network)
        return s

    def _param_key(self):
        return (_OPT_EPOCH[0],) + tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ready_core(self, defer_pack=False) -> RenderCore:
        """The RenderCore on the parameters' device with up-to-date packed weights (defer_pack: the caller packs itself)."""
        dev = self.density.beta.device
        if dev.type != "cuda":
            raise I2SDFError("I2SDFNetwork (i2sdf_b200) needs its parameters on a CUDA device; there is no CPU path")
        if self._core_obj is None or self._core_obj.device != dev:
            self._core_obj = RenderCore(self._model_conf, dev)
            self._packed_key = None
        if not defer_pack:
            key = self._param_key()
            if key != self._packed_key:
                self.pack_weights()
                self._packed_key = key
        return self._core_obj

    def effective_weights(self):
        """W = g v / ||v|| of every layer (mlp.py:71-72) + biases, SDF stack first, then radiance, then light.  On CUDA with
        all layers weight-normed (the shipped configs) this is ONE launch (and one for its backward), csrc/wnorm.cu."""
        layers = [l for st in self._stacks() for l in st.layers()]
        bs = [l.bias for l in layers]
        if len(layers) <= 28 and all(hasattr(l, "weight_g") for l in layers) and layers[0].weight_v.is_cuda:
            from .autograd import weight_norm_all
            return weight_norm_all(layers), bs
        return [_effective_weight(l) for l in layers], bs

    @torch.no_grad()
    def pack_weights(self):
        Ws, bs = self.effective_weights()
        self._core_obj.pack(Ws, bs)

    # ------------------------------------------------------------------ RNG tapes (training)
    def _draw_sampler_tape(self, R, device):
        """Same draws, same order, same devices as the reference (ray_sampler.py:39,190,223,233)."""
        rs = self.ray_sampler
        n_out = rs.N_samples + 2 + rs.N_samples_extra
        return {
            "jitter": torch.rand(R, rs.N_samples_eval, device=device),
            "u_final": torch.rand(R, rs.N_samples, device=device),
            "extra_perm": lambda n: torch.randperm(n)[:rs.N_samples_extra],          # CPU generator, as the reference
            "eik_idx_fn": lambda: torch.randint(n_out, (R,), device=device),
        }

    # ------------------------------------------------------------------ forward
    def forward(self, input, predict_only=False):
        if self.training:
            from .autograd import forward_train
            # forward_train packs the effective weights it builds for autograd (every step: the optimizer moved them)
            return forward_train(self, self._ready_core(defer_pack=True), input, predict_only)
        return self._forward_nograd(self._ready_core(), input, predict_only)

    @torch.no_grad()
    def _forward_nograd(self, core, input, predict_only):
        o, d, dnorm = core.rays(input["uv"], input["pose"], input["intrinsics"])
        R = o.shape[0]
        beta = self.density.beta.detach()
        z, z_eik = core.sample(o, d, beta, None, group=self.convergence_group_eval)
        # grad_x is always evaluated (the reference's returns_grad is True in eval, network/__init__.py:109); this also
        # keeps predict_only calls on the same kernels, hence bit-identical to the full call
        out = core.render(o, d, dnorm, z, beta, want_normal=True, want_light=self.use_light)
        res = {"rgb_values": out["rgb"], "depth_values": out["depth"], "weight_sum": out["weight_sum"][:, None]}
        if self.use_light:
            res["light_mask"] = out["light"][:, None]
        if not predict_only:
            res["normal_map"] = out["normal"]
        return res


class I2SDFLoss(nn.Module):
    """Loss of the reconstruction stage; consumes I2SDFNetwork outputs (reference: model/network/__init__.py:289-406).
    The whole loss and its backward seed are one kernel (csrc/loss.cu); CPU tensors are rejected."""

    def __init__(self, eikonal_weight=0.1, smooth_weight=0.0, mask_weight=0.0, depth_weight=0.1, normal_weight=0.05,
                 angular_weight=0.05, bubble_weight=0.0, min_bubble_iter=0, max_bubble_iter=None, smooth_iter=None,
                 light_mask_weight=0.0, eikonal_weight_bubble=0.0):
        super().__init__()
        self.eikonal_weight, self.smooth_weight, self.mask_weight = eikonal_weight, smooth_weight, mask_weight
        self.depth_weight, self.normal_weight, self.angular_weight = depth_weight, normal_weight, angular_weight
        self.bubble_weight, self.light_mask_weight = bubble_weight, light_mask_weight
        self.min_bubble_iter, self.max_bubble_iter, self.smooth_iter = min_bubble_iter, max_bubble_iter, smooth_iter
        self.rgb_loss = F.l1_loss
        # rays sharded over ranks: divide by the GLOBAL counts (parallel.use_global_loss_means); None = this call's own counts
        self.means_group = None
        if self.bubble_weight > 0 and self.max_bubble_iter is not None and self.smooth_iter < self.max_bubble_iter:
            self.smooth_iter = self.max_bubble_iter

    # Masked means are evaluated as sum(mask * x) / sum(mask): same value as the reference's x[mask].mean()
    # (model/network/__init__.py:320-329) — NaN for an empty mask included — but without the boolean-index gather, whose
    # nonzero() forces a device->host sync in the middle of every training step.
    @staticmethod
    def _masked_mean(x, mask):
        m = mask.flatten().bool()
        return torch.where(m, x, torch.zeros((), dtype=x.dtype, device=x.device)).sum() / m.sum()

    @classmethod
    def _masked_normal_l1(cls, normal, normal_gt, mask):
        return cls._masked_mean(torch.abs(1 - torch.sum(normal * normal_gt.reshape(-1, 3), dim=-1)), mask)

    def get_rgb_loss(self, rgb_values, rgb_gt):
        return self.rgb_loss(rgb_values, rgb_gt.reshape(-1, 3))

    def get_eikonal_loss(self, grad_theta):
        return ((grad_theta.norm(2, dim=1) - 1) ** 2).mean()

    def get_mask_loss(self, mask_pred, mask_gt):
        return F.binary_cross_entropy(mask_pred.clip(1e-3, 1.0 - 1e-3), mask_gt)

    def get_depth_loss(self, depth, depth_gt, depth_mask):
        return self._masked_mean((depth.flatten() - depth_gt.flatten()) ** 2, depth_mask)

    def get_normal_l1_loss(self, normal, normal_gt, normal_mask):
        return self._masked_normal_l1(normal, normal_gt, normal_mask)

    def get_normal_angular_loss(self, normal, normal_gt, normal_mask):
        dot = torch.sum(normal * normal_gt.reshape(-1, 3), dim=-1)
        return self._masked_mean((torch.acos(torch.clamp(dot, -1.0 + 1e-6, 1.0 - 1e-6)) / math.tau).clamp_max(0.5).abs(), normal_mask)

    def forward(self, model_outputs, ground_truth, current_step):
        if not model_outputs["rgb_values"].is_cuda:
            # no CPU path in the product: I2SDFNetwork only produces CUDA tensors.  (_forward_torch below is the PyTorch
            # restatement the tests check the kernel against; it is not reachable from here.)
            from ._lib import I2SDFError
            raise I2SDFError("I2SDFLoss runs on CUDA tensors only (csrc/loss.cu, no CPU fallback); got model outputs on "
                             f"{model_outputs['rgb_values'].device}")
        return self._forward_fused(model_outputs, ground_truth, current_step)

    def _shard_denominators(self, n_rays, n_eik, n_bubble, depth_mask, normal_mask, dev):
        """Divisors of the means when the batch is sharded over the ranks of `means_group` (SURVEY.md §8(e) caveat 2): the
        reference's means run over the whole batch (model/network/__init__.py:308-329), so every rank divides its sums by
        (global count) / world — the average over ranks of the returned losses, and of the parameter gradients that
        `parallel.allreduce_gradients` averages, is then the single-GPU loss / gradient of the whole batch even when the shards
        are ragged or their mask counts differ.  One 40-byte SUM all-reduce in stream order; no host sync.  Returns float32 [5]
        in the I2SDF_LOSS_DENOM_* order (rays, eikonal rows, bubble points, depth-mask count, normal-mask count) or None."""
        import torch.distributed as dist
        g = self.means_group
        if g is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(g) == 1:
            return None
        cnt = torch.zeros(5, dtype=torch.float64, device=dev)
        cnt[0].fill_(float(n_rays))
        cnt[1].fill_(float(n_eik))
        cnt[2].fill_(float(n_bubble))
        if depth_mask is not None:
            cnt[3].copy_(depth_mask.to(dev).sum())
        if normal_mask is not None:
            cnt[4].copy_(normal_mask.to(dev).sum())
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=g)
        return (cnt / dist.get_world_size(g)).float()

    _TERMS = ("loss", "rgb_loss", "eikonal_loss", "smooth_loss", "mask_loss", "depth_loss", "normal_loss", "angular_loss",
              "bubble_loss", "light_mask_loss")

    def _forward_fused(self, out, gt, current_step):
        """CUDA tensors: every term, the total and d loss / d output in ONE launch (csrc/loss.cu, i2sdf_loss_forward); the
        reference's ~60 small loss / loss-backward kernels per step become 1 + a multi-tensor scale in backward."""
        from .autograd import fused_loss
        smooth_on = self.smooth_iter is None or current_step > self.smooth_iter
        has_n = "normal" in gt and (self.normal_weight > 0 or self.angular_weight > 0)
        sel = dict(
            rgb=out["rgb_values"], rgb_gt=gt["rgb"],
            grad_theta=out.get("grad_theta"),
            diff_norm=out["diff_norm"] if (smooth_on and self.smooth_weight > 0 and "diff_norm" in out) else None,
            weight_sum=out["weight_sum"] if ("mask" in gt and self.mask_weight > 0) else None, mask_gt=gt.get("mask"),
            depth=out["depth_values"] if ("depth" in gt and self.depth_weight > 0) else None, depth_gt=gt.get("depth"), depth_mask=gt.get("depth_mask"),
            normal=out["normal_values"] if has_n else None, normal_gt=gt.get("normal"), normal_mask=gt.get("normal_mask"),
            surface_sdf=out["surface_sdf"] if ("surface_sdf" in out and self.bubble_weight > 0) else None,
            light=out["light_mask"] if ("light_mask" in out and self.light_mask_weight > 0) else None, light_gt=gt.get("light_mask"))
        w = dict(w_eik=self.eikonal_weight, w_smooth=self.smooth_weight, w_mask=self.mask_weight, w_depth=self.depth_weight,
                 w_normal=self.normal_weight, w_angular=self.angular_weight, w_bubble=self.bubble_weight, w_light=self.light_mask_weight)
        sel["denom"] = self._shard_denominators(
            sel["rgb"].shape[0], 0 if sel["grad_theta"] is None else sel["grad_theta"].shape[0],
            0 if sel["surface_sdf"] is None else sel["surface_sdf"].numel(),
            sel["depth_mask"] if sel["depth"] is not None else None, sel["normal_mask"] if sel["normal"] is not None else None,
            sel["rgb"].device)
        terms = fused_loss(sel, w)
        res = {"loss": terms[0]}
        det = terms.detach()
        for i, k in enumerate(self._TERMS[1:], 1):
            res[k] = det[i]
        return res

    def _forward_torch(self, model_outputs, ground_truth, current_step):
        """Plain PyTorch restatement of the reference's loss (model/network/__init__.py:338-406).  CHECKER ONLY: the tests compare
        the CUDA kernel with it (and it with the oracle's loss, which the fixtures pin on the reference); forward() never calls it."""
        dev = model_outputs["rgb_values"].device
        zero = lambda: torch.zeros((), device=dev)                # noqa: E731   (a fill kernel: no host->device copy, no sync)
        if self.means_group is not None:
            has_d = "depth" in ground_truth and self.depth_weight > 0
            has_nn = "normal" in ground_truth and (self.normal_weight > 0 or self.angular_weight > 0)
            den = self._shard_denominators(
                model_outputs["rgb_values"].shape[0], model_outputs["grad_theta"].shape[0] if "grad_theta" in model_outputs else 0,
                model_outputs["surface_sdf"].numel() if ("surface_sdf" in model_outputs and self.bubble_weight > 0) else 0,
                ground_truth["depth_mask"] if has_d else None, ground_truth["normal_mask"] if has_nn else None, dev)
            if den is not None:
                return self._forward_torch_sharded(model_outputs, ground_truth, current_step, den)
        terms = {"rgb_loss": self.get_rgb_loss(model_outputs["rgb_values"], ground_truth["rgb"])}
        terms["eikonal_loss"] = self.get_eikonal_loss(model_outputs["grad_theta"]) if "grad_theta" in model_outputs else zero()
        smooth_on = self.smooth_iter is None or current_step > self.smooth_iter
        terms["smooth_loss"] = model_outputs["diff_norm"].mean() if (smooth_on and self.smooth_weight > 0 and "diff_norm" in model_outputs) else zero()
        terms["mask_loss"] = self.get_mask_loss(model_outputs["weight_sum"], ground_truth["mask"]) if ("mask" in ground_truth and self.mask_weight > 0) else zero()
        terms["depth_loss"] = self.get_depth_loss(model_outputs["depth_values"], ground_truth["depth"], ground_truth["depth_mask"]) if ("depth" in ground_truth and self.depth_weight > 0) else zero()
        has_n = "normal" in ground_truth
        terms["normal_loss"] = self.get_normal_l1_loss(model_outputs["normal_values"], ground_truth["normal"], ground_truth["normal_mask"]) if (has_n and self.normal_weight > 0) else zero()
        # the reference's "angular" term re-uses the L1 normal loss (model/network/__init__.py:368-371)
        terms["angular_loss"] = self.get_normal_l1_loss(model_outputs["normal_values"], ground_truth["normal"], ground_truth["normal_mask"]) if (has_n and self.angular_weight > 0) else zero()
        terms["bubble_loss"] = model_outputs["surface_sdf"].abs().mean() if ("surface_sdf" in model_outputs and self.bubble_weight > 0) else zero()
        terms["light_mask_loss"] = self.get_mask_loss(model_outputs["light_mask"].reshape(-1, 1), ground_truth["light_mask"].reshape(-1, 1)) if ("light_mask" in model_outputs and self.light_mask_weight > 0) else zero()
        weights = {"rgb_loss": 1.0, "eikonal_loss": self.eikonal_weight, "smooth_loss": self.smooth_weight,
                   "mask_loss": self.mask_weight, "depth_loss": self.depth_weight, "normal_loss": self.normal_weight,
                   "angular_loss": self.angular_weight, "bubble_loss": self.bubble_weight,
                   "light_mask_loss": self.light_mask_weight}
        loss = terms["rgb_loss"]
        for k in ("eikonal_loss", "smooth_loss", "mask_loss", "depth_loss", "normal_loss", "angular_loss", "bubble_loss", "light_mask_loss"):
            loss = loss + weights[k] * terms[k]
        out = {"loss": loss}
        out.update(terms)
        return out

    def _forward_torch_sharded(self, out, gt, current_step, den):
        """_forward_torch with every mean written as (this shard's sum) / den[...] (see _shard_denominators)."""
        dev = out["rgb_values"].device
        zero = lambda: torch.zeros((), device=dev)                # noqa: E731
        n_ray, n_eik, n_bub, n_depth, n_normal = den[0], den[1], den[2], den[3], den[4]

        def msum(x, mask):
            return torch.where(mask.flatten().bool(), x, torch.zeros((), dtype=x.dtype, device=dev)).sum()

        def bce_sum(p, t):
            return F.binary_cross_entropy(p.clip(1e-3, 1.0 - 1e-3), t, reduction="sum")

        def normal_l1():
            return msum(torch.abs(1 - torch.sum(out["normal_values"] * gt["normal"].reshape(-1, 3), dim=-1)), gt["normal_mask"]) / n_normal

        terms = {"rgb_loss": (out["rgb_values"] - gt["rgb"].reshape(-1, 3)).abs().sum() / (3 * n_ray)}
        terms["eikonal_loss"] = ((out["grad_theta"].norm(2, dim=1) - 1) ** 2).sum() / n_eik if "grad_theta" in out else zero()
        smooth_on = self.smooth_iter is None or current_step > self.smooth_iter
        terms["smooth_loss"] = out["diff_norm"].sum() / n_ray if (smooth_on and self.smooth_weight > 0 and "diff_norm" in out) else zero()
        terms["mask_loss"] = bce_sum(out["weight_sum"], gt["mask"]) / n_ray if ("mask" in gt and self.mask_weight > 0) else zero()
        terms["depth_loss"] = (msum((out["depth_values"].flatten() - gt["depth"].flatten()) ** 2, gt["depth_mask"]) / n_depth
                               if ("depth" in gt and self.depth_weight > 0) else zero())
        has_n = "normal" in gt
        terms["normal_loss"] = normal_l1() if (has_n and self.normal_weight > 0) else zero()
        terms["angular_loss"] = normal_l1() if (has_n and self.angular_weight > 0) else zero()
        terms["bubble_loss"] = out["surface_sdf"].abs().sum() / n_bub if ("surface_sdf" in out and self.bubble_weight > 0) else zero()
        terms["light_mask_loss"] = (bce_sum(out["light_mask"].reshape(-1, 1), gt["light_mask"].reshape(-1, 1)) / n_ray
                                    if ("light_mask" in out and self.light_mask_weight > 0) else zero())
        weights = {"eikonal_loss": self.eikonal_weight, "smooth_loss": self.smooth_weight, "mask_loss": self.mask_weight,
                   "depth_loss": self.depth_weight, "normal_loss": self.normal_weight, "angular_loss": self.angular_weight,
                   "bubble_loss": self.bubble_weight, "light_mask_loss": self.light_mask_weight}
        loss = terms["rgb_loss"]
        for k, w in weights.items():
            loss = loss + w * terms[k]
        res = {"loss": loss}
        res.update(terms)
        return res
