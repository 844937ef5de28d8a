"""Training-mode forward/backward of I2SDFNetwork on the CUDA core.

torch.autograd is used only as the tape between the loss (I2SDFLoss, PyTorch) and three custom Functions whose
forward AND backward are library calls (C ABI): the main-pass MLP chain, the compositing, and stand-alone SDF
(+grad_x) evaluations for the eikonal / smoothness / bubble terms.  Weight-norm (W = g v/||v||) stays a PyTorch op
so the Functions see the effective weights as differentiable inputs and return dL/dW for them.

Reference behaviour reproduced (file:line in jingsenzhu/i2-sdf): model/network/__init__.py:99-125 (main pass),
:162-170 (light mask, detached), :175-209 (training extras), mlp.py:107-143 (create_graph gradient).
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch.autograd import Function


# parameter -> persistent gradient view (parallel.GradBucket): the weight-norm backward writes dg / dv straight into these
import weakref


def register_grad_view(param, view):
    param._i2sdf_grad_view = view         # an attribute of the Parameter object (tensors do not hash / compare as dictionary keys)


def _zeros_like_all(groups):
    """Zero-initialised gradient buffers for several lists of tensors, carved from ONE flat allocation (one fill kernel)."""
    flat_n = sum(t.numel() for g in groups for t in g)
    if flat_n == 0:
        return [[] for _ in groups]
    ref = next(t for g in groups for t in g)
    flat = torch.zeros(flat_n + 4 * sum(len(g) for g in groups), dtype=ref.dtype, device=ref.device)
    out, off = [], 0
    for g in groups:
        cur = []
        for t in g:
            n = t.numel()
            cur.append(flat[off:off + n].view(t.shape))
            off += (n + 3) // 4 * 4                      # keep every buffer 16-byte aligned
        out.append(cur)
    return out


def _split_params(params, n_sdf, n_col, n_light):
    i = 0
    out = []
    for n in (n_sdf, n_sdf, n_col, n_col, n_light, n_light):
        out.append(list(params[i:i + n]))
        i += n
    return out      # W_sdf, b_sdf, W_col, b_col, W_light, b_light


class _PointsFn(Function):
    """(o, d, z) -> per-sample sdf, grad_x sdf, rgb, light-mask.  Differentiable w.r.t. the effective weights.
    extra_pts [E,3] (or None; fused tensor-core path only): explicit points evaluated in the SAME launches, appended
    after the R*N ray samples; their grad_x sdf is returned as a fifth output (eikonal / smoothness points)."""

    @staticmethod
    def forward(ctx, core, o, d, z, extra_pts, want_grad, n_sdf, n_col, n_light, *params):
        out = core.points_forward(o, d, z, want_grad, n_light > 0, save=True, extra_pts=extra_pts)
        ctx.core, ctx.cfg = core, (want_grad, n_sdf, n_col, n_light)
        ctx.fused = out["fused"]
        ctx.set_materialize_grads(False)
        E = 0 if extra_pts is None else extra_pts.shape[0]
        M = z.shape[0] * (z.shape[1] - 1)
        keep = [o, d, z, out["act"], out["feat"] if out["feat"] is not None else o.new_empty(0), out["s_rgb"]]
        keep.append(out["s_light"] if out["s_light"] is not None else o.new_empty(0))
        keep.append(extra_pts.detach() if E else o.new_empty(0))
        ctx.save_for_backward(*keep, *params)
        s_grad = out["s_grad"][:M] if want_grad else o.new_empty(0)
        s_light = out["s_light"] if n_light > 0 else o.new_empty(0)
        x_grad = out["s_grad"][M:] if E else o.new_empty(0)
        if not want_grad:
            ctx.mark_non_differentiable(s_grad)
        if n_light == 0:
            ctx.mark_non_differentiable(s_light)
        if not E:
            ctx.mark_non_differentiable(x_grad)
        return out["s_sdf"][:M], s_grad, out["s_rgb"][:M], s_light, x_grad

    @staticmethod
    def backward(ctx, g_sdf, g_grad, g_rgb, g_light, g_xgrad):
        core = ctx.core
        want_grad, n_sdf, n_col, n_light = ctx.cfg
        saved = ctx.saved_tensors
        o, d, z, act, feat, s_rgb, s_light, extra_pts = saved[:8]
        W_sdf, b_sdf, W_col, b_col, W_l, b_l = _split_params(saved[8:], n_sdf, n_col, n_light)
        R, N = z.shape[0], z.shape[1] - 1
        M = R * N
        E = extra_pts.shape[0]
        dW_sdf, db_sdf, dW_col, db_col, dW_l, db_l = _zeros_like_all([W_sdf, b_sdf, W_col, b_col, W_l, b_l])
        if ctx.fused:
            # plane slots: one chain kernel + one weight-gradient launch cover both stacks, first and second order
            if not want_grad:
                g_grad = None
            if g_sdf is not None or g_rgb is not None or g_grad is not None or (E and g_xgrad is not None):
                # the upstream arrays cover the M ray samples; the appended points only carry an upstream of grad_x sdf (no concatenation
                # with zero blocks: i2sdf_fused_backward_ex takes the two pieces as they are)
                core.fused_backward(M + E, act, dW_sdf, db_sdf, rays=(o, d, z, N), pts=extra_pts if E else None, s_rgb=s_rgb, g_sdf=g_sdf,
                                    g_grad=g_grad, g_rgb=g_rgb, dW_col=dW_col, db_col=db_col, m_up=M, g_grad_tail=g_xgrad if E else None)
            if n_light > 0 and g_light is not None:
                # the head's own parameters only: its input features are detached (network/__init__.py:165)
                hidden = act.data_ptr() + act.numel() - (M + E) * core.desc.light_hidden * 4     # behind the plane slots (i2sdf_b200.h)
                core.light_backward(W_l, b_l, feat[:M], s_light, g_light, dW_l, db_l, hidden_ptr=hidden)
            return (None,) * 9 + tuple(dW_sdf + db_sdf + dW_col + db_col + dW_l + db_l)
        g_feat_ptr, ld = None, 256
        if g_rgb is not None:
            g_x = core.color_backward(W_col, b_col, d, N, feat, s_rgb, g_rgb, dW_col, db_col)
            ed = 3 + 6 * core.desc.multires_d
            g_feat_ptr, ld = g_x.data_ptr() + 4 * ed, 288
        if n_light > 0 and g_light is not None:
            core.light_backward(W_l, b_l, feat, s_light, g_light, dW_l, db_l)
        if g_sdf is not None or g_feat_ptr is not None or (want_grad and g_grad is not None):
            core.sdf_backward(W_sdf, M, act, dW_sdf, db_sdf, rays=(o, d, z, N), g_sdf=g_sdf, g_feat=g_feat_ptr, g_feat_ld=ld,
                              g_grad=g_grad if want_grad else None)
        if g_rgb is not None:
            del g_x
        return (None,) * 9 + tuple(dW_sdf + db_sdf + dW_col + db_col + dW_l + db_l)


class _CompositeFn(Function):
    """Laplace density + alpha compositing + per-ray reductions; differentiable w.r.t. per-sample inputs and beta."""

    @staticmethod
    def forward(ctx, core, z, dnorm, beta_param, s_sdf, s_rgb, s_grad, s_light, want_normal, want_light):
        sg = s_grad if want_normal else None
        sl = s_light if want_light else None
        rgb, depth, wsum, normal, light = core.composite_forward(z, dnorm, beta_param, s_sdf, s_rgb, sg, sl)
        ctx.core, ctx.cfg = core, (want_normal, want_light)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(z, dnorm, beta_param, s_sdf, s_rgb, s_grad, s_light)
        normal = normal if want_normal else z.new_empty(0)
        light = light if want_light else z.new_empty(0)
        return rgb, depth, wsum, normal, light

    @staticmethod
    def backward(ctx, g_rgb, g_depth, g_wsum, g_normal, g_light):
        want_normal, want_light = ctx.cfg
        z, dnorm, beta_param, s_sdf, s_rgb, s_grad, s_light = ctx.saved_tensors
        o_sdf, o_rgb, o_grad, o_light, o_beta = ctx.core.composite_backward(
            z, dnorm, beta_param, s_sdf, s_rgb, s_grad if want_normal else None, s_light if want_light else None,
            g_rgb, g_depth, g_wsum, g_normal if want_normal else None, g_light if want_light else None)
        return (None, None, None, o_beta.reshape(beta_param.shape), o_sdf, o_rgb,
                o_grad if want_normal else None, o_light if want_light else None, None, None)


class _SdfPointsFn(Function):
    """x [M,3] -> (sdf [M], grad_x sdf [M,3]); differentiable (incl. second order) w.r.t. the SDF stack's weights."""

    @staticmethod
    def forward(ctx, core, pts, want_grad, n_sdf, *params):
        pts = pts.detach().contiguous().float()
        M = pts.shape[0]
        # sdf-only points with a backward (the bubble loss, network/__init__.py:196-201) also ride the tensor-core chain: its sdf + grad_x
        # table is the one that saves plane slots, the unused grad_x costs a reverse sweep on a few thousand points
        fused = bool(core.fused_sdf)
        act = core.sdf_saved_buffer(M, want_grad or fused)
        sdf, _, grad = core.sdf_forward(pts, want_grad=(want_grad or fused), save_act=act)
        ctx.core, ctx.cfg = core, (want_grad, n_sdf)
        ctx.fused = fused
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(pts, act, *params)
        if not want_grad:
            grad = pts.new_empty(0)
            ctx.mark_non_differentiable(grad)
        return sdf, grad

    @staticmethod
    def backward(ctx, g_sdf, g_grad):
        want_grad, n_sdf = ctx.cfg
        saved = ctx.saved_tensors
        pts, act = saved[:2]
        W, b = list(saved[2:2 + n_sdf]), list(saved[2 + n_sdf:2 + 2 * n_sdf])
        dW, db = _zeros_like_all([W, b])
        if g_sdf is not None or (want_grad and g_grad is not None):
            if ctx.fused:
                ctx.core.fused_backward(pts.shape[0], act, dW, db, pts=pts, g_sdf=g_sdf, g_grad=g_grad if want_grad else None)
            else:
                ctx.core.sdf_backward(W, pts.shape[0], act, dW, db, pts=pts, g_sdf=g_sdf, g_grad=g_grad if want_grad else None)
        return (None,) * 4 + tuple(dW + db)


def _sdf_params(model):
    W, b = model.implicit_network.effective()
    return [w.contiguous() for w in W], list(b)


def sdf_with_grad(model, x):
    """(sdf [M], None, grad [M,3]) with autograd support — ImplicitNetwork.gradient in training (mlp.py:107-118)."""
    core = model._ready_core()
    W, b = _sdf_params(model)
    sdf, grad = _SdfPointsFn.apply(core, x, True, len(W), *W, *b)
    return sdf, None, grad


def forward_train(model, core, input, predict_only=False):
    """I2SDFNetwork.forward with self.training == True (model/network/__init__.py:80-209)."""
    ov = getattr(model, "_tape_override", None) or {}
    # effective weights W = g v/|v| (mlp.py:71-72) once per step: differentiable inputs of the Functions below, and the source
    # of the packed device copies every kernel of this step reads (two launches)
    Ws, bs = model.effective_weights()
    core.pack(Ws, bs)
    model._packed_key = None        # an eval-mode forward after this one re-checks (and re-packs: the optimizer will have stepped)
    o, d, dnorm = core.rays(input["uv"], input["pose"], input["intrinsics"])
    R, dev = o.shape[0], o.device
    beta = model.density.beta
    if "z_all" in ov:                                   # test hook: reference z's
        z, z_eik = ov["z_all"].to(dev).contiguous(), ov["z_eik"].to(dev).reshape(-1).contiguous()
    else:
        tape = {k: ov[k] for k in ("jitter", "u_final", "extra_perm", "eik_idx") if k in ov} or model._draw_sampler_tape(R, dev)
        z, z_eik = core.sample(o, d, beta.detach(), tape, defer_sync=True,      # resolved below, behind the main pass
                                group=getattr(model, "convergence_group", None))
    n_sdf = model.implicit_network.num_layers - 1
    n_col = model.rendering_network.num_layers - 1
    n_light = (model.light_network.num_layers - 1) if model.use_light else 0
    W_sdf, b_sdf = Ws[:n_sdf], bs[:n_sdf]
    W_col, b_col = Ws[n_sdf:n_sdf + n_col], bs[n_sdf:n_sdf + n_col]
    W_l, b_l = Ws[n_sdf + n_col:], bs[n_sdf + n_col:]
    want_grad = bool(model.use_normal)                  # returns_grad (network/__init__.py:109) in training
    params = [w.contiguous() for w in W_sdf] + list(b_sdf) + [w.contiguous() for w in W_col] + list(b_col) + \
             [w.contiguous() for w in W_l] + list(b_l)
    # ---- eikonal / smoothness points (network/__init__.py:175-193).  Drawn here, before the main pass: on the fused
    # tensor-core path they are appended to the main-pass launches (same RNG calls in the same order as the reference:
    # nothing random happens in between)
    pts = None
    if not predict_only:
        bsph = model.scene_bounding_sphere
        eik_u = ov["eik_uniform"].to(dev) if "eik_uniform" in ov else torch.empty(R, 3, device=dev).uniform_(-bsph, bsph)
        near = o + z_eik[:, None] * d
        nbr_u = ov["nbr_uniform"].to(dev) if "nbr_uniform" in ov else torch.empty_like(near).uniform_(-0.005, 0.005)
        pts = torch.cat([eik_u, near, near + nbr_u], 0)
    ride = pts is not None and core.fused_main
    s_sdf, s_grad, s_rgb, s_light, x_grad = _PointsFn.apply(core, o, d, z, pts if ride else None, want_grad, n_sdf, n_col, n_light, *params)
    rgb, depth, wsum, normal, light = _CompositeFn.apply(core, z, dnorm, beta, s_sdf, s_rgb, s_grad, s_light,
                                                         want_grad and not predict_only, n_light > 0)
    res = {"rgb_values": rgb, "depth_values": depth, "weight_sum": wsum[:, None]}
    if n_light > 0:
        res["light_mask"] = light[:, None]
    # the sampler's round count is read back only now: the main pass is queued behind it, so the device stays busy while the
    # host waits, then replays the reference's one CPU-generator draw (randperm(n), ray_sampler.py:223)
    core.sampler_resolve()
    if predict_only:
        return res
    sdf_params = [w.contiguous() for w in W_sdf] + list(b_sdf)
    if ride:
        g = x_grad
    else:
        _, g = _SdfPointsFn.apply(core, pts, True, n_sdf, *sdf_params)
    res["grad_theta"] = g[:2 * R]
    nrm = F.normalize(g[R:], dim=1, eps=1e-6)
    res["diff_norm"] = torch.norm(nrm[:R] - nrm[R:], dim=1)
    # ---- bubble loss points (:196-201)
    if "pointcloud" in input:
        idx = int(ov["bubble_cam_idx"]) if "bubble_cam_idx" in ov else np.random.randint(0, R)
        sp = torch.cat([input["pointcloud"].to(dev).float(), o[idx][None]], 0)
        ssdf, _ = _SdfPointsFn.apply(core, sp, False, n_sdf, *sdf_params)
        res["surface_sdf"] = ssdf[:-1, None]
    if model.use_normal:
        res["normal_values"] = normal
    return res


# ----------------------------------------------------------------------------------------------------------------------
# I2SDFLoss on CUDA tensors: one launch for every term + the gradient w.r.t. every model output (csrc/loss.cu)
# ----------------------------------------------------------------------------------------------------------------------
_LOSS_PRED = ("rgb", "grad_theta", "diff_norm", "weight_sum", "depth", "normal", "surface_sdf", "light")     # differentiable inputs
_LOSS_AUX = ("rgb_gt", "mask_gt", "depth_gt", "depth_mask", "normal_gt", "normal_mask", "light_gt")


class _LossFn(Function):
    @staticmethod
    def forward(ctx, weights, aux, *preds):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        a = _lib.LossArgs()
        dev = preds[0].device
        keep = []

        def prep(t, dtype=torch.float32):
            t = t.detach()
            if t.dtype != dtype or not t.is_contiguous() or t.device != dev:
                t = t.to(device=dev, dtype=dtype).contiguous()
            keep.append(t)
            return t

        P = {k: (None if t is None else prep(t)) for k, t in zip(_LOSS_PRED, preds)}
        R = P["rgb"].shape[0]
        a.R, a.n_eik = R, (0 if P["grad_theta"] is None else P["grad_theta"].shape[0])
        a.n_bubble = 0 if P["surface_sdf"] is None else P["surface_sdf"].numel()
        for k in _LOSS_PRED:
            if P[k] is not None:
                rows = {"grad_theta": a.n_eik, "surface_sdf": a.n_bubble}.get(k, R)
                width = 3 if k in ("rgb", "grad_theta", "normal") else 1
                if P[k].numel() != rows * width:
                    raise _lib.I2SDFError(f"I2SDFLoss: {k} has {P[k].numel()} elements, expected {rows * width}")
                setattr(a, k, P[k].data_ptr())
        need = {"rgb": ("rgb_gt",), "weight_sum": ("mask_gt",), "depth": ("depth_gt", "depth_mask"), "normal": ("normal_gt", "normal_mask"),
                "light": ("light_gt",)}
        for k, deps in need.items():
            if P[k] is None:
                continue
            for dname in deps:
                t = aux.get(dname)
                if t is None:
                    raise _lib.I2SDFError(f"I2SDFLoss: ground truth for '{k}' is missing ({dname})")
                t = prep(t, torch.bool if dname.endswith("_mask") else torch.float32)
                if t.numel() != R * (3 if dname in ("rgb_gt", "normal_gt") else 1):
                    raise _lib.I2SDFError(f"I2SDFLoss: {dname} has the wrong size")
                setattr(a, dname, t.data_ptr())
        for k, v in weights.items():
            setattr(a, k, float(v))
        if aux.get("denom") is not None:            # sharded batch: divisors of the means (I2SDFLoss._shard_denominators)
            den = prep(aux["denom"])
            if den.numel() != 5:
                raise _lib.I2SDFError("I2SDFLoss: denom must hold 5 divisors (rays, eikonal rows, bubble points, depth count, normal count)")
            a.denom = den.data_ptr()
        # one flat buffer: terms[10] (padded to 16) + the gradient of every differentiable input that needs one
        want = [i for i, k in enumerate(_LOSS_PRED) if P[k] is not None and ctx.needs_input_grad[2 + i]]
        sizes = [(P[_LOSS_PRED[i]].numel() + 3) // 4 * 4 for i in want]
        flat = torch.empty(16 + sum(sizes), dtype=torch.float32, device=dev)
        a.terms = flat.data_ptr()
        grads, off = [None] * len(_LOSS_PRED), 16
        for i, n in zip(want, sizes):
            k = _LOSS_PRED[i]
            grads[i] = flat[off:off + P[k].numel()].view(preds[i].shape)
            setattr(a, "g_" + k, grads[i].data_ptr())
            off += n
        with torch.cuda.device(dev):
            _lib.check(lib.i2sdf_loss_forward(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "i2sdf_loss_forward")
        ctx.grads = grads
        ctx.flat = flat
        ctx.set_materialize_grads(False)
        del keep
        return flat[:10]

    @staticmethod
    def backward(ctx, g_terms):
        out = [None, None] + [None] * len(_LOSS_PRED)
        if g_terms is None:
            return tuple(out)
        live = [g for g in ctx.grads if g is not None]
        scaled = torch._foreach_mul(live, g_terms[0]) if live else []          # upstream of `loss` (terms[0]); the other terms are logged detached
        it = iter(scaled)
        for i, g in enumerate(ctx.grads):
            if g is not None:
                out[2 + i] = next(it)
        return tuple(out)


def fused_loss(sel, weights):
    """sel: dict with the _LOSS_PRED / _LOSS_AUX entries (None = term off) -> terms [10] (terms[0] = loss, differentiable)."""
    aux = {k: sel.get(k) for k in _LOSS_AUX + ("denom",)}
    return _LossFn.apply(weights, aux, *[sel.get(k) for k in _LOSS_PRED])


# ----------------------------------------------------------------------------------------------------------------------
# weight norm of every layer in one launch per direction (csrc/wnorm.cu)
# ----------------------------------------------------------------------------------------------------------------------
class _WeightNormAll(Function):
    """(g_0, v_0, g_1, v_1, ...) -> (W_0, W_1, ...) with W = g v / ||v||_row  (torch._weight_norm(v, g, 0) per layer)."""

    @staticmethod
    def _launch(gs, vs, norms, Ws=None, dWs=None, dgs=None, dvs=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        b = _lib.WnormBatch()
        b.n = len(gs)
        for i, (g, v, nrm) in enumerate(zip(gs, vs, norms)):
            j = b.jobs[i]
            j.g, j.v, j.norm, j.rows, j.cols = g.data_ptr(), v.data_ptr(), nrm.data_ptr(), v.shape[0], v.shape[1]
            if Ws is not None:
                j.W = Ws[i].data_ptr()
            if dWs is not None and dWs[i] is not None:
                j.dW, j.dg, j.dv = dWs[i].data_ptr(), dgs[i].data_ptr(), dvs[i].data_ptr()
        dev = vs[0].device
        with torch.cuda.device(dev):
            _lib.check(lib.i2sdf_weight_norm(C.byref(b), 0 if Ws is not None else 1, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                       "i2sdf_weight_norm")

    @staticmethod
    def forward(ctx, *gv):
        gs = [t.detach().contiguous() for t in gv[0::2]]
        vs = [t.detach().contiguous() for t in gv[1::2]]
        dev = vs[0].device
        rows = [v.shape[0] for v in vs]
        flat_n = torch.empty(sum((r + 3) // 4 * 4 for r in rows), device=dev)          # all norms in one allocation
        norms, off = [], 0
        for r in rows:
            norms.append(flat_n[off:off + r])
            off += (r + 3) // 4 * 4
        Ws = [torch.empty_like(v) for v in vs]
        _WeightNormAll._launch(gs, vs, norms, Ws=Ws)
        ctx.save_for_backward(*gs, *vs, flat_n)
        ctx.rows = rows
        # [g_0, v_0, g_1, v_1, ...]: (parameter, its persistent gradient view) where a GradBucket registered one
        ctx.grad_views = [(weakref.ref(t), getattr(t, "_i2sdf_grad_view", None)) for t in gv]
        ctx.set_materialize_grads(False)
        return tuple(Ws)

    @staticmethod
    def backward(ctx, *dWs):
        n = len(ctx.rows)
        saved = ctx.saved_tensors
        gs, vs, flat_n = saved[:n], saved[n:2 * n], saved[2 * n]
        norms, off = [], 0
        for r in ctx.rows:
            norms.append(flat_n[off:off + r])
            off += (r + 3) // 4 * 4
        dWs = [None if d is None else d.contiguous() for d in dWs]
        if all(d is None for d in dWs):
            return (None,) * (2 * n)
        gvw = ctx.grad_views
        direct = {}            # output index -> parameter whose .grad becomes the bucket view this backward wrote into

        def out_like(t, i):
            # Straight into the parameter's persistent gradient view (parallel.GradBucket) - and then the parameter's .grad is SET to that
            # view here and autograd gets no gradient for this input: handing the view back would make AccumulateGrad clone it (a view
            # is never "stolen": 27 device-to-device copies per step in the round-2 trace, tools/gpu_gaps.py).  Only while the parameter
            # has no .grad (with zero_grad(set_to_none=False) autograd must accumulate, so the ordinary path is taken).
            if gvw is not None and gvw[i][1] is not None:
                prm, w = gvw[i][0](), gvw[i][1]
                if prm is not None and prm.grad is None and w.shape == t.shape and w.device == t.device:
                    direct[i] = prm
                    return w
            return torch.empty_like(t)
        dgs = [None if d is None else out_like(g, 2 * i) for i, (d, g) in enumerate(zip(dWs, gs))]
        dvs = [None if d is None else out_like(v, 2 * i + 1) for i, (d, v) in enumerate(zip(dWs, vs))]
        _WeightNormAll._launch(gs, vs, norms, dWs=dWs, dgs=dgs, dvs=dvs)
        out = []
        for dg, dv in zip(dgs, dvs):
            out += [dg, dv]
        for i, prm in direct.items():
            prm.grad = out[i]
            out[i] = None
        return tuple(out)


def weight_norm_all(layers):
    """Effective weights of weight-normed nn.Linear layers (attributes weight_g [out,1], weight_v [out,in]) on a CUDA device."""
    args = []
    for lin in layers:
        args += [lin.weight_g, lin.weight_v]
    return list(_WeightNormAll.apply(*args))
