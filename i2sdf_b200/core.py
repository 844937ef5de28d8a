"""RenderCore: owns one i2sdf_handle (C ABI) on one GPU and the workspace tensors the calls need.

torch is used here for device memory and streams only; every tensor is handed to the library as a raw pointer.
"""
import ctypes as C
import math
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import Desc, check


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _f32(t: torch.Tensor, device) -> torch.Tensor:
    if t.dtype != torch.float32 or t.device != device or not t.is_contiguous():
        t = t.to(device=device, dtype=torch.float32).contiguous()
    return t


class RenderCore:
    def __init__(self, model_conf: dict, device: torch.device):
        if device.type != "cuda":
            raise _lib.I2SDFError("i2sdf_b200 runs on CUDA devices only (no CPU fallback); move the module to a B200")
        # (the cross-check backend lives in its own library: only a process that sets one of the I2SDF_SIMT* switches ever loads it)
        self.lib = _lib.load_check() if _lib.wants_check_backend() else _lib.load()
        self.device = device
        imp, ren, smp = model_conf["implicit_network"], model_conf["rendering_network"], model_conf["ray_sampler"]
        if ren.get("mode", "nerf") != "nerf":
            raise _lib.I2SDFError("rendering_network.mode must be 'nerf' (the only mode the shipped configs use)")
        if imp.get("embed_type") != "positional" or ren.get("embed_type") != "positional":
            raise _lib.I2SDFError("only embed_type 'positional' is supported")
        dims = list(imp["dims"])
        skip = list(imp.get("skip_in", ()))
        if len(skip) > 1:
            raise _lib.I2SDFError("at most one skip connection is supported")
        light = model_conf.get("light_network")
        fvs = model_conf["feature_vector_size"]
        if any(d != 256 for d in dims + list(ren["dims"])) or fvs != 256:
            raise _lib.I2SDFError("hidden width / feature_vector_size must be 256")
        d = Desc()
        d.abi_version = _lib.ABI_VERSION
        d.hidden, d.feature_size = 256, fvs
        d.n_sdf_layers, d.sdf_skip_layer, d.multires_x = len(dims) + 1, (skip[0] if skip else -1), imp["multires"]
        d.n_color_layers, d.multires_d = len(ren["dims"]) + 1, ren["multires"]
        d.n_light_layers = 2 if light else 0
        d.light_hidden = light["dims"][0] if light else 128
        d.n_samples, d.n_samples_eval, d.n_samples_extra = smp["N_samples"], smp["N_samples_eval"], smp["N_samples_extra"]
        d.beta_iters, d.max_total_iters = smp["beta_iters"], smp["max_total_iters"]
        d.near_ = float(smp["near"])
        d.far_ = 2.0 * float(model_conf.get("scene_bounding_sphere", 1.0))
        d.eps, d.add_tiny = float(smp["eps"]), float(smp.get("add_tiny", 0.0))
        d.beta_min = float(model_conf["density"].get("beta_min", 1e-4))
        # host tables with the reference's own torch ops so index arithmetic is bit-identical
        # (ray_sampler.py:30, :76, :188, :225)
        d.lemma2_coeff = float(1.0 / (4.0 * torch.log(torch.tensor(d.eps + 1.0))))
        u_up = torch.linspace(0.0, 1.0, steps=d.n_samples_eval)
        u_fin = torch.linspace(0.0, 1.0, steps=d.n_samples)
        t_in = torch.linspace(0.0, 1.0, steps=d.n_samples_eval)
        eidx = torch.stack([torch.linspace(0, d.n_samples_eval * (k + 1) - 1, d.n_samples_extra).long()
                            for k in range(d.max_total_iters)]).to(torch.int32).contiguous()
        self._tables = (u_up, u_fin, t_in, eidx)
        d.u_up = C.cast(u_up.data_ptr(), C.POINTER(C.c_float))
        d.u_final = C.cast(u_fin.data_ptr(), C.POINTER(C.c_float))
        d.t_init = C.cast(t_in.data_ptr(), C.POINTER(C.c_float))
        d.extra_idx = C.cast(eidx.data_ptr(), C.POINTER(C.c_int32))
        self.desc = d
        self.n_out = d.n_samples + 2 + d.n_samples_extra
        h = C.c_void_p()
        with torch.cuda.device(device):
            check(self.lib.i2sdf_create(C.byref(d), device.index or 0, C.byref(h)), "i2sdf_create")
        self.h = h
        self.n_layers = self.lib.i2sdf_num_layers(h)
        tcbits = self.lib.i2sdf_uses_tensor_cores(h)
        self.uses_tensor_cores = bool(tcbits & 1)
        self.uses_tensor_cores_main = bool(tcbits & 2)
        # what the training forwards save: plane slots + fused tensor-core backward, or fp32 pre-activations + layer-wise backward
        self.fused_main = bool(self.lib.i2sdf_saved_format(h, 0))
        self.fused_sdf = bool(self.lib.i2sdf_saved_format(h, 1))
        self._ws = None
        self._ws_rays = -1
        self._packed_refs = None
        self.rounds_log = []

    def close(self):
        if getattr(self, "h", None):
            self.lib.i2sdf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace(self, R: int) -> torch.Tensor:
        if self._ws is None or R > self._ws_rays:
            nbytes = self.lib.i2sdf_workspace_bytes(self.h, R, 0)
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            self._ws_rays = R
            self._ws_bytes = nbytes
        return self._ws

    # ---- measurement hook
    def profile(self, enable: bool):
        check(self.lib.i2sdf_profile_enable(self.h, int(enable)), "i2sdf_profile_enable")

    def profile_read(self):
        ms = (C.c_float * 7)()
        n = (C.c_int64 * 7)()
        check(self.lib.i2sdf_profile_read_n(self.h, 7, ms, n), "i2sdf_profile_read_n")
        kinds = ("sampler_sdf", "main_mlp", "sampler_rays", "misc", "backward_chain", "weight_grads", "light_head")
        return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(kinds)}

    # ---- plane slots (HBM format of the fused training path; see csrc/planes.cuh)
    def planes_pack(self, X: torch.Tensor, columns: int = 256) -> torch.Tensor:
        X = _f32(X, self.device)
        M, width = X.shape
        slot = torch.empty(self.lib.i2sdf_planes_slot_bytes(M, columns), dtype=torch.uint8, device=self.device)
        check(self.lib.i2sdf_planes_pack(self.h, _ptr(X), width, width, M, columns, _ptr(slot), self._stream()), "i2sdf_planes_pack")
        return slot

    def planes_unpack(self, slot: torch.Tensor, M: int, width: int = 256, columns: int = 256) -> torch.Tensor:
        X = torch.empty(M, width, device=self.device)
        check(self.lib.i2sdf_planes_unpack(self.h, _ptr(slot), columns, M, _ptr(X), width, width, self._stream()), "i2sdf_planes_unpack")
        return X

    def planes_wgrad(self, P_slots, X_slots, M: int, rows: int, cols: int, x_columns: int = 256, colsum: bool = False):
        """dW [rows, cols] = sum_t P_t^T X_t on the tensor cores (+ column sums of P_0)."""
        n = len(P_slots)
        dW = torch.zeros(rows, cols, device=self.device)
        cs = torch.zeros(256, device=self.device) if colsum else None
        Pp = (C.c_void_p * n)(*[t.data_ptr() for t in P_slots])
        Xp = (C.c_void_p * n)(*[t.data_ptr() for t in X_slots])
        check(self.lib.i2sdf_planes_wgrad(self.h, n, Pp, Xp, x_columns, M, _ptr(dW), cols, rows, cols, _ptr(cs), self._stream()),
              "i2sdf_planes_wgrad")
        return (dW, cs[:rows]) if colsum else dW

    # ---- weights
    def pack(self, weights: List[torch.Tensor], biases: List[torch.Tensor]):
        """weights[i]: effective [out,in] fp32 weight of layer i (SDF layers, colour layers, light layers)."""
        if len(weights) != self.n_layers or len(biases) != self.n_layers:
            raise _lib.I2SDFError(f"expected {self.n_layers} layers, got {len(weights)}")
        ws = [_f32(w.detach(), self.device) for w in weights]
        bs = [_f32(b.detach(), self.device) for b in biases]
        Wp = (C.c_void_p * self.n_layers)(*[w.data_ptr() for w in ws])
        Bp = (C.c_void_p * self.n_layers)(*[b.data_ptr() for b in bs])
        check(self.lib.i2sdf_pack_weights(self.h, Wp, Bp, self._stream()), "i2sdf_pack_weights")
        self._packed_refs = (ws, bs)     # keep alive until the async pack kernels ran

    # ---- entry points
    def rays(self, uv, pose, intr):
        uv, pose, intr = _f32(uv, self.device), _f32(pose, self.device), _f32(intr, self.device)
        B, P, _ = uv.shape
        o = torch.empty(B * P, 3, device=self.device)
        d = torch.empty(B * P, 3, device=self.device)
        dn = torch.empty(B * P, device=self.device)
        check(self.lib.i2sdf_rays(self.h, _ptr(uv), _ptr(pose), _ptr(intr), B, P, _ptr(o), _ptr(d), _ptr(dn), self._stream()), "i2sdf_rays")
        return o, d, dn

    def sdf_saved_buffer(self, M: int, want_grad: bool) -> torch.Tensor:
        """Buffer for sdf_forward(save_act=...): plane slots when the tensor-core chain serves the call, else fp32 [L-1,M,256]."""
        if self.fused_sdf and want_grad:
            return torch.empty(self.lib.i2sdf_sdf_saved_bytes(self.h, M), dtype=torch.uint8, device=self.device)
        return torch.empty(self.desc.n_sdf_layers - 1, M, 256, device=self.device)

    def sdf_forward(self, pts, want_feat=False, want_grad=False, save_act=None):
        pts = _f32(pts.detach(), self.device)
        M = pts.shape[0]
        sdf = torch.empty(M, device=self.device)
        feat = torch.empty(M, 256, device=self.device) if want_feat else None
        grad = torch.empty(M, 3, device=self.device) if want_grad else None
        if M == 0:
            return sdf, feat, grad
        ws = self.workspace(1)
        check(self.lib.i2sdf_sdf_forward(self.h, _ptr(pts), M, _ptr(sdf), _ptr(feat), _ptr(grad), _ptr(save_act),
                                         _ptr(ws), self._ws_bytes, self._stream()), "i2sdf_sdf_forward")
        return sdf, feat, grad

    def sdf_grid(self, gx, gy, gz, affine=None):
        """sdf of the regular grid meshgrid(gx, gy, gz) (numpy 'xy' order, utils/plots.py:445-446), points generated on the device."""
        gx, gy, gz = (_f32(torch.as_tensor(a), self.device) for a in (gx, gy, gz))
        aff = None if affine is None else _f32(torch.as_tensor(affine).reshape(12), self.device)
        n = gx.numel() * gy.numel() * gz.numel()
        out = torch.empty(n, device=self.device)
        with torch.cuda.device(self.device):
            check(self.lib.i2sdf_sdf_grid(self.h, _ptr(gx), _ptr(gy), _ptr(gz), gx.numel(), gy.numel(), gz.numel(), _ptr(aff), _ptr(out), self._stream()),
                  "i2sdf_sdf_grid")
        self._keep_grid = (gx, gy, gz, aff)
        return out

    def sample(self, o, d, beta_param, tape: Optional[Dict[str, torch.Tensor]] = None, want_info=False, defer_sync=False, group=None,
               staged=False):
        """ErrorBoundSampler.get_z_vals.  tape (training): jitter [R,128], u_final [R,64] fp32;
        extra_perm: callable n -> LongTensor[32] (drawn after one 8-byte D2H of n) or int tensor; eik_idx [R].

        defer_sync (training step): the reference draws randperm(n)[:32] on the CPU generator AFTER it knows n = 128 * rounds
        (ray_sampler.py:223), which costs a device->host round trip in the middle of the step.  Instead every candidate
        (n = 128 .. 128 * max_iters) is drawn from the SAME generator state, the device picks the row that applies, and
        sampler_resolve() - called by the caller once it has queued enough work behind the sampler - reads n back and
        replays the one draw that happened, so the host generator ends in exactly the reference's state.

        group (torch.distributed process group, world > 1): the rays of this call are one shard of a batch; the convergence
        word of every round is MAX-all-reduced over the group so that all shards run the rounds the whole batch would
        (the reference's test is batch-global, ray_sampler.py:151).  staged=True walks the same stage-by-stage entry point
        (i2sdf_sampler_step) without a group: same launches as the sharded path minus the exchange (tests)."""
        self.sampler_resolve()
        o, d = _f32(o, self.device), _f32(d, self.device)
        R = o.shape[0]
        tape = tape or {}
        ws = self.workspace(R)
        jit = _f32(tape["jitter"], self.device) if "jitter" in tape else None
        ufin = _f32(tape["u_final"], self.device) if "u_final" in tape else None
        st = self._stream()
        sharded = group is not None and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1
        if sharded or staged:
            import torch.distributed as dist
            args = (self.h, _ptr(o), _ptr(d), R, _ptr(beta_param), _ptr(jit), _ptr(ufin))
            check(self.lib.i2sdf_sampler_step(*args, 0, 0, _ptr(ws), self._ws_bytes, st), "i2sdf_sampler_step")
            off = self.lib.i2sdf_sampler_beta_max(self.h, R, _ptr(ws)) - ws.data_ptr()
            beta_max = ws[off:off + 4 * self.desc.max_total_iters].view(torch.float32)
            for k in range(self.desc.max_total_iters):
                check(self.lib.i2sdf_sampler_step(*args, 1, k, _ptr(ws), self._ws_bytes, st), "i2sdf_sampler_step")
                if sharded:
                    dist.all_reduce(beta_max[k:k + 1], op=dist.ReduceOp.MAX, group=group)   # 4 bytes, in stream order
                check(self.lib.i2sdf_sampler_step(*args, 2, k, _ptr(ws), self._ws_bytes, st), "i2sdf_sampler_step")
        else:
            check(self.lib.i2sdf_sampler_rounds(self.h, _ptr(o), _ptr(d), R, _ptr(beta_param), _ptr(jit), _ptr(ufin),
                                                _ptr(ws), self._ws_bytes, st), "i2sdf_sampler_rounds")
        info = torch.zeros(2, dtype=torch.int32, device=self.device)
        extra = None
        if defer_sync and callable(tape.get("extra_perm")):
            ep = tape["extra_perm"]
            state = torch.get_rng_state()
            cands = []
            for k in range(self.desc.max_total_iters):
                torch.set_rng_state(state)
                cands.append(ep(self.desc.n_samples_eval * (k + 1)).to(torch.int32))
            torch.set_rng_state(state)
            # staged through a persistent PINNED buffer: a pageable host->device copy is synchronous - the host would block
            # here until the device has run everything queued so far (the sampler rounds), which is the very wait this path removes
            if getattr(self, "_cand_host", None) is None:
                self._cand_host = torch.zeros(self.desc.max_total_iters, self.desc.n_samples_extra, dtype=torch.int32).pin_memory()
            torch.stack(cands, out=self._cand_host)
            table = torch.empty(self._cand_host.shape, dtype=torch.int32, device=self.device)
            table.copy_(self._cand_host, non_blocking=True)
            eik = tape["eik_idx"].to(device=self.device, dtype=torch.int32).contiguous() if "eik_idx" in tape else None
            if callable(tape.get("eik_idx_fn")):
                eik = tape["eik_idx_fn"]().to(device=self.device, dtype=torch.int32).contiguous()
            z = torch.empty(R, self.n_out, device=self.device)
            z_eik = torch.empty(R, device=self.device) if eik is not None else None
            check(self.lib.i2sdf_sampler_finalize_candidates(self.h, R, _ptr(beta_param), _ptr(table), _ptr(eik), _ptr(z), _ptr(z_eik),
                                                             _ptr(info), _ptr(ws), self._ws_bytes, st), "i2sdf_sampler_finalize_candidates")
            if getattr(self, "_info_host", None) is None:
                self._info_host = torch.zeros(2, dtype=torch.int32).pin_memory()
            self._info_host.copy_(info, non_blocking=True)
            if torch.cuda.is_current_stream_capturing():
                # CUDA-graph capture (i2sdf_b200/graph.py): the candidate upload and the round-count read-back above are graph nodes on
                # the two pinned buffers; the host side (drawing the candidates, replaying the generator) is graph_pre_replay / graph_resolve
                self._graph_ep = ep
                return (z, z_eik, info) if want_info else (z, z_eik)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._pending = (state, ep, ev)
            return (z, z_eik, info) if want_info else (z, z_eik)
        if "extra_perm" in tape:
            ep = tape["extra_perm"]
            if callable(ep):
                check(self.lib.i2sdf_sampler_info(self.h, R, _ptr(beta_param), _ptr(info), _ptr(ws), self._ws_bytes, st), "i2sdf_sampler_info")
                ep = ep(int(info[1].item()))
            extra = ep.to(device=self.device, dtype=torch.int32).contiguous()
        eik = tape["eik_idx"].to(device=self.device, dtype=torch.int32).contiguous() if "eik_idx" in tape else None
        if callable(tape.get("eik_idx_fn")):
            eik = tape["eik_idx_fn"]().to(device=self.device, dtype=torch.int32).contiguous()
        z = torch.empty(R, self.n_out, device=self.device)
        z_eik = torch.empty(R, device=self.device) if eik is not None else None
        check(self.lib.i2sdf_sampler_finalize(self.h, R, _ptr(beta_param), _ptr(extra), _ptr(eik), _ptr(z), _ptr(z_eik),
                                              _ptr(info), _ptr(ws), self._ws_bytes, st), "i2sdf_sampler_finalize")
        if want_info:
            return z, z_eik, info
        return z, z_eik

    def sampler_resolve(self):
        """Second half of sample(defer_sync=True): wait for the round count and replay the host generator draw that applied."""
        if torch.cuda.is_current_stream_capturing():
            return
        pend = getattr(self, "_pending", None)
        if pend is None:
            return
        self._pending = None
        state, ep, ev = pend
        ev.synchronize()
        torch.set_rng_state(state)
        ep(int(self._info_host[1]))
        self.rounds_log.append(int(self._info_host[0]))          # sampler rounds of that forward (bench.py reports them)
        del self.rounds_log[:-256]

    def graph_pre_replay(self):
        """Host half of a captured training forward, before the replay: draw this step's extra-sample candidates (one per possible
        round count, all from the same generator state, ray_sampler.py:223) into the pinned table the graph uploads."""
        ep = getattr(self, "_graph_ep", None)
        if ep is None:
            return
        state = torch.get_rng_state()
        cands = []
        for k in range(self.desc.max_total_iters):
            torch.set_rng_state(state)
            cands.append(ep(self.desc.n_samples_eval * (k + 1)).to(torch.int32))
        torch.set_rng_state(state)
        torch.stack(cands, out=self._cand_host)
        self._graph_state = state

    def graph_resolve(self):
        """... and after the replay has finished (the caller synchronised): replay the ONE draw that applied, so that the host generator
        ends where the reference's would (the round count sits in the pinned info buffer)."""
        state = getattr(self, "_graph_state", None)
        if state is None:
            return
        self._graph_state = None
        torch.set_rng_state(state)
        self._graph_ep(int(self._info_host[1]))
        self.rounds_log.append(int(self._info_host[0]))
        del self.rounds_log[:-256]

    def sampler_round_debug(self, z, sdf, beta_param, beta_in, upsample: bool, u_tape=None):
        dev = self.device
        z, sdf, beta_in = _f32(z, dev), _f32(sdf, dev), _f32(beta_in, dev)
        u_tape = None if u_tape is None else _f32(u_tape, dev)
        R, n = z.shape
        ns = self.desc.n_samples_eval if upsample else self.desc.n_samples
        out = dict(beta=torch.empty(R, device=dev), cdf=torch.empty(R, n, device=dev),
                   inds=torch.empty(R, ns, dtype=torch.int32, device=dev), samples=torch.empty(R, ns, device=dev))
        if upsample:
            out["z_merged"] = torch.empty(R, n + ns, device=dev)
            out["src"] = torch.empty(R, n + ns, dtype=torch.int32, device=dev)
        check(self.lib.i2sdf_sampler_round_debug(
            self.h, _ptr(z), _ptr(sdf), R, n, _ptr(beta_param), _ptr(beta_in), int(upsample),
            _ptr(u_tape), _ptr(out["beta"]), _ptr(out["cdf"]), _ptr(out["inds"]),
            _ptr(out["samples"]), _ptr(out.get("z_merged")), _ptr(out.get("src")), self._stream()), "i2sdf_sampler_round_debug")
        return out

    # ---- training split: forward pieces + backward
    def backward_workspace(self, M: int) -> torch.Tensor:
        nbytes = self.lib.i2sdf_backward_workspace_bytes(self.h, M)
        if getattr(self, "_bws", None) is None or self._bws.numel() < nbytes:
            self._bws = None
            self._bws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._bws

    @staticmethod
    def _ptr_array(tensors):
        return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def points_forward(self, o, d, z, want_grad, want_light, save=True, extra_pts=None):
        """Main-pass chain on the R*N ray samples (+ optional explicit points appended: fused tensor-core path only)."""
        dev = self.device
        o, d, z = _f32(o, dev), _f32(d, dev), _f32(z, dev)
        R, N = z.shape[0], z.shape[1] - 1
        E = 0 if extra_pts is None else extra_pts.shape[0]
        if E:
            extra_pts = _f32(extra_pts.detach(), dev)
        M = R * N + E
        fused = save and self.fused_main                          # plane slots: features stay inside the saved state
        out = dict(s_sdf=torch.empty(M, device=dev), s_rgb=torch.empty(M, 3, device=dev))
        # (with a light head the features are ALSO written as fp32: the head is a second pass over them, forward and backward)
        out["feat"] = None if (fused and not want_light) else torch.empty(M, 256, device=dev)
        out["s_grad"] = torch.empty(M, 3, device=dev) if (want_grad or fused) else None
        out["s_light"] = torch.empty(R * N, device=dev) if want_light else None
        out["act"] = torch.empty(self.lib.i2sdf_saved_bytes_points(self.h, M), dtype=torch.uint8, device=dev) if save else None
        out["fused"] = fused
        if E and not fused:
            raise _lib.I2SDFError("points_forward: appended points need the fused tensor-core path")
        ws = self.workspace(R)
        check(self.lib.i2sdf_points_forward_ex(self.h, _ptr(o), _ptr(d), _ptr(z), R, N, _ptr(extra_pts), E, _ptr(out["s_sdf"]),
                                               _ptr(out["s_grad"]), _ptr(out["s_rgb"]), _ptr(out["s_light"]), _ptr(out["feat"]),
                                               _ptr(out["act"]), _ptr(ws), self._ws_bytes, self._stream()), "i2sdf_points_forward_ex")
        self._keep_fwd = extra_pts
        return out

    def composite_forward(self, z, dnorm, beta_param, s_sdf, s_rgb, s_grad, s_light):
        dev = self.device
        R, N = z.shape[0], z.shape[1] - 1
        rgb, depth, wsum = torch.empty(R, 3, device=dev), torch.empty(R, device=dev), torch.empty(R, device=dev)
        normal = torch.empty(R, 3, device=dev) if s_grad is not None else None
        light = torch.empty(R, device=dev) if s_light is not None else None
        check(self.lib.i2sdf_composite_forward(self.h, _ptr(z), _ptr(dnorm), _ptr(s_sdf), _ptr(s_rgb), _ptr(s_grad), _ptr(s_light),
                                               _ptr(beta_param), R, N, _ptr(rgb), _ptr(depth), _ptr(wsum), _ptr(normal), _ptr(light),
                                               _ptr(None), self._stream()), "i2sdf_composite_forward")
        return rgb, depth, wsum, normal, light

    def composite_backward(self, z, dnorm, beta_param, s_sdf, s_rgb, s_grad, s_light, g_rgb, g_depth, g_wsum, g_normal, g_light):
        dev = self.device
        R, N = z.shape[0], z.shape[1] - 1
        M = R * N
        f = lambda t: None if t is None else _f32(t, dev)        # noqa: E731
        g_rgb, g_depth, g_wsum, g_normal, g_light = f(g_rgb), f(g_depth), f(g_wsum), f(g_normal), f(g_light)
        o_sdf, o_rgb = torch.empty(M, device=dev), torch.empty(M, 3, device=dev)
        o_grad = torch.empty(M, 3, device=dev) if s_grad is not None else None
        o_light = torch.empty(M, device=dev) if s_light is not None else None
        o_beta = torch.zeros(1, device=dev)
        check(self.lib.i2sdf_composite_backward(self.h, _ptr(z), _ptr(dnorm), _ptr(s_sdf), _ptr(s_rgb), _ptr(s_grad), _ptr(s_light),
                                                _ptr(beta_param), R, N, _ptr(g_rgb), _ptr(g_depth), _ptr(g_wsum),
                                                _ptr(g_normal if s_grad is not None else None), _ptr(g_light if s_light is not None else None),
                                                _ptr(o_sdf), _ptr(o_rgb), _ptr(o_grad), _ptr(o_light), _ptr(o_beta), self._stream()),
              "i2sdf_composite_backward")
        return o_sdf, o_rgb, o_grad, o_light, o_beta

    def color_backward(self, Ws, bs, dirs, ns, feat, s_rgb, g_rgb, dWs, dbs):
        dev = self.device
        M = feat.shape[0]
        g_x = torch.empty(M, 288, device=dev)
        bws = self.backward_workspace(M)
        g_rgb = _f32(g_rgb, dev)
        check(self.lib.i2sdf_color_backward(self.h, self._ptr_array(Ws), self._ptr_array(bs), _ptr(dirs), ns, _ptr(feat), _ptr(s_rgb),
                                            _ptr(g_rgb), M, self._ptr_array(dWs), self._ptr_array(dbs), _ptr(g_x), _ptr(bws),
                                            bws.numel(), self._stream()), "i2sdf_color_backward")
        return g_x

    def light_backward(self, Ws, bs, feat, s_light, g_light, dWs, dbs, hidden_ptr=0):
        """hidden_ptr: device address of the head's saved hidden pre-activations (fused path), 0 = recompute."""
        M = feat.shape[0]
        bws = self.backward_workspace(M)
        g_light = _f32(g_light, self.device)
        check(self.lib.i2sdf_light_backward(self.h, self._ptr_array(Ws), self._ptr_array(bs), _ptr(feat), C.c_void_p(hidden_ptr), _ptr(s_light), _ptr(g_light), M,
                                            self._ptr_array(dWs), self._ptr_array(dbs), _ptr(bws), bws.numel(), self._stream()),
              "i2sdf_light_backward")

    def sdf_backward(self, Ws, M, act, dWs, dbs, pts=None, rays=None, g_sdf=None, g_feat=None, g_feat_ld=256, g_grad=None):
        """rays = (o, d, z [R,zstride], ns).  g_feat may be a (data_ptr, ld) view into a wider buffer."""
        dev = self.device
        bws = self.backward_workspace(M)
        g_sdf = None if g_sdf is None else _f32(g_sdf, dev)
        g_grad = None if g_grad is None else _f32(g_grad, dev)
        if rays is not None:
            o, d, z, ns = rays
            po, pd, pz, zs, pp = _ptr(o), _ptr(d), _ptr(z), z.shape[1], _ptr(None)
        else:
            po = pd = pz = _ptr(None)
            zs, ns, pp = 0, 1, _ptr(pts)
        gf = C.c_void_p(0)
        if g_feat is not None:
            gf = C.c_void_p(g_feat if isinstance(g_feat, int) else g_feat.data_ptr())
        check(self.lib.i2sdf_sdf_backward(self.h, self._ptr_array(Ws), pp, po, pd, pz, zs, ns, M, _ptr(act), _ptr(g_sdf), gf, g_feat_ld,
                                          _ptr(g_grad), self._ptr_array(dWs), self._ptr_array(dbs), _ptr(bws), bws.numel(), self._stream()),
              "i2sdf_sdf_backward")
        self._keep = (g_sdf, g_grad)

    def fused_backward(self, M, saved, dW_sdf, db_sdf, pts=None, rays=None, s_rgb=None, g_sdf=None, g_grad=None, g_rgb=None,
                       dW_col=None, db_col=None, m_up=None, g_grad_tail=None):
        """Tensor-core backward on plane slots (i2sdf_fused_backward).  rays = (o, d, z [R,zstride], ns); with rays AND pts
        the first R*ns points are ray samples and pts are the points appended after them."""
        dev = self.device
        bws = self.backward_workspace(M)
        f = lambda t: None if t is None else _f32(t, dev)        # noqa: E731
        g_sdf, g_grad, g_rgb, g_grad_tail = f(g_sdf), f(g_grad), f(g_rgb), f(g_grad_tail)
        m_rays = 0
        if rays is not None:
            o, d, z, ns = rays
            po, pd, pz, zs, pp = _ptr(o), _ptr(d), _ptr(z), z.shape[1], _ptr(pts)
            m_rays = z.shape[0] * ns
        else:
            po = pd = pz = _ptr(None)
            zs, ns, pp = 0, 1, _ptr(pts)
        null = C.POINTER(C.c_void_p)()
        # m_up: the upstream arrays cover the first m_up points; the appended points' grad_x upstream comes separately (g_grad_tail)
        check(self.lib.i2sdf_fused_backward_ex(self.h, pp, po, pd, pz, zs, ns, M, m_rays, _ptr(saved), _ptr(s_rgb), _ptr(g_sdf), _ptr(g_grad), _ptr(g_rgb),
                                               M if m_up is None else m_up, _ptr(g_grad_tail),
                                               self._ptr_array(dW_sdf), self._ptr_array(db_sdf),
                                               self._ptr_array(dW_col) if dW_col else null, self._ptr_array(db_col) if db_col else null,
                                               _ptr(bws), bws.numel(), self._stream()), "i2sdf_fused_backward_ex")
        self._keep = (g_sdf, g_grad, g_rgb, g_grad_tail)

    def render(self, o, d, dnorm, z, beta_param, want_normal=True, want_light=False, per_sample=False, save=None):
        """Main pass + compositing.  z [R,N+1].  Returns dict of per-ray tensors (+ per-sample ones if asked)."""
        dev = self.device
        o, d, dnorm, z = _f32(o, dev), _f32(d, dev), _f32(dnorm, dev), _f32(z, dev)
        R, N1 = z.shape
        N = N1 - 1
        ws = self.workspace(R)
        out = dict(rgb=torch.empty(R, 3, device=dev), depth=torch.empty(R, device=dev), weight_sum=torch.empty(R, device=dev))
        if want_normal:
            out["normal"] = torch.empty(R, 3, device=dev)
        if want_light:
            out["light"] = torch.empty(R, device=dev)
        ps = {}
        if per_sample:
            ps = dict(s_sdf=torch.empty(R * N, device=dev), s_rgb=torch.empty(R * N, 3, device=dev), s_w=torch.empty(R * N, device=dev))
            if want_normal:
                ps["s_grad"] = torch.empty(R * N, 3, device=dev)
            if want_light:
                ps["s_light"] = torch.empty(R * N, device=dev)
        save_bytes = 0 if save is None else save.numel() * save.element_size()
        check(self.lib.i2sdf_render_forward(
            self.h, _ptr(o), _ptr(d), _ptr(dnorm), _ptr(z), R, N, _ptr(beta_param),
            _ptr(out["rgb"]), _ptr(out["depth"]), _ptr(out["weight_sum"]), _ptr(out.get("normal")), _ptr(out.get("light")),
            _ptr(ps.get("s_sdf")), _ptr(ps.get("s_grad")), _ptr(ps.get("s_rgb")), _ptr(ps.get("s_w")), _ptr(ps.get("s_light")),
            _ptr(save), save_bytes, _ptr(ws), self._ws_bytes, self._stream()), "i2sdf_render_forward")
        out.update(ps)
        return out
