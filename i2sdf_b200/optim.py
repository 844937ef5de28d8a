"""Adam for the reconstruction trainer as ONE kernel launch per step.

    opt = i2sdf_b200.optim.Adam(model.get_param_groups(lr), eps=1e-15)        # instead of torch.optim.Adam(...)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma)                # unchanged (reads / writes param_groups[i]['lr'])

Reference: model/trainer/recon.py:201-207 builds torch.optim.Adam(self.model.get_param_groups(lr), eps=1e-15) +
ExponentialLR.  Same update rule, same state layout (`step`, `exp_avg`, `exp_avg_sq`: state dicts interchange with
torch.optim.Adam), no weight decay / amsgrad / maximize (the reference uses none).  CUDA fp32 parameters only: like the rest of
the package there is no CPU path.
"""
import ctypes as C
import math
import os

import torch

from . import _lib


_FAST_OFF = os.environ.get("I2SDF_ADAM_FAST", "1") == "0"        # measurement switch (tools/step_times.py)


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False, maximize=False):
        if weight_decay != 0.0 or amsgrad or maximize:
            raise _lib.I2SDFError("i2sdf_b200.optim.Adam: weight_decay / amsgrad / maximize are not supported (the reference uses none)")
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0, amsgrad=False, maximize=False))
        self._fresh_steps = {}
        self._batches = {}
        self._graph_scalars = {}
        self._plans = {}

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        if not hasattr(self, "_fresh_steps"):
            self._fresh_steps = {}
        if not hasattr(self, "_batches"):
            self._batches = {}
        if not hasattr(self, "_graph_scalars"):
            self._graph_scalars = {}
        if not hasattr(self, "_plans"):
            self._plans = {}
        for group in self.param_groups:
            if self._fast_step(group, lib):
                continue
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            if len(ps) < len(group["params"]):
                # a parameter that sits this step out keeps its own count (torch.optim.Adam's per-parameter `step`): detach it from
                # the counter object it may share with the parameters that do step
                for p in group["params"]:
                    st = self.state.get(p) if p.grad is None else None
                    if st and "step" in st:
                        st["step"] = st["step"].clone()
            beta1, beta2 = group["betas"]
            # the step counters are host tensors (torch.optim.Adam's state layout); parameters created together share ONE
            # tensor object so that a step costs one host increment, not one per parameter
            seen = set()
            for p in ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse:
                    raise _lib.I2SDFError("i2sdf_b200.optim.Adam needs contiguous fp32 CUDA parameters with dense gradients")
                st = self.state[p]
                if len(st) == 0:
                    shared = self._fresh_steps.get(id(group))
                    if shared is None or float(shared) != 0.0:          # (a counter that already advanced is not handed to new parameters)
                        shared = self._fresh_steps[id(group)] = torch.tensor(0.0, dtype=torch.float32)
                    st["step"] = shared
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                sp = st["step"]
                if id(sp) not in seen:
                    seen.add(id(sp))
                    sp += 1
            # one launch per distinct step count: normally ONE for the whole group; a parameter that got its first gradient later
            # than the others (a loss term switched on mid-training) runs on its own count, as in torch.optim.Adam
            by_step = {}
            for p in ps:
                by_step.setdefault(float(self.state[p]["step"]), []).append(p)
            dev = ps[0].device
            with torch.cuda.device(dev):
                stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                for t, plist in by_step.items():
                    grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in plist]
                    # the job tables (pointers of parameter / gradient / moments) are rebuilt only when a pointer changed: with
                    # parallel.GradBucket every pointer is stable from step to step and a step costs two float stores per table
                    key = (id(group), tuple(p.data_ptr() for p in plist), tuple(g.data_ptr() for g in grads),
                           tuple(self.state[p]["exp_avg"].data_ptr() for p in plist), tuple(self.state[p]["exp_avg_sq"].data_ptr() for p in plist))
                    cached = self._batches.get(id(group))
                    if cached is None or cached[0] != key:
                        batches = []
                        for lo in range(0, len(plist), _lib.ADAM_MAX_JOBS):
                            b = _lib.AdamBatch()
                            chunk = plist[lo:lo + _lib.ADAM_MAX_JOBS]
                            b.n = len(chunk)
                            for i, p in enumerate(chunk):
                                st, j = self.state[p], b.jobs[i]
                                j.param, j.grad, j.numel = p.data_ptr(), grads[lo + i].data_ptr(), p.numel()
                                j.exp_avg, j.exp_avg_sq = st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                            batches.append(b)
                        cached = (key, batches)
                        if len(by_step) == 1:
                            self._batches[id(group)] = cached
                    # steady state from the next step on (_fast_step): every parameter of the group has a gradient and one shared counter
                    if len(by_step) == 1 and len(plist) == len(group["params"]) and not torch.cuda.is_current_stream_capturing():
                        st0 = self.state[plist[0]]
                        for p in plist:                  # equal counts (one launch): share ONE counter object again, e.g. after load_state_dict
                            self.state[p]["step"] = st0["step"]
                        self._plans[id(group)] = dict(params=list(group["params"]), state0=st0, step=st0["step"], batches=cached[1], dev=dev,
                                                      jobs=[b.jobs[i] for b in cached[1] for i in range(b.n)])
                    capturing = torch.cuda.is_current_stream_capturing()
                    scal = None
                    if capturing:
                        # CUDA-graph capture (i2sdf_b200/graph.py): the step's two scalars come from device memory, refreshed on every replay by
                        # a captured copy from a pinned host buffer that prepare_replay() fills
                        gi = self.param_groups.index(group)
                        if gi not in self._graph_scalars:
                            self._graph_scalars[gi] = (torch.zeros(2, dtype=torch.float32).pin_memory(), torch.zeros(2, dtype=torch.float32, device=dev), plist[0])
                        host, scal, _ = self._graph_scalars[gi]
                        host[0], host[1] = group["lr"] / (1.0 - beta1 ** t), math.sqrt(1.0 - beta2 ** t)
                        scal.copy_(host, non_blocking=True)
                    for b in cached[1]:
                        b.beta1, b.beta2, b.eps = beta1, beta2, group["eps"]
                        b.one_minus_beta1, b.one_minus_beta2 = 1.0 - beta1, 1.0 - beta2
                        b.step_size = group["lr"] / (1.0 - beta1 ** t)
                        b.bias_correction2_sqrt = math.sqrt(1.0 - beta2 ** t)
                        if scal is not None:
                            _lib.check(lib.i2sdf_adam_step_dev(C.byref(b), C.c_void_p(scal.data_ptr()), stream), "i2sdf_adam_step_dev")
                        else:
                            _lib.check(lib.i2sdf_adam_step(C.byref(b), stream), "i2sdf_adam_step")
                    del grads
        return loss

    def _fast_step(self, group, lib) -> bool:
        """Steady-state step of one group: same parameter objects as when the job tables were built, all with dense contiguous gradients,
        one shared step counter, unchanged optimizer state (load_state_dict / add_param_group / a parameter without a gradient fall back to
        the general path, which rebuilds the plan).  Host cost: one pass over the parameters to refresh the gradient pointers (autograd
        allocates new gradient tensors every step unless a parallel.GradBucket pins them) - the general path's per-step dictionary
        lookups, float() conversions and pointer-key tuples cost 0.5-0.8 ms per step for 44 tensors (tools/host_profile.py), which the
        GPU spends idle at the step boundary once the rest of the step is as short as it is."""
        plan = self._plans.get(id(group))
        if plan is None or _FAST_OFF:
            return False

        def drop():
            # back to the general path: the shared job tables carry this path's gradient pointers, so its pointer-keyed cache is stale too
            self._plans.pop(id(group), None)
            self._batches.pop(id(group), None)
            return False
        if torch.cuda.is_current_stream_capturing():
            return drop()
        params, pp = group["params"], plan["params"]
        if len(params) != len(pp):
            return drop()
        st0 = self.state.get(params[0])
        if st0 is None or st0 is not plan["state0"] or st0.get("step") is not plan["step"]:
            return drop()
        keep = []                                    # the gradient tensors stay referenced until the launch is queued
        for p, q, j in zip(params, pp, plan["jobs"]):
            g = p.grad
            if p is not q or g is None or g.is_sparse or not g.is_contiguous() or j.param != p.data_ptr():
                return drop()
            gp = g.data_ptr()
            if j.grad != gp:
                j.grad = gp
            keep.append(g)
        sp = plan["step"]
        sp += 1
        t = float(sp)
        beta1, beta2 = group["betas"]
        dev = plan["dev"]
        with torch.cuda.device(dev):
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for b in plan["batches"]:
                b.beta1, b.beta2, b.eps = beta1, beta2, group["eps"]
                b.one_minus_beta1, b.one_minus_beta2 = 1.0 - beta1, 1.0 - beta2
                b.step_size = group["lr"] / (1.0 - beta1 ** t)
                b.bias_correction2_sqrt = math.sqrt(1.0 - beta2 ** t)
                _lib.check(lib.i2sdf_adam_step(C.byref(b), stream), "i2sdf_adam_step")
        del keep
        return True

    def prepare_replay(self):
        """Before replaying a CUDA graph that captured step(): advance the step counters and refresh the scalars the captured kernels read."""
        for gi, (host, _, p0) in self._graph_scalars.items():
            group = self.param_groups[gi]
            beta1, beta2 = group["betas"]
            seen = set()
            for p in group["params"]:
                st = self.state.get(p)
                if st and id(st["step"]) not in seen:
                    seen.add(id(st["step"]))
                    st["step"] += 1
            t = float(self.state[p0]["step"])
            host[0], host[1] = group["lr"] / (1.0 - beta1 ** t), math.sqrt(1.0 - beta2 ** t)

    def state_dict(self):
        """torch.optim.Adam's layout with ONE step tensor PER PARAMETER: inside this optimizer the parameters of a group share a
        step tensor object (one host increment per step); saved as is, `torch.save` would keep that aliasing and a
        torch.optim.Adam loading the checkpoint would then advance the shared counter once per parameter."""
        sd = super().state_dict()
        sd["state"] = {k: ({**v, "step": v["step"].clone()} if "step" in v else dict(v)) for k, v in sd["state"].items()}
        return sd
