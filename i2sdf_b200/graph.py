"""A whole training step as ONE CUDA graph: forward (sampler rounds, main pass, compositing) + I2SDFLoss + backward + gradient
all-reduce + Adam + weight re-pack, replayed with one launch per step.

Why: a 1024-ray step is ~60 dependent kernel launches issued by ~2.7 ms of Python (autograd tape, ctypes calls).  One B200 hides that
behind its 5.5 ms of kernels, but a shard of 1024 / 8 rays (BASELINE.json configs[4], strong scaling) needs only ~1.1 ms of device time
and the step becomes host-bound (bench.py `strong_scaling.host_enqueue_ms_per_step`).  Replaying a captured graph removes the host from
the step; what stays on the host is what the reference does on the host:
  * the CPU-generator draw of the extra-sample permutation (ray_sampler.py:223) - candidates for every possible round count are drawn
    before the replay into a pinned table the graph uploads, and the one draw that applied is replayed afterwards, so the host generator
    ends where the reference's would (`strict_rng`; off: the generator advances by the largest candidate, no wait on the device);
  * Adam's bias-correction scalars (two floats per step through a pinned buffer).
Everything else - the device RNG draws (torch's graph-safe Philox offsets), the NCCL all-reduce of the gradient bucket, the kernels of
libi2sdf_b200.so (launched on the capturing stream through the same C ABI) - is inside the graph.

    step = GraphedTrainStep(model, loss_fn, optimizer, model_input, ground_truth, current_step=0)
    for model_input, ground_truth in batches:            # same shapes as the example batch
        loss = step(model_input, ground_truth)           # a device tensor; .item() only when it is logged

Re-capture (build a new GraphedTrainStep) when something baked in changes: the batch shape, the learning rate (ExponentialLR steps once
per epoch in the reference), or a loss term switching on (smooth_iter / bubble iterations).
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist


class GraphedTrainStep:
    def __init__(self, model, loss_fn, optimizer, model_input: Dict[str, torch.Tensor], ground_truth: Dict[str, torch.Tensor],
                 current_step: int = 0, group: Optional["dist.ProcessGroup"] = None, bucket=None, warmup: int = 3, strict_rng: bool = True):
        from .optim import Adam
        from .parallel import GradBucket
        if not model.training:
            raise RuntimeError("GraphedTrainStep captures a TRAINING step: call model.train() first")
        if not isinstance(optimizer, Adam):
            raise TypeError("GraphedTrainStep needs i2sdf_b200.optim.Adam (its step reads the per-step scalars from device memory under capture)")
        self.model, self.loss_fn, self.opt, self.group = model, loss_fn, optimizer, group
        self.current_step, self.strict_rng = current_step, strict_rng
        self.dev = model.density.beta.device
        self.world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        self.bucket = bucket if bucket is not None else GradBucket(model.parameters())
        self.inp = {k: v.to(self.dev).clone() for k, v in model_input.items()}
        self.gt = {k: v.to(self.dev).clone() for k, v in ground_truth.items()}
        self.core = model._ready_core(defer_pack=True)
        # warm-up on a side stream (allocator, lazily created pinned buffers, per-device kernel attributes), as torch.cuda.graph asks
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(max(1, warmup)):
                self._eager_step()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.core.sampler_resolve()
        self.graph = torch.cuda.CUDAGraph()
        self.bucket.zero()
        self.core.graph_pre_replay()                     # (no-op before the first capture: sets nothing)
        with torch.cuda.graph(self.graph):
            out = self.model(self.inp)
            self.terms = self.loss_fn(out, self.gt, self.current_step)
            self.loss = self.terms["loss"]
            self.loss.backward()
            self.bucket.allreduce(self.group)
            self.opt.step()
        self._pending = False
        # the capture itself did not run anything: the optimizer's host step counter moved by one without an update -> take it back
        for st in {id(v["step"]): v["step"] for v in self.opt.state.values() if "step" in v}.values():
            st -= 1

    def _eager_step(self):
        out = self.model(self.inp)
        loss = self.loss_fn(out, self.gt, self.current_step)["loss"]
        self.bucket.zero()
        loss.backward()
        self.bucket.allreduce(self.group)
        self.opt.step()

    def finish(self):
        """Wait for the last replay and bring the host generator where the reference's would be (strict_rng)."""
        if self._pending:
            self._done.synchronize()
            self._pending = False
            if self.strict_rng:
                self.core.graph_resolve()

    def __call__(self, model_input: Dict[str, torch.Tensor], ground_truth: Dict[str, torch.Tensor]) -> torch.Tensor:
        if self.strict_rng:
            self.finish()                                # the previous step's round count decides which host draw happened
        for k, v in model_input.items():
            self.inp[k].copy_(v, non_blocking=True)
        for k, v in ground_truth.items():
            self.gt[k].copy_(v, non_blocking=True)
        self.core.graph_pre_replay()
        self.opt.prepare_replay()
        self.graph.replay()
        self._done = torch.cuda.Event()
        self._done.record(torch.cuda.current_stream(self.dev))
        self._pending = True
        if not self.strict_rng:
            # no wait: the host generator advances by the draw for the largest round count (deterministic, not the reference's sequence)
            st = getattr(self.core, "_graph_state", None)
            if st is not None:
                torch.set_rng_state(st)
                self.core._graph_ep(self.core.desc.n_samples_eval * self.core.desc.max_total_iters)
                self.core._graph_state = None
        self.model._packed_key = None                    # an eval forward after this re-packs (the optimizer moved the weights)
        return self.loss
