"""Multi-GPU plumbing for the per-ray path: one process per GPU, rays sharded, weights replicated.

The reference has no distributed code (`strategy=None` at main_recon.py:112); rays are independent
units, so the B200 design is plain data parallelism:
  * inference: shard rays / pixel chunks, no collective;
  * training: every rank renders its shard, then ONE all-reduce of the flat gradient (800 955 fp32 = 3.2 MB for
    config/synthetic.yml) over NCCL/NVLink per step.
These helpers only touch torch tensors and torch.distributed, so they run under gloo on CPU for the tests.
(A torch DistributedDataParallel wrapper around I2SDFNetwork also works: the custom autograd Functions return
ordinary parameter gradients, so DDP's bucket hooks fire as usual.)

Parity caveats of sharding, all inherited from the reference's batch-global semantics: the
sampler's convergence test is global over the rays of ONE forward call (ray_sampler.py:151), so a shard may stop
up-sampling one round earlier than the full batch would (`use_global_convergence` restores the batch-global test);
masked-mean losses average per shard (`use_global_loss_means` restores the batch-global means).
"""
from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced partition of n items: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rays(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    """Split a reference-style input dict across ranks along the ray axis.

    Training layout (dataset/train_dataset.py:169-192): uv [R,1,2], pose [R,4,4], intrinsics [R,4,4] -> split dim 0.
    Eval layout (dataset/eval_dataset.py:150-168): uv [1,P,2], pose [1,4,4], intrinsics [1,4,4] -> split uv dim 1."""
    uv = batch["uv"]
    out = dict(batch)
    if uv.shape[0] == 1 and uv.shape[1] > 1:
        lo, hi = shard_bounds(uv.shape[1], rank, world)
        out["uv"] = uv[:, lo:hi].contiguous()
        return out
    lo, hi = shard_bounds(uv.shape[0], rank, world)
    for k, v in batch.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == uv.shape[0] and k != "pointcloud":
            out[k] = v[lo:hi].contiguous()
    return out


def seed_rank(seed: int, rank: int, device: Optional[torch.device] = None) -> int:
    """Per-rank RNG streams for the training-mode draws of a sharded step (SURVEY.md §8(e) caveat 3).

    The reference seeds ONE process (`pl.seed_everything(args.seed)`, main_recon.py:63) and then draws, per step, the sampler's
    jitter / inverse-CDF uniforms (ray_sampler.py:39,190: device generator), the extra-sample permutation (:223: CPU generator),
    the eikonal pick (:233) and the eikonal / neighbour points (network/__init__.py:178,186).  Ranks that shared that seed would
    draw the SAME jitter and eikonal points for different rays; every rank therefore gets its own stream, a fixed function of
    (seed, rank): rank 0 keeps the reference's seed, so a 1-GPU run reproduces the unsharded run draw for draw.  Seeds the CPU
    generator and `device`'s generator (default: the current CUDA device if there is one); returns the seed used."""
    s = int(seed) + 0x9E3779B1 * int(rank)          # distinct, reproducible, rank 0 == seed
    s &= (1 << 63) - 1
    torch.manual_seed(s)                             # CPU generator (+ all CUDA generators, as torch.manual_seed does)
    if device is not None and torch.device(device).type == "cuda":
        with torch.cuda.device(device):
            torch.cuda.manual_seed(s)
    return s


def use_global_convergence(model, group: Optional[dist.ProcessGroup] = None, enable: bool = True, eval_forwards: bool = False):
    """Make the error-bounded sampler's convergence test global over all ranks' rays (strict sharding parity).

    The reference stops up-sampling when `beta.max() <= beta0` over the rays of ONE forward call (ray_sampler.py:151), so a
    shard on its own may stop a round earlier than the full batch would and then draws different z's.  With this switch
    every round MAX-all-reduces its 4-byte convergence word over `group` (NCCL, in stream order, no host sync): the
    sharded step / render then equals the single-GPU one of the whole batch bit for bit (tests/test_multigpu.py).

    Training forwards only, unless `eval_forwards` is set: an eval-mode forward that enters a collective deadlocks as soon
    as the ranks do not ALL call it the same number of times - rank-0-only validation / plotting, or `render_image(group=)`
    dealing an odd number of chunks.  With `eval_forwards=True` every rank must issue the same eval forwards (e.g. one
    sharded batch rendered by all ranks together); `render_image(group=)` switches it off around its own chunk loop."""
    if enable and not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("use_global_convergence needs an initialised torch.distributed process group")
    g = (group if group is not None else dist.group.WORLD) if enable else None
    model.convergence_group = g
    model.convergence_group_eval = g if eval_forwards else None
    return model


def use_global_loss_means(loss_fn, group: Optional[dist.ProcessGroup] = None, enable: bool = True):
    """Make I2SDFLoss divide by the counts of the WHOLE batch (strict sharding parity of the loss and its gradients).

    The reference's loss terms are means over all rays of the batch, two of them over masked subsets
    (model/network/__init__.py:320-329).  A mean of per-shard means equals the batch mean only if every shard has the same
    count — false for the depth / normal masks in general and for ragged shards.  With this switch every rank divides its
    sums by (global count) / world (one 40-byte SUM all-reduce per step, in stream order): averaging the ranks' losses, and
    the gradients as `allreduce_gradients(average=True)` does, then gives exactly the single-GPU values."""
    if enable and not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("use_global_loss_means needs an initialised torch.distributed process group")
    loss_fn.means_group = (group if group is not None else dist.group.WORLD) if enable else None
    return loss_fn


class GradBucket:
    """ONE persistent flat gradient buffer for all parameters: the all-reduce runs on it directly, the optimizer reads it.

    Round 1 flattened with `torch.cat` before and un-flattened with `_foreach_copy_` after every all-reduce (two extra passes
    over 3.2 MB + their launches, SCALE_r01: +0.28 ms per step at 8 GPUs against ~0.03 ms of wire time).  Here every
    parameter's `.grad` IS a view of the flat buffer: the weight-norm backward (the last kernel of `loss.backward()`, which
    produces 99.5 % of the gradient bytes) writes dg / dv straight into its views (`autograd._WeightNormAll`), the few
    tensors autograd produced elsewhere (biases, density.beta) are copied in by one multi-tensor kernel, parameters without
    a gradient contribute zeros - so the flat size is identical on every rank whatever terms a shard happened to have.

        bucket = GradBucket(model.parameters())
        loss.backward(); bucket.allreduce(group); opt.step(); bucket.zero()          # zero() instead of opt.zero_grad()
    """

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradBucket: no parameters")
        ref = self.params[0]
        offs, n = [], 0
        for p in self.params:
            if p.dtype != ref.dtype or p.device != ref.device:
                raise ValueError("GradBucket: parameters must share dtype and device")
            offs.append(n)
            n += (p.numel() + 3) // 4 * 4                   # 16-byte aligned views
        self.flat = torch.zeros(n, dtype=ref.dtype, device=ref.device)
        self.views = [self.flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, self.params)]
        from . import autograd as _ag
        for p, v in zip(self.params, self.views):
            _ag.register_grad_view(p, v)

    def gather(self) -> int:
        """After backward: make every parameter's .grad the bucket view (copying what was produced elsewhere, zeroing what has
        no gradient).  Returns the number of tensors that had to be copied."""
        src, dst, zero = [], [], []
        for p, v in zip(self.params, self.views):
            g = p.grad
            if g is None:
                zero.append(v)
            elif g.data_ptr() != v.data_ptr() or g.stride() != v.stride():
                src.append(g)
                dst.append(v)
            p.grad = v
        if zero:
            torch._foreach_zero_(zero)
        if src:
            torch._foreach_copy_(dst, src)
        return len(src)

    def allreduce(self, group: Optional[dist.ProcessGroup] = None, average: bool = True) -> int:
        self.gather()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if average and dist.get_backend(group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=group)       # the division happens inside NCCL
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
                if average:
                    self.flat /= dist.get_world_size(group)
        return self.flat.numel()

    def zero(self):
        """Replacement for optimizer.zero_grad(set_to_none=True): the next backward writes fresh values into the views."""
        for p in self.params:
            p.grad = None


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                        average: bool = True) -> int:
    """One all-reduce of all gradients as a single flat bucket; returns the number of elements reduced.

    Stateless variant (flatten / un-flatten every call); `GradBucket` is the persistent one.  Parameters without a gradient
    contribute ZEROS (and receive the reduced value): the flat buffer has the same size on every rank even when a term was
    absent on one shard - dropping them per rank, as round 1 did, made NCCL hang or mis-assign silently."""
    ps = [p for p in params if p.requires_grad]
    if not ps:
        return 0
    world = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
    if world == 1:
        return sum(p.numel() for p in ps if p.grad is not None)
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in ps]
    flat = torch.cat([g.reshape(-1) for g in grads])                       # one kernel
    if average and dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)           # the division happens inside NCCL
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat /= world
    views = [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)]
    torch._foreach_copy_(grads, views)                                     # un-flatten in a couple of multi-tensor kernels
    for p, g in zip(ps, grads):
        if p.grad is None:
            p.grad = g
    return flat.numel()
