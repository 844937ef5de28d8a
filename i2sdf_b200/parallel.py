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


def use_global_convergence(model, group: Optional[dist.ProcessGroup] = None, enable: bool = True):
    """Make the error-bounded sampler's convergence test global over all ranks' rays (strict sharding parity).

    The reference stops up-sampling when `beta.max() <= beta0` over the rays of ONE forward call (ray_sampler.py:151), so a
    shard on its own may stop a round earlier than the full batch would and then draws different z's.  With this switch
    every round MAX-all-reduces its 4-byte convergence word over `group` (NCCL, in stream order, no host sync): the
    sharded render then equals the single-GPU render of the whole batch bit for bit (tests/test_multigpu.py)."""
    if enable and not (dist.is_available() and dist.is_initialized()):
        raise RuntimeError("use_global_convergence needs an initialised torch.distributed process group")
    model.convergence_group = (group if group is not None else dist.group.WORLD) if enable else None
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


def allreduce_gradients(params: Iterable[torch.nn.Parameter], group: Optional[dist.ProcessGroup] = None,
                        average: bool = True) -> int:
    """One all-reduce of all gradients as a single flat bucket; returns the number of elements reduced."""
    ps = [p for p in params if p.grad is not None]
    if not ps:
        return 0
    grads = [p.grad for p in ps]
    flat = torch.cat([g.reshape(-1) for g in grads])                       # one kernel
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        world = dist.get_world_size(group)
        if average and dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)       # the division happens inside NCCL
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat /= world
    views = [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)]
    torch._foreach_copy_(grads, views)                                     # un-flatten in a couple of multi-tensor kernels
    return flat.numel()
