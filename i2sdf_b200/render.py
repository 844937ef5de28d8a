"""Whole-image rendering driver (next-tier row of SURVEY.md §8(f), rank 2).

Replaces, for the B200 core, the chunk loop the reference's eval systems run around the model:
`utils.split_input` / `utils.merge_output` (utils/__init__.py:35-84) inside `VolumeRenderSystem.test_step`
(model/eval/recon.py:161-203) and the pixel grid of `PlotDataset.get_uv` (dataset/eval_dataset.py:144-148).
The uv grid is generated on the device, chunks go through `I2SDFNetwork.forward` (eval) and the per-ray outputs are
written straight into pre-allocated image buffers: no index_select copies, no list-of-dicts merge.

Note the reference's batch-global sampler convergence (ray_sampler.py:151): results depend on the chunk size in the
reference too; `split_n_pixels` here means exactly what the yaml's `train.split_n_pixels` means there.
"""
from typing import Dict, Sequence

import torch


def pixel_grid(img_res: Sequence[int], device=None) -> torch.Tensor:
    """uv [H*W, 2] = (x, y) per pixel in row-major order — identical to PlotDataset.get_uv (mgrid, flip, reshape)."""
    H, W = int(img_res[0]), int(img_res[1])
    ys = torch.arange(H, device=device, dtype=torch.float32).repeat_interleave(W)
    xs = torch.arange(W, device=device, dtype=torch.float32).repeat(H)
    return torch.stack([xs, ys], dim=-1)


@torch.no_grad()
def render_image(model, pose: torch.Tensor, intrinsics: torch.Tensor, img_res: Sequence[int], split_n_pixels: int = 65536,
                 predict_only: bool = False, group=None, assemble: bool = True) -> Dict[str, torch.Tensor]:
    """Render one full view.  pose, intrinsics: [4,4] (or [1,4,4]).  Returns {key: [H*W, C]}: merge_output's dict, except that the
    entries it leaves 1-D (depth_values [H*W], utils/__init__.py:76-78) come back as [H*W, 1] - its callers reshape them to
    (1, H*W, 1) right away (model/eval/recon.py:181-182).

    group (torch.distributed process group): the view's chunks are dealt to the ranks round-robin — chunk i of the single-GPU
    loop goes to rank i % world, so every chunk is the SAME set of rays as without sharding and (the sampler's convergence test
    being global per forward call, ray_sampler.py:151) renders to the same bits.  No collective on the data path (SURVEY.md
    §8(e): inference has none); with `assemble` the ranks' disjoint pieces are summed into the full image on every rank by one
    all-reduce per output key at the end, otherwise each rank returns its own chunks and zeros elsewhere."""
    if model.training:
        raise RuntimeError("render_image is an inference driver: call model.eval() first")
    dev = model.density.beta.device
    uv = pixel_grid(img_res, dev)
    total = uv.shape[0]
    pose = pose.reshape(1, 4, 4).to(dev).float()
    intrinsics = intrinsics.reshape(1, 4, 4).to(dev).float()
    if group is not None:
        return _render_image_sharded(model, uv, pose, intrinsics, split_n_pixels, predict_only, group, assemble)
    out: Dict[str, torch.Tensor] = {}
    for lo in range(0, total, split_n_pixels):
        hi = min(lo + split_n_pixels, total)
        res = model({"uv": uv[None, lo:hi], "pose": pose, "intrinsics": intrinsics}, predict_only=predict_only)
        for k, v in res.items():
            v2 = v.reshape(hi - lo, -1)
            if k not in out:
                out[k] = torch.empty(total, v2.shape[1], device=dev, dtype=v2.dtype)
            out[k][lo:hi] = v2
    return out


def _render_image_sharded(model, uv, pose, intrinsics, split_n_pixels, predict_only, group, assemble):
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev, total = uv.device, uv.shape[0]
    out: Dict[str, torch.Tensor] = {}
    # The ranks render DIFFERENT chunks, possibly different numbers of them: a sampler that MAX-all-reduces its convergence word
    # (parallel.use_global_convergence(eval_forwards=True)) would deadlock on an odd chunk count and, with equal counts, couple the
    # convergence of unrelated chunks.  Each chunk is its own forward call, as in the single-GPU loop: switch the collective off here.
    saved_group = getattr(model, "convergence_group_eval", None)
    model.convergence_group_eval = None
    try:
        for i, lo in enumerate(range(0, total, split_n_pixels)):
            if i % world != rank:
                continue
            hi = min(lo + split_n_pixels, total)
            res = model({"uv": uv[None, lo:hi], "pose": pose, "intrinsics": intrinsics}, predict_only=predict_only)
            for k, v in res.items():
                v2 = v.reshape(hi - lo, -1)
                if k not in out:
                    out[k] = torch.zeros(total, v2.shape[1], device=dev, dtype=v2.dtype)
                out[k][lo:hi] = v2
    finally:
        model.convergence_group_eval = saved_group
    if assemble and world > 1:
        # a rank with no chunk (more ranks than chunks) learns keys / widths / dtypes from rank 0, which always owns chunk 0
        meta = [[(k, v.shape[1], str(v.dtype).replace("torch.", "")) for k, v in sorted(out.items())] if rank == 0 else None]
        dist.broadcast_object_list(meta, src=dist.get_global_rank(group, 0), group=group)
        for k, width, dt in meta[0]:
            if k not in out:
                out[k] = torch.zeros(total, width, device=dev, dtype=getattr(torch, dt))
            dist.all_reduce(out[k], op=dist.ReduceOp.SUM, group=group)       # disjoint pieces + zeros: the sum is exact
    return out
