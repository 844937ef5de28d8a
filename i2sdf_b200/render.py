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
                 predict_only: bool = False) -> Dict[str, torch.Tensor]:
    """Render one full view.  pose, intrinsics: [4,4] (or [1,4,4]).  Returns {key: [H*W, C]} like merge_output."""
    if model.training:
        raise RuntimeError("render_image is an inference driver: call model.eval() first")
    dev = model.density.beta.device
    uv = pixel_grid(img_res, dev)
    total = uv.shape[0]
    pose = pose.reshape(1, 4, 4).to(dev).float()
    intrinsics = intrinsics.reshape(1, 4, 4).to(dev).float()
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
