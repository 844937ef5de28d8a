"""Synthetic inputs of the measured workloads: rays and training targets.

Neutral ground shared by both arms of bench.py, smoke() and the tests: nothing here touches a kernel or the oracle.
Rays: pose = I with t = (0, 0, -1.5) (a camera outside the geometric-init sphere of radius 0.6, looking down +z),
fx = fy = 300, cx = 160, cy = 120, uv ~ U([0, 320] x [0, 240]); eval layout uv [1, R, 2] (dataset/eval_dataset.py:150-168),
training layout uv [R, 1, 2] with per-ray pose / intrinsics (dataset/train_dataset.py:169-192)."""
from typing import Dict

import torch


def synthetic_rays(R: int, seed: int = 1, train_layout: bool = False) -> Dict[str, torch.Tensor]:
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


def make_train_gt(R: int, seed: int, light: bool = False, bubble_points: int = 0) -> Dict[str, torch.Tensor]:
    """Targets with the keys / shapes of the collated training batch (dataset/train_dataset.py:176-192): rgb, depth + mask,
    normal + mask (, light_mask; `pointcloud` [Bb, 3] on the init sphere's surface rides in the INPUT dict, see bubble_input)."""
    g = torch.Generator().manual_seed(seed)
    gt = {
        "rgb": torch.rand(R, 3, generator=g),
        "depth": torch.rand(R, generator=g) * 2 + 0.5,
        "depth_mask": torch.ones(R, dtype=torch.bool),
        "normal": torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1),
        "normal_mask": torch.ones(R, dtype=torch.bool),
    }
    if light:
        gt["light_mask"] = (torch.rand(R, 1, generator=g) > 0.9).float()
    return gt


def bubble_points(n: int, seed: int = 11, radius: float = 0.6) -> torch.Tensor:
    """Stand-in for the trainer's sampled bubble point cloud (model/trainer/recon.py:156-170): n points near the surface."""
    g = torch.Generator().manual_seed(seed)
    p = torch.nn.functional.normalize(torch.randn(n, 3, generator=g), dim=-1)
    return p * (radius + 0.02 * torch.randn(n, 1, generator=g))
