"""SDF evaluation of regular grids for mesh extraction and validation plots.

Replaces the loops `for pnts in torch.split(grid_points, ...): z.append(model.implicit_network(pnts)[:, 0])` of the
reference (model/eval/recon.py:46-56 low-resolution pass, :87-90 full pass over the PCA-aligned grid; utils/plots.py:188-199
validation surface) and their grid builders `get_grid_uniform` / `get_grid` (utils/plots.py:440-489).  The axis arrays are
built exactly as the reference builds them (numpy float64 linspace / arange, cast to float32); the points themselves are
never materialised: the kernel derives each point from its index, runs the sdf-only tensor-core chain and writes 4 bytes.
The marching-cubes step stays with the caller (skimage, CPU), which receives the same [ny, nx, nz]-ordered volume.
"""
from typing import Optional, Sequence

import numpy as np
import torch


def grid_axes_uniform(resolution: int, grid_boundary: Sequence[float] = (-2.0, 2.0)):
    """The axes of get_grid_uniform (utils/plots.py:440-451)."""
    x = np.linspace(grid_boundary[0], grid_boundary[1], resolution)
    return x, x, x


def grid_axes_from_points(points: torch.Tensor, resolution: int, input_min=None, input_max=None, eps: float = 0.1):
    """The axes of get_grid (utils/plots.py:453-489): `resolution` samples along the shortest bounding-box axis, the same step on the others."""
    if input_min is None or input_max is None:
        input_min = torch.min(points, dim=0)[0].squeeze().cpu().numpy()
        input_max = torch.max(points, dim=0)[0].squeeze().cpu().numpy()
    bb = input_max - input_min
    sa = int(np.argmin(bb))
    short = np.linspace(input_min[sa] - eps, input_max[sa] + eps, resolution)
    length = np.max(short) - np.min(short)
    step = length / (short.shape[0] - 1)
    axes = []
    for a in range(3):
        axes.append(short if a == sa else np.arange(input_min[a] - eps, input_max[a] + step + eps, step))
    return tuple(axes), length, sa


def grid_points(x, y, z) -> torch.Tensor:
    """The explicit point list the reference builds (np.meshgrid + vstack, utils/plots.py:445-446) - for tests and small grids."""
    xx, yy, zz = np.meshgrid(x, y, z)
    return torch.tensor(np.vstack([xx.ravel(), yy.ravel(), zz.ravel()]).T, dtype=torch.float)


@torch.no_grad()
def sdf_grid(model, x, y, z, rotation: Optional[torch.Tensor] = None, translation: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sdf [ny * nx * nz] of the grid meshgrid(x, y, z), in the reference's point order, on the model's device.

    rotation [3, 3] / translation [3]: evaluate at  R p + t  instead of p - eval/recon.py:80-84 maps the aligned grid back with
    `vecs^T p + s_mean`, i.e. rotation = vecs.T, translation = s_mean."""
    core = model._ready_core()
    aff = None
    if rotation is not None or translation is not None:
        R = torch.eye(3) if rotation is None else torch.as_tensor(rotation).float().cpu()
        t = torch.zeros(3) if translation is None else torch.as_tensor(translation).float().cpu()
        aff = torch.cat([R.reshape(9), t.reshape(3)])
    f32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64)).float()         # noqa: E731  (float64 axis -> float32, as torch.tensor(..., dtype=float) does)
    return core.sdf_grid(f32(x), f32(y), f32(z), aff)


def sdf_volume(model, x, y, z, **kw) -> np.ndarray:
    """The array the reference hands to marching cubes: z.reshape(ny, nx, nz).transpose([1, 0, 2])  (utils/plots.py:204-206)."""
    v = sdf_grid(model, x, y, z, **kw).cpu().numpy().astype(np.float32)
    return v.reshape(len(y), len(x), len(z)).transpose([1, 0, 2])
