"""Device-resident ray feed for the reconstruction trainer.

The reference feeds its trainer one PIXEL per dataset item: `ReconDataset.__getitem__` (dataset/train_dataset.py:169-192)
returns (global pixel index, image index, {uv [1,2], intrinsics [4,4], pose [4,4]}, {rgb [3], ...}) and `collate_fn` (:194-209)
stacks `batch_size` of them with a 4-worker DataLoader(shuffle=True) (model/trainer/recon.py).  At B200 kernel speeds (6 ms
per 1024-ray step) assembling 1024 dicts per step in Python workers is the bottleneck, so the tables live on the GPU and one
batch is a handful of gathers:

    feed = RayFeed.from_dataset(train_dataset, device)            # or RayFeed(uv=..., intrinsics_all=..., ...)
    for indices, img_indices, model_input, ground_truth in feed.batches(batch_size):   # one epoch, shuffled without replacement
        ...

Same tuple, keys, shapes and dtypes as the reference's collated batch (uv [R,1,2], intrinsics [R,4,4], pose [R,4,4], rgb [R,3],
mask / light_mask [R,1], depth [R], depth_mask [R] bool, normal [R,3], normal_mask [R] bool).  `BubblePDF` is the device-side
state of the bubble step (model/trainer/recon.py:142-170: `update_pdf`, `sample_bubble`).  Only tensor ops: the feed is plumbing
around the kernels, runs wherever its tensors live (the CPU tests drive it on the CPU), and draws its randomness from the
generator of its device like the reference's DataLoader sampler (CPU `torch.randperm`) and `torch.multinomial` do.
"""
from typing import Dict, Iterator, Optional, Tuple

import torch


class RayFeed:
    _GT_IMAGES = ("rgb_images", "mask_images", "lightmask_images", "depth_images", "depth_masks", "normal_images", "normal_masks")

    def __init__(self, uv: torch.Tensor, intrinsics_all: torch.Tensor, pose_all: torch.Tensor, rgb_images: torch.Tensor,
                 mask_images: Optional[torch.Tensor] = None, lightmask_images: Optional[torch.Tensor] = None,
                 depth_images: Optional[torch.Tensor] = None, depth_masks: Optional[torch.Tensor] = None,
                 normal_images: Optional[torch.Tensor] = None, normal_masks: Optional[torch.Tensor] = None,
                 device: Optional[torch.device] = None):
        dev = torch.device(device) if device is not None else rgb_images.device
        mv = lambda t: None if t is None else t.to(dev).contiguous()          # noqa: E731
        self.device = dev
        self.uv = mv(uv)                                   # [HW,2]  (u, v) pixel coordinates as the dataset builds them (:71-74)
        self.intrinsics_all, self.pose_all = mv(intrinsics_all), mv(pose_all)     # [N,4,4]
        self.rgb_images = mv(rgb_images)                   # [N,HW,3]
        self.mask_images, self.lightmask_images = mv(mask_images), mv(lightmask_images)         # [N,HW,1]
        self.depth_images, self.depth_masks = mv(depth_images), mv(depth_masks)                 # [N,HW], [N,HW] bool
        self.normal_images, self.normal_masks = mv(normal_images), mv(normal_masks)             # [N,HW,3], [N,HW] bool
        self.n_images, self.total_pixels = self.rgb_images.shape[0], self.rgb_images.shape[1]
        if self.uv.shape[0] != self.total_pixels:
            raise ValueError("uv and rgb_images disagree about the number of pixels per image")
        if (self.depth_images is None) != (self.depth_masks is None) or (self.normal_images is None) != (self.normal_masks is None):
            raise ValueError("depth / normal images need their masks")

    @classmethod
    def from_dataset(cls, ds, device) -> "RayFeed":
        """Build from a reference `ReconDataset` instance (reads the attributes its constructor fills, :42-167)."""
        kw = dict(uv=ds.uv, intrinsics_all=ds.intrinsics_all, pose_all=ds.pose_all, rgb_images=ds.rgb_images)
        if getattr(ds, "use_mask", False):
            kw["mask_images"] = ds.mask_images
        if getattr(ds, "use_lightmask", False):
            kw["lightmask_images"] = ds.lightmask_images
        if getattr(ds, "use_depth", False) or getattr(ds, "use_bubble", False):
            kw["depth_images"], kw["depth_masks"] = ds.depth_images, ds.depth_masks
        if getattr(ds, "use_normal", False):
            kw["normal_images"], kw["normal_masks"] = ds.normal_images, ds.normal_masks
        return cls(device=device, **kw)

    def __len__(self) -> int:
        return self.n_images * self.total_pixels           # ReconDataset.__len__ (:166-167)

    def gather(self, indices: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
        """The collated batch of the dataset items `indices` (global pixel indices, int64 [R])."""
        indices = indices.to(device=self.device, dtype=torch.long)
        img = torch.div(indices, self.total_pixels, rounding_mode="floor")
        pix = indices - img * self.total_pixels
        sample = {"uv": self.uv[pix].unsqueeze(1), "intrinsics": self.intrinsics_all[img], "pose": self.pose_all[img]}
        gt = {"rgb": self.rgb_images[img, pix]}
        if self.mask_images is not None:
            gt["mask"] = self.mask_images[img, pix]
        if self.lightmask_images is not None:
            gt["light_mask"] = self.lightmask_images[img, pix]
        if self.depth_images is not None:
            gt["depth"] = self.depth_images[img, pix]
            gt["depth_mask"] = self.depth_masks[img, pix]
        if self.normal_images is not None:
            gt["normal"] = self.normal_images[img, pix]
            gt["normal_mask"] = self.normal_masks[img, pix]
        return indices, img, sample, gt

    def batches(self, batch_size: int, shuffle: bool = True, drop_last: bool = False, generator: Optional[torch.Generator] = None,
                group=None) -> Iterator[Tuple[torch.Tensor, torch.Tensor, dict, dict]]:
        """One epoch: every pixel of every image exactly once (DataLoader(shuffle=True) semantics).

        group (a torch.distributed process group, rays sharded over its ranks): `batch_size` stays the GLOBAL batch of the
        reference's trainer; every rank walks the same shuffled epoch — the permutation's seed is drawn on rank 0 and broadcast
        (8 bytes per epoch) — and yields its contiguous share of each global batch (parallel.shard_bounds), so the ranks'
        batches are disjoint, cover the global batch, and the union over an epoch is every pixel exactly once."""
        n = len(self)
        rank, world = 0, 1
        if group is not None:
            import torch.distributed as dist
            from .parallel import shard_bounds
            rank, world = dist.get_rank(group), dist.get_world_size(group)
            if shuffle:
                seed = torch.randint(0, 2 ** 62, (1,), generator=generator if (generator is None or generator.device.type == "cpu") else None)
                seed = seed.to(self.device)
                dist.broadcast(seed, src=dist.get_global_rank(group, 0), group=group)
                generator = torch.Generator(device=self.device).manual_seed(int(seed.item()))
        order = torch.randperm(n, device=self.device, generator=generator) if shuffle else torch.arange(n, device=self.device)
        for lo in range(0, n, batch_size):
            idx = order[lo:lo + batch_size]
            if drop_last and idx.numel() < batch_size:
                return
            if world > 1:
                a, b = shard_bounds(idx.numel(), rank, world)
                idx = idx[a:b]
            yield self.gather(idx)

    def random_batch(self, batch_size: int, generator: Optional[torch.Generator] = None):
        """A batch of uniformly drawn pixels (with replacement across calls): for benchmarks and smoke runs."""
        return self.gather(torch.randint(len(self), (batch_size,), device=self.device, generator=generator))


class BubblePDF:
    """Importance-sampling state of the bubble step (model/trainer/recon.py:142-170), on the device of the point cloud.

    pointcloud [P,3]; pointlinks [N*HW] int64: global pixel index -> point index or -1 (dataset/train_dataset.py:109-128)."""

    def __init__(self, pointcloud: torch.Tensor, pointlinks: torch.Tensor, pdf_prune: float = 0.0, pdf_max: Optional[float] = None,
                 uniform: bool = False, device: Optional[torch.device] = None):
        dev = torch.device(device) if device is not None else pointcloud.device
        self.pointcloud = pointcloud.to(dev).contiguous()
        self.pointlinks = pointlinks.to(dev).contiguous()
        self.pdf = torch.zeros(self.pointcloud.shape[0], device=dev)
        self.sample_count = torch.zeros(self.pointcloud.shape[0], dtype=torch.long, device=dev)
        self.pdf_prune, self.pdf_max, self.uniform = pdf_prune, pdf_max, uniform

    @torch.no_grad()
    def update_pdf(self, value: torch.Tensor, idx: torch.Tensor, group=None) -> None:
        """pdf[point of pixel idx] = clamp / prune(value)   (recon.py:142-152); pixels without a point are skipped.

        group: rays sharded over the ranks of a process group (SURVEY.md §8(e) caveat 4).  The reference's PDF is one buffer
        updated with the errors of the WHOLE batch; a rank on its own would only see its shard.  With `group` the (value, idx)
        pairs of all ranks are all-gathered (ragged shards padded) and applied in rank order, so every rank holds the PDF the
        single-process trainer would hold."""
        if group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(group)
            if world > 1:
                dev = self.pdf.device
                value, idx = value.to(dev).reshape(-1).float(), idx.to(dev).reshape(-1).long()
                n = torch.tensor([value.numel()], device=dev, dtype=torch.long)
                sizes = [torch.zeros_like(n) for _ in range(world)]
                dist.all_gather(sizes, n, group=group)
                sizes = [int(t.item()) for t in sizes]
                cap = max(sizes)
                pv, pi = value.new_zeros(cap), idx.new_full((cap,), -1)
                pv[:value.numel()], pi[:idx.numel()] = value, idx
                gv, gi = [torch.empty_like(pv) for _ in range(world)], [torch.empty_like(pi) for _ in range(world)]
                dist.all_gather(gv, pv, group=group)
                dist.all_gather(gi, pi, group=group)
                value = torch.cat([t[:k] for t, k in zip(gv, sizes)])
                idx = torch.cat([t[:k] for t, k in zip(gi, sizes)])
        value = value.to(self.pdf.device).clone()
        if self.pdf_max is not None:
            value = value.clamp(max=self.pdf_max)
        value[value < self.pdf_prune] = 0
        link = self.pointlinks[idx.to(self.pdf.device)]
        mask = link != -1
        self.pdf[link[mask]] = value[mask]

    @torch.no_grad()
    def sample_bubble(self, batch_size: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """[batch_size,3] surface points ~ pdf without replacement (recon.py:154-168); uniform variant: a random subset."""
        if self.uniform:
            sel = torch.randperm(self.pointcloud.shape[0], device=self.pointcloud.device, generator=generator)[:batch_size]
            return self.pointcloud[sel]
        cand = torch.where(self.pdf > 0)[0]
        if cand.numel() >= (1 << 24):
            raise RuntimeError("PDF capacity exceeds torch.multinomial's limit of 2^24 categories")    # the reference exits here
        pick = torch.multinomial(self.pdf[cand], batch_size, replacement=False, generator=generator)
        self.sample_count[cand[pick]] += 1
        return self.pointcloud[cand[pick]]
