"""Host logic of the device-resident ray feed (i2sdf_b200/feed.py) against a literal restatement of the reference's per-pixel
dataset items + collate (dataset/train_dataset.py:169-209) and of its bubble PDF (model/trainer/recon.py:142-168)."""
import numpy as np
import torch

from i2sdf_b200.feed import BubblePDF, RayFeed


def _tables(n_img=3, h=4, w=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    hw = h * w
    vv, uu = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    uv = torch.stack([uu, vv], 0).float().reshape(2, -1).t().contiguous()           # (u, v) per pixel, row-major (:71-74)
    t = dict(uv=uv, intrinsics_all=torch.rand(n_img, 4, 4, generator=g), pose_all=torch.rand(n_img, 4, 4, generator=g),
             rgb_images=torch.rand(n_img, hw, 3, generator=g), mask_images=(torch.rand(n_img, hw, 1, generator=g) > 0.5).float(),
             lightmask_images=(torch.rand(n_img, hw, 1, generator=g) > 0.8).float(), depth_images=torch.rand(n_img, hw, generator=g) * 6,
             depth_masks=torch.rand(n_img, hw, generator=g) > 0.3,
             normal_images=torch.nn.functional.normalize(torch.randn(n_img, hw, 3, generator=g), dim=-1),
             normal_masks=torch.rand(n_img, hw, generator=g) > 0.2)
    return t, hw


def _reference_item(t, hw, idx):
    """ReconDataset.__getitem__ with every use_* flag on."""
    pidx, img = idx % hw, idx // hw
    sample = {"uv": t["uv"][pidx].unsqueeze(0), "intrinsics": t["intrinsics_all"][img], "pose": t["pose_all"][img]}
    gt = {"rgb": t["rgb_images"][img][pidx], "mask": t["mask_images"][img][pidx], "light_mask": t["lightmask_images"][img][pidx],
          "depth": t["depth_images"][img][pidx], "depth_mask": t["depth_masks"][img][pidx],
          "normal": t["normal_images"][img][pidx], "normal_mask": t["normal_masks"][img][pidx]}
    return idx, img, sample, gt


def _collate(items):
    cols = list(zip(*items))
    out = []
    for entry in cols:
        if isinstance(entry[0], dict):
            out.append({k: torch.stack([o[k] for o in entry]) for k in entry[0]})
        else:
            out.append(torch.LongTensor(entry))
    return tuple(out)


def test_gather_equals_collated_dataset_items():
    t, hw = _tables()
    feed = RayFeed(**t)
    idx = torch.tensor([0, 7, 19, 20, 41, 59, 33, 33])
    ref = _collate([_reference_item(t, hw, int(i)) for i in idx])
    got = feed.gather(idx)
    assert torch.equal(got[0], ref[0]) and torch.equal(got[1], ref[1])
    for a, b in ((got[2], ref[2]), (got[3], ref[3])):
        assert set(a) == set(b)
        for k in b:
            assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
            assert torch.equal(a[k], b[k]), k


def test_epoch_visits_every_pixel_once_and_optional_tables_are_optional():
    t, hw = _tables(n_img=2, h=3, w=3)
    feed = RayFeed(uv=t["uv"], intrinsics_all=t["intrinsics_all"], pose_all=t["pose_all"], rgb_images=t["rgb_images"])
    seen = []
    for indices, img, sample, gt in feed.batches(batch_size=5, generator=torch.Generator().manual_seed(1)):
        assert set(gt) == {"rgb"} and sample["uv"].shape == (indices.numel(), 1, 2)
        assert torch.equal(img, indices // hw)
        seen.append(indices)
    seen = torch.cat(seen)
    assert seen.numel() == len(feed) == 2 * hw and torch.equal(seen.sort().values, torch.arange(2 * hw))
    assert sum(1 for _ in feed.batches(5, drop_last=True)) == (2 * hw) // 5
    b = feed.random_batch(7, generator=torch.Generator().manual_seed(2))
    assert b[2]["pose"].shape == (7, 4, 4)


def test_bubble_pdf_update_and_sampling():
    n_pix, n_pts = 40, 25
    g = torch.Generator().manual_seed(3)
    links = -torch.ones(n_pix, dtype=torch.long)
    valid = torch.randperm(n_pix, generator=g)[:n_pts].sort().values
    links[valid] = torch.arange(n_pts)
    cloud = torch.rand(n_pts, 3, generator=g)
    bub = BubblePDF(cloud, links, pdf_prune=0.2, pdf_max=0.9)
    idx = torch.arange(n_pix)
    value = torch.rand(n_pix, generator=g)
    bub.update_pdf(value, idx)
    # literal restatement of recon.py:142-152
    v = value.clone().clamp(max=0.9)
    v[v < 0.2] = 0
    ref = torch.zeros(n_pts)
    m = links != -1
    ref[links[m]] = v[m]
    assert torch.equal(bub.pdf, ref)
    pts = bub.sample_bubble(8, generator=torch.Generator().manual_seed(4))
    assert pts.shape == (8, 3)
    picked = torch.where(bub.sample_count > 0)[0]
    assert picked.numel() == 8 and (bub.pdf[picked] > 0).all()            # without replacement, only points with mass
    assert all(any(torch.equal(p, cloud[j]) for j in picked) for p in pts)
    uni = BubblePDF(cloud, links, uniform=True)
    assert uni.sample_bubble(5).shape == (5, 3)


# ---- rays sharded over ranks (gloo, world_size 2): disjoint shares of the same shuffled epoch, one PDF ------------------------
def _shard_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    t, hw = _tables()
    feed = RayFeed(**t)
    torch.manual_seed(1234 + rank)                       # ranks do NOT share a seed: the epoch's permutation comes from rank 0
    got = [(b[0].numpy().copy(), b[3]["rgb"].numpy().copy()) for b in feed.batches(7, group=dist.group.WORLD)]      # numpy: pickled by value
    # bubble PDF: every rank reports the errors of its own (ragged) shard
    links = torch.arange(3 * hw) % 11
    links[::5] = -1
    pdf = BubblePDF(torch.rand(11, 3, generator=torch.Generator().manual_seed(0)), links, pdf_prune=0.1, pdf_max=0.8)
    g = torch.Generator().manual_seed(50 + rank)
    n = 9 if rank == 0 else 4
    val, idx = torch.rand(n, generator=g), torch.randint(3 * hw, (n,), generator=g)
    pdf.update_pdf(val, idx, group=dist.group.WORLD)
    q.put((rank, got, pdf.pdf.numpy().copy(), val.numpy().copy(), idx.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_epoch_and_all_gathered_pdf_updates():
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    t, hw = _tables()
    n = 3 * hw
    as_t = lambda pairs: [(torch.from_numpy(i), torch.from_numpy(c)) for i, c in pairs]          # noqa: E731
    b0, b1 = as_t(res[0][1]), as_t(res[1][1])
    assert len(b0) == len(b1) == -(-n // 7)
    seen = []
    for k, ((i0, rgb0), (i1, rgb1)) in enumerate(zip(b0, b1)):
        size = min(7, n - 7 * k)
        assert i0.numel() == -(-size // 2) and i1.numel() == size // 2            # shard_bounds: rank 0 takes the odd one
        assert torch.equal(rgb0, t["rgb_images"].reshape(-1, 3)[i0])
        seen += i0.tolist() + i1.tolist()
    assert sorted(seen) == list(range(n))                                         # the epoch covers every pixel exactly once
    assert seen != list(range(n))                                                 # ... shuffled
    # PDF: both ranks hold what one process would hold after the updates of the whole batch, applied in rank order
    links = torch.arange(n) % 11
    links[::5] = -1
    ref = BubblePDF(torch.rand(11, 3, generator=torch.Generator().manual_seed(0)), links, pdf_prune=0.1, pdf_max=0.8)
    ref.update_pdf(torch.from_numpy(np.concatenate([res[0][3], res[1][3]])), torch.from_numpy(np.concatenate([res[0][4], res[1][4]])))
    assert np.array_equal(res[0][2], ref.pdf.numpy()) and np.array_equal(res[1][2], ref.pdf.numpy()) and float(ref.pdf.sum()) > 0
