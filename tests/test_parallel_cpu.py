"""CPU, gloo, world_size 2: the host-side logic of the N>1 path (ray sharding + flat gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from i2sdf_b200 import configs
from i2sdf_b200.parallel import allreduce_gradients, shard_bounds, shard_rays


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import contextlib
    import io
    from i2sdf_b200.network import I2SDFNetwork
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = I2SDFNetwork(configs.model_conf("synthetic"))
    for p in m.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    n = allreduce_gradients(m.parameters())
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in m.parameters())
    # ray sharding: gather shard sizes and first elements
    R = 1025
    batch = {"uv": torch.arange(R * 2, dtype=torch.float32).reshape(R, 1, 2), "pose": torch.zeros(R, 4, 4),
             "intrinsics": torch.zeros(R, 4, 4), "pointcloud": torch.zeros(7, 3)}
    sh = shard_rays(batch, rank, world)
    sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([sh["uv"].shape[0]]))
    q.put((rank, n, ok, [int(s) for s in sizes], sh["pose"].shape[0], sh["pointcloud"].shape[0], float(sh["uv"][0, 0, 0])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_ray_sharding():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n, ok, sizes, npose, ncloud, first in res:
        assert n == 800955 and ok                     # one flat 3.2 MB bucket
        assert sizes == [513, 512] and npose == sizes[rank] and ncloud == 7
    assert res[0][6] == 0.0 and res[1][6] == 513 * 2.0


def test_shard_bounds_cover_everything_once():
    for n in (0, 1, 7, 1024, 307200):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                lo, hi = shard_bounds(n, r, world)
                covered += list(range(lo, hi)) if n < 5000 else [lo, hi]
            if n < 5000:
                assert covered == list(range(n))
    eval_batch = {"uv": torch.zeros(1, 10, 2), "pose": torch.zeros(1, 4, 4), "intrinsics": torch.zeros(1, 4, 4)}
    a, b = shard_rays(eval_batch, 0, 2), shard_rays(eval_batch, 1, 2)
    assert a["uv"].shape == (1, 5, 2) and b["uv"].shape == (1, 5, 2) and a["pose"].shape == (1, 4, 4)


# ---- batch-global loss means over ragged shards (SURVEY.md §8(e) caveat 2) ---------------------------------------------------
def _loss_case(R, seed):
    g = torch.Generator().manual_seed(seed)
    out = {"rgb_values": torch.rand(R, 3, generator=g), "depth_values": torch.rand(R, generator=g) * 3,
           "weight_sum": torch.rand(R, 1, generator=g), "grad_theta": torch.randn(2 * R, 3, generator=g),
           "diff_norm": torch.rand(R, generator=g), "normal_values": torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1),
           "surface_sdf": torch.randn(4, 1, generator=g) * 0.1, "light_mask": torch.rand(R, 1, generator=g)}
    gt = {"rgb": torch.rand(R, 3, generator=g), "depth": torch.rand(R, generator=g) * 3, "depth_mask": torch.rand(R, generator=g) > 0.4,
          "normal": torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1), "normal_mask": torch.rand(R, generator=g) > 0.6,
          "mask": (torch.rand(R, 1, generator=g) > 0.5).float(), "light_mask": (torch.rand(R, 1, generator=g) > 0.5).float()}
    gt["depth_mask"][0] = True
    gt["normal_mask"][0] = True
    return out, gt


_LOSS_KW = dict(eikonal_weight=0.1, smooth_weight=0.01, mask_weight=0.2, depth_weight=0.1, normal_weight=0.05, angular_weight=0.05,
                bubble_weight=0.5, light_mask_weight=0.5)
_SHARDS = (11, 6)            # ragged on purpose: the per-ray means need the global ray count too


def _loss_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from i2sdf_b200.network import I2SDFLoss
    from i2sdf_b200.parallel import use_global_loss_means
    out, gt = _loss_case(_SHARDS[rank], 100 + rank)
    leaves = {k: v.clone().requires_grad_(True) for k, v in out.items()}
    fn = use_global_loss_means(I2SDFLoss(**_LOSS_KW))
    res = fn._forward_torch(leaves, gt, 10)          # the PyTorch restatement (the CUDA kernel takes the same divisors: GPU test)
    res["loss"].backward()
    q.put((rank, float(res["loss"]), {k: float(v) for k, v in res.items()}, {k: v.grad.numpy().copy() for k, v in leaves.items()}))      # numpy: pickled by value
    dist.barrier()
    dist.destroy_process_group()


def test_global_loss_means_equal_the_single_process_loss():
    from i2sdf_b200.network import I2SDFLoss
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loss_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the whole batch in one process, the reference's formulas
    cases = [_loss_case(_SHARDS[r], 100 + r) for r in range(world)]
    out = {k: torch.cat([c[0][k] for c in cases]).clone().requires_grad_(True) for k in cases[0][0]}
    gt = {k: torch.cat([c[1][k] for c in cases]) for k in cases[0][1]}
    fn = I2SDFLoss(**_LOSS_KW)
    whole = fn._forward_torch(out, gt, 10)
    whole["loss"].backward()
    assert abs(sum(r[1] for r in res) / world - float(whole["loss"])) < 1e-6 * abs(float(whole["loss"]))
    for k in ("rgb_loss", "eikonal_loss", "smooth_loss", "mask_loss", "depth_loss", "normal_loss", "bubble_loss", "light_mask_loss"):
        assert abs(sum(r[2][k] for r in res) / world - float(whole[k])) < 2e-6 * max(abs(float(whole[k])), 1e-3), k
    # gradient of a shard's outputs, averaged as the gradient all-reduce averages parameter gradients, = its slice of the whole
    # batch's gradient
    for k in out:
        got = torch.cat([torch.from_numpy(r[3][k]) for r in res]) / world
        assert torch.allclose(got, out[k].grad, rtol=1e-5, atol=1e-8), k
    # without the switch the per-shard means differ (this is what the test guards against)
    plain = [I2SDFLoss(**_LOSS_KW)._forward_torch(c[0], c[1], 10) for c in cases]
    assert abs(sum(float(p["depth_loss"]) for p in plain) / world - float(whole["depth_loss"])) > 1e-4


def test_seed_rank_streams():
    from i2sdf_b200.parallel import seed_rank
    assert seed_rank(42, 0) == 42                       # rank 0 keeps the reference's seed: same draws as the unsharded run
    a = torch.rand(4)
    torch.manual_seed(42)
    assert torch.equal(a, torch.rand(4))
    seeds = {seed_rank(42, r) for r in range(8)}
    assert len(seeds) == 8
    seed_rank(42, 3)
    b = torch.rand(4)
    seed_rank(42, 3)
    assert torch.equal(b, torch.rand(4)) and not torch.equal(a, b)


# ---- sharded whole-image render (inference: chunks dealt to ranks, no data-path collective) -------------------------------------
class _FakeEvalModel:
    """Stands in for I2SDFNetwork(eval) on the CPU: per-ray outputs that depend on the ray AND on the whole chunk (like the
    sampler's batch-global convergence test), so a different chunking would change the result."""
    training = False

    class density:                      # noqa: N801  (attribute path model.density.beta.device)
        beta = torch.zeros(())

    def __call__(self, inp, predict_only=False):
        uv = inp["uv"][0]
        chunk_stat = uv.sum() * 1e-3
        rgb = torch.stack([uv[:, 0], uv[:, 1], uv[:, 0] * 0 + chunk_stat], -1) + float(inp["pose"][0, 0, 3])
        out = {"rgb_values": rgb, "depth_values": uv[:, 0] + uv[:, 1], "weight_sum": (uv[:, :1] > 2).float()}
        if not predict_only:
            out["normal_map"] = rgb * 0.5
        return out


class _FakeCollectiveEvalModel(_FakeEvalModel):
    """Stand-in for a model after parallel.use_global_convergence(eval_forwards=True): every eval forward enters a collective, as the
    real sampler does once per round.  render_image(group=) deals the ranks DIFFERENT numbers of chunks; it must switch that off."""

    def __call__(self, inp, predict_only=False):
        if self.convergence_group_eval is not None:
            dist.all_reduce(torch.zeros(1), group=self.convergence_group_eval)       # would deadlock on an odd chunk count
        return super().__call__(inp, predict_only)


def _render_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world, timeout=__import__("datetime").timedelta(seconds=60))
    from i2sdf_b200.parallel import use_global_convergence
    from i2sdf_b200.render import render_image
    pose = torch.eye(4)
    pose[0, 3] = 0.25
    res = {}
    # five chunks over two ranks (3 + 2) through a model whose eval forwards would enter a collective: no hang, same image, switch restored
    fm = use_global_convergence(_FakeCollectiveEvalModel(), eval_forwards=True)
    guarded = render_image(fm, pose, torch.eye(4), (7, 9), split_n_pixels=13, group=dist.group.WORLD)
    assert fm.convergence_group_eval is not None and fm.convergence_group is not None
    use_global_convergence(fm)                                  # default: training forwards only
    assert fm.convergence_group_eval is None and fm.convergence_group is not None
    res["guarded"] = ({k: v.numpy().copy() for k, v in guarded.items()}, {})
    for name, res_hw, split in (("five_chunks", (7, 9), 13), ("one_chunk", (3, 4), 100)):
        full = render_image(_FakeEvalModel(), pose, torch.eye(4), res_hw, split_n_pixels=split, group=dist.group.WORLD)
        own = render_image(_FakeEvalModel(), pose, torch.eye(4), res_hw, split_n_pixels=split, group=dist.group.WORLD, assemble=False)
        res[name] = ({k: v.numpy().copy() for k, v in full.items()}, {k: v.numpy().copy() for k, v in own.items()})
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_render_image_equals_the_single_process_image():
    import numpy as np
    from i2sdf_b200.render import render_image
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_render_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pose = torch.eye(4)
    pose[0, 3] = 0.25
    single = render_image(_FakeEvalModel(), pose, torch.eye(4), (7, 9), split_n_pixels=13)
    for rank in range(world):                                   # the guarded render (advisor finding r01: eval collective + odd chunk count)
        for k, v in single.items():
            assert np.array_equal(res[rank]["guarded"][0][k], v.numpy()), (rank, k)
    for name, res_hw, split in (("five_chunks", (7, 9), 13), ("one_chunk", (3, 4), 100)):
        single = render_image(_FakeEvalModel(), pose, torch.eye(4), res_hw, split_n_pixels=split)
        for rank in range(world):
            full, own = res[rank][name]
            assert set(full) == set(single)
            for k, v in single.items():
                assert np.array_equal(full[k], v.numpy()), (name, rank, k)          # every rank holds the single-process image, bit for bit
        # without assembly: the ranks' pieces are disjoint and together cover the image
        total = res_hw[0] * res_hw[1]
        owner = np.arange(total) // split % world
        for k, v in single.items():
            for rank in range(world):
                own = res[rank][name][1]
                if k in own:
                    assert np.array_equal(own[k][owner == rank], v.numpy()[owner == rank]) and not own[k][owner != rank].any()
                else:
                    assert not (owner == rank).any()


# ---- persistent flat gradient bucket (parallel.GradBucket) -------------------------------------------------------------------
def _bucket_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from i2sdf_b200.parallel import GradBucket
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(())), torch.nn.Parameter(torch.randn(2, 2))]
    bucket = GradBucket(ps)
    n_flat = bucket.flat.numel()
    ok = True
    for step in range(2):
        bucket.zero()
        assert all(p.grad is None for p in ps)
        # rank 1 has no gradient for parameter 1 (a loss term absent on its shard); parameter 3 gets none on any rank
        loss = (ps[0] * (rank + 1.0 + step)).sum() + ps[2] * 3.0
        if rank == 0:
            loss = loss + (ps[1] * 2.0).sum()
        loss.backward()
        n = bucket.allreduce()
        ok &= n == n_flat
        ok &= all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(ps, bucket.views))                 # every .grad IS its bucket view
        ok &= torch.allclose(ps[0].grad, torch.full((5, 3), 1.5 + step)) and torch.allclose(ps[1].grad, torch.full((7,), 1.0))
        ok &= torch.allclose(ps[2].grad, torch.tensor(3.0)) and torch.equal(ps[3].grad, torch.zeros(2, 2))
    q.put((rank, bool(ok), n_flat))
    dist.barrier()
    dist.destroy_process_group()


def test_grad_bucket_allreduces_in_place_with_uneven_none_grads():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and res[0][2] == res[1][2] == 16 + 8 + 4 + 4
