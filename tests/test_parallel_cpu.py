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
