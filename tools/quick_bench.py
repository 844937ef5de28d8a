"""Ad-hoc device timing of individual entry points (development aid; bench.py is the contract)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs          # noqa: E402
from i2sdf_b200.network import I2SDFNetwork  # noqa: E402
from i2sdf_b200 import synthetic as orc  # noqa: E402   # (neutral input generator: perf tools do not touch oracle/)


def timeit(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    dev = torch.device("cuda:0")
    R = int(os.environ.get("R", 1024))
    conf = configs.model_conf("synthetic")
    torch.manual_seed(0)
    m = I2SDFNetwork(conf)
    with torch.no_grad():
        m.density.beta.fill_(0.01)
    m = m.to(dev).eval()
    core = m._ready_core()
    print("tensor cores:", core.uses_tensor_cores)
    inp = {k: v.to(dev) for k, v in orc.synthetic_rays(R, seed=1).items()}
    M = R * 128
    pts = (torch.rand(M, 3, device=dev) - 0.5) * 3
    t = timeit(lambda: core.sdf_forward(pts))
    print(f"sdf_forward (sdf only) M={M}: {t:.3f} ms  {M / t / 1e3:.1f} M pts/s  {M * 918016 / t / 1e9:.1f} TFLOP/s algorithmic")
    Mf = R * 97
    ptsf = pts[:Mf].contiguous()
    t = timeit(lambda: core.sdf_forward(ptsf, want_grad=True))
    print(f"sdf_forward (sdf + grad_x, chain F.. R..) M={Mf}: {t:.3f} ms  {Mf / t / 1e3:.1f} M pts/s")
    o, d, dn = core.rays(inp["uv"], inp["pose"], inp["intrinsics"])
    beta = m.density.beta.detach()
    t = timeit(lambda: core.sample(o, d, beta))
    z, _, info = core.sample(o, d, beta, want_info=True)
    print(f"sampler R={R}: {t:.3f} ms rounds={int(info[0])}")
    t = timeit(lambda: core.render(o, d, dn, z, beta))
    print(f"render (main pass + composite) R={R}: {t:.3f} ms")
    t = timeit(lambda: m(inp))
    print(f"forward eval R={R}: {t:.3f} ms  {R * 97 / t / 1e3:.2f} M ray-samples/s  {R / t * 1e3:.0f} rays/s")


if __name__ == "__main__":
    main()
