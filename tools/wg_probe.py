"""GPU probe of the MN-major weight-gradient kernel: prints the error of both descriptor variants (env I2SDF_WG_VARIANT)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2sdf_b200 import configs
from i2sdf_b200.network import I2SDFNetwork
m = I2SDFNetwork(configs.model_conf("synthetic")).cuda()
core = m._ready_core()
g = torch.Generator().manual_seed(1)
M = int(os.environ.get("M", 1000))
P = torch.randn(M, 256, generator=g).cuda(); X = torch.randn(M, 256, generator=g).cuda()
sP, sX = core.planes_pack(P), core.planes_pack(X)
r = lambda t: core.planes_unpack(core.planes_pack(t), M).double()
ref = r(P).T @ r(X)
dW = core.planes_wgrad([sP], [sX], M, 256, 256)
torch.cuda.synchronize()
err = (dW.double() - ref).abs().max().item()
print("variant", os.environ.get("I2SDF_WG_VARIANT", "0"), "M", M, "max abs err", err, "ref max", ref.abs().max().item())
if os.environ.get("TIME"):
    M = 99328
    P = torch.randn(M, 256, generator=g).cuda(); X = torch.randn(M, 256, generator=g).cuda()
    sP, sX = core.planes_pack(P), core.planes_pack(X)
    for n in (1, 2):
        for _ in range(3): core.planes_wgrad([sP] * n, [sX] * n, M, 256, 256)
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10): core.planes_wgrad([sP] * n, [sX] * n, M, 256, 256, colsum=True)
        e1.record(); torch.cuda.synchronize()
        print(f"wgrad {n} term(s) M={M}: {e0.elapsed_time(e1) / 10 * 1000:.1f} us per call (incl. zero fill)")
