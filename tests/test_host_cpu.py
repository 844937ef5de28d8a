"""CPU: host-side logic and the C-ABI library surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

from i2sdf_b200 import _lib, configs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "i2sdf_b200.h")).read()
    declared = set(re.findall(r"\b(i2sdf_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.i2sdf_abi_version() == 1


def test_desc_struct_layout_matches_header():
    # 15 int32 + 6 float + 4 pointers, natural alignment
    assert ctypes.sizeof(_lib.Desc) == 15 * 4 + 6 * 4 + 4 + 4 * 8


def test_module_surface_matches_reference_state_dict():
    from i2sdf_b200.network import I2SDFNetwork
    torch.manual_seed(0)
    m = I2SDFNetwork(configs.model_conf("synthetic_light_mask"))
    keys = set(m.state_dict().keys())
    want = {f"implicit_network.lin{l}.{s}" for l in range(7) for s in ("bias", "weight_g", "weight_v")}
    want |= {f"rendering_network.lin{l}.{s}" for l in range(4) for s in ("bias", "weight_g", "weight_v")}
    want |= {f"light_network.lin{l}.{s}" for l in range(2) for s in ("bias", "weight_g", "weight_v")}
    want |= {"density.beta"}
    assert keys == want
    assert m.implicit_network.lin2.weight_v.shape == (217, 256)       # layer feeding the skip concat
    assert m.implicit_network.lin6.weight_v.shape == (257, 256)
    assert m.rendering_network.lin0.weight_v.shape == (256, 283)
    assert sum(p.numel() for p in m.parameters()) == 635965
    m2 = I2SDFNetwork(configs.model_conf("synthetic"))
    assert sum(p.numel() for p in m2.parameters()) == 800955


def test_geometric_init_reproduces_reference_weights():
    """Same seed -> same weights as the reference constructor (fixture weights were made by the reference ctor,
    then perturbed only in the embedding columns / colour heads; untouched tensors must match bit for bit)."""
    import numpy as np
    from i2sdf_b200.network import I2SDFNetwork
    w = np.load(os.path.join(ROOT, "tests", "golden", "weights_synthetic.npz"))
    torch.manual_seed(0)
    m = I2SDFNetwork(configs.model_conf("synthetic"))
    sd = m.state_dict()
    for k in ("implicit_network.lin1.weight_v", "implicit_network.lin5.weight_g", "implicit_network.lin8.weight_v",
              "implicit_network.lin8.bias", "implicit_network.lin3.weight_v"):
        assert torch.equal(sd[k], torch.from_numpy(w[k])), k
    assert torch.equal(sd["implicit_network.lin0.weight_v"][:, :3], torch.from_numpy(w["implicit_network.lin0.weight_v"][:, :3]))
    # colour stack was scaled by 1.5 after construction
    assert torch.allclose(sd["rendering_network.lin2.weight_v"] * 1.5, torch.from_numpy(w["rendering_network.lin2.weight_v"]))


def test_forward_on_cpu_fails_loudly():
    from i2sdf_b200.network import I2SDFNetwork
    m = I2SDFNetwork(configs.model_conf("synthetic")).eval()
    with pytest.raises(_lib.I2SDFError):
        m({"uv": torch.zeros(1, 4, 2), "pose": torch.eye(4)[None], "intrinsics": torch.eye(4)[None]})


def test_unsupported_configs_are_rejected():
    from i2sdf_b200.network import I2SDFNetwork
    conf = configs.model_conf("synthetic")
    conf["bg_network"] = {}
    with pytest.raises(_lib.I2SDFError):
        I2SDFNetwork(conf)
    conf = configs.model_conf("synthetic")
    conf["rendering_network"]["mode"] = "idr"
    with pytest.raises(_lib.I2SDFError):
        I2SDFNetwork(conf)


def test_pixel_grid_matches_reference_get_uv():
    """dataset/eval_dataset.py:144-148: np.mgrid -> flip -> reshape(2,-1).T"""
    import numpy as np
    from i2sdf_b200.render import pixel_grid
    H, W = 5, 7
    uv = np.mgrid[0:H, 0:W].astype(np.int32)
    ref = torch.from_numpy(np.flip(uv, axis=0).copy()).float().reshape(2, -1).transpose(1, 0)
    assert torch.equal(pixel_grid((H, W)), ref)


def test_any_optimizer_step_invalidates_the_packed_weights_key():
    """network._OPT_EPOCH is part of the key an eval forward checks before reusing the packed weights; fused optimizers do
    not bump tensor versions, so the global optimizer post-step hook has to."""
    import torch
    from i2sdf_b200 import network
    lin = torch.nn.Linear(2, 2)
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    lin(torch.ones(1, 2)).sum().backward()
    before = network._OPT_EPOCH[0]
    opt.step()
    assert network._OPT_EPOCH[0] == before + 1


def test_loss_restatement_matches_the_oracle_loss_on_cpu():
    """I2SDFLoss._forward_torch (the PyTorch restatement the CUDA kernel is checked against on the GPU) vs the oracle's
    recon_loss, which the training fixtures pin on the reference's loss values (model/network/__init__.py:338-406)."""
    import torch
    from i2sdf_b200.network import I2SDFLoss
    from oracle import i2sdf_oracle as orc
    g = torch.Generator().manual_seed(0)
    R = 50
    out = {"rgb_values": torch.rand(R, 3, generator=g), "depth_values": torch.rand(R, generator=g) * 3, "weight_sum": torch.rand(R, 1, generator=g),
           "normal_values": torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=1), "grad_theta": torch.randn(2 * R, 3, generator=g),
           "diff_norm": torch.rand(R, generator=g), "surface_sdf": torch.randn(9, 1, generator=g), "light_mask": torch.rand(R, 1, generator=g)}
    gt = {"rgb": torch.rand(R, 3, generator=g), "depth": torch.rand(R, generator=g) * 3, "depth_mask": torch.rand(R, generator=g) > 0.3,
          "normal": torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=1), "normal_mask": torch.rand(R, generator=g) > 0.4,
          "mask": (torch.rand(R, 1, generator=g) > 0.5).float(), "light_mask": (torch.rand(R, 1, generator=g) > 0.7).float()}
    kw = dict(eikonal_weight=0.1, smooth_weight=0.01, depth_weight=0.1, normal_weight=0.05, bubble_weight=0.5, light_mask_weight=0.5,
              mask_weight=0.2)
    mine = I2SDFLoss(smooth_iter=10, **kw)._forward_torch(out, gt, 100)
    ref = orc.recon_loss(out, gt, angular_weight=0.05, smooth_active=True, **kw)
    assert abs(float(mine["loss"]) - float(ref)) < 1e-6 * abs(float(ref))
    assert set(mine) == {"loss", "rgb_loss", "eikonal_loss", "smooth_loss", "mask_loss", "depth_loss", "normal_loss", "angular_loss",
                         "bubble_loss", "light_mask_loss"}
    assert float(mine["angular_loss"]) == float(mine["normal_loss"])         # the reference's "angular" term IS the L1 normal loss (:368-371)
    early = I2SDFLoss(smooth_iter=1000, **kw)._forward_torch(out, gt, 100)
    with pytest.raises(_lib.I2SDFError):                                    # the module itself has no CPU path
        I2SDFLoss(**kw)(out, gt, 100)
    assert float(early["smooth_loss"]) == 0.0                                # smoothness term only after smooth_iter (:347-351)


def test_opt_in_helpers_fail_loudly_off_the_gpu():
    """No CPU paths: the one-launch Adam rejects CPU parameters and unsupported options; the global-convergence switch needs
    an initialised process group."""
    import pytest
    import torch
    from i2sdf_b200 import _lib
    from i2sdf_b200.optim import Adam
    from i2sdf_b200.parallel import use_global_convergence
    lin = torch.nn.Linear(3, 2)
    opt = Adam(lin.parameters(), lr=1e-3, eps=1e-15)
    lin(torch.ones(1, 3)).sum().backward()
    with pytest.raises(_lib.I2SDFError):
        opt.step()
    with pytest.raises(_lib.I2SDFError):
        Adam(lin.parameters(), weight_decay=0.1)
    with pytest.raises(RuntimeError):
        use_global_convergence(lin)
    sd = opt.state_dict()
    assert set(sd) == {"state", "param_groups"} and sd["param_groups"][0]["eps"] == 1e-15


def test_ctypes_structs_match_the_header_field_by_field(tmp_path):
    """Every struct of include/i2sdf_b200.h against its ctypes mirror in i2sdf_b200/_lib.py: a C probe compiled with gcc prints
    sizeof and the offset of every field; the binding must agree (a silent mismatch would hand the kernels shifted pointers)."""
    import ctypes
    import subprocess
    from i2sdf_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = {"i2sdf_desc": _lib.Desc, "i2sdf_loss_args": _lib.LossArgs, "i2sdf_wnorm_job": _lib.WnormJob, "i2sdf_wnorm_batch": _lib.WnormBatch,
             "i2sdf_adam_job": _lib.AdamJob, "i2sdf_adam_batch": _lib.AdamBatch}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "i2sdf_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'printf("{cname} . %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['return 0; }']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    seen = 0
    for ln in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines():
        cname, fname, val = ln.split()
        cls = pairs[cname]
        want = ctypes.sizeof(cls) if fname == "." else getattr(cls, fname).offset
        assert int(val) == want, (cname, fname, int(val), want)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
    assert _lib.WNORM_MAX_JOBS == 28 and _lib.ADAM_MAX_JOBS == 64          # I2SDF_WNORM_MAX_JOBS / I2SDF_ADAM_MAX_JOBS


def test_c_abi_from_plain_c(tmp_path):
    """The boundary is a C ABI: tests/cabi/probe.c (C11, libdl only, no torch, no CUDA headers) binds the library the way a
    foreign-function stub would, checks argument validation, and creates a handle — which must fail LOUDLY (I2SDF_E_NOGPU and a
    message, no crash, no fallback) on a host without a B200 and succeed on one."""
    import subprocess
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cabi", "probe.c"),
                    "-o", str(exe), "-ldl"], check=True)
    r = subprocess.run([str(exe), _lib.LIB_PATH], check=True, capture_output=True, text=True, timeout=120)
    lines = {ln.split()[0]: ln for ln in r.stdout.splitlines()}
    assert lines["abi"].split()[1] == lines["abi"].split()[3] == str(_lib.ABI_VERSION)
    assert lines["slot256"].split()[1] == str(_lib.load().i2sdf_planes_slot_bytes(1000, 256))
    assert " rc -1 " in lines["create_bad"] and "handle null" in lines["create_bad"] and "256" in lines["create_bad"]
    assert " rc -1 " in lines["loss_null"] and "required" in lines["loss_null"]
    if torch.cuda.is_available():
        assert " rc 0 " in lines["create_ok"] and lines["destroy"].endswith("rc 0")
    else:
        assert " rc -3 " in lines["create_nogpu"] and "handle null" in lines["create_nogpu"]  # I2SDF_E_NOGPU


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under i2sdf_b200/ may import it (statically: every import statement of every
    module; dynamically: importing the whole package in a fresh interpreter leaves no `oracle*` module loaded), and the C sources
    do not mention it either."""
    import ast
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "i2sdf_b200")
    mods = [f for f in os.listdir(pkg) if f.endswith(".py")]
    assert len(mods) >= 8
    for f in mods:
        tree = ast.parse(open(os.path.join(pkg, f)).read())
        for node in ast.walk(tree):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            assert not any(n == "oracle" or n.startswith("oracle.") for n in names), (f, names)
    code = ("import sys, importlib; sys.path.insert(0, %r);\n"
            "[importlib.import_module('i2sdf_b200.' + m) for m in %r];\n"
            "bad = [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]; print('LOADED', bad)") % (ROOT, [m[:-3] for m in mods if m != "__init__.py"])
    r = subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True, timeout=300)
    assert "LOADED []" in r.stdout, r.stdout
    for f in os.listdir(os.path.join(pkg, "csrc")):
        if f.endswith((".cu", ".cuh")):
            assert "oracle" not in open(os.path.join(pkg, "csrc", f)).read().lower(), f


def test_synthetic_inputs_flop_model_and_grid_axes():
    """The neutral input generator equals the oracle's; the published FLOP model gives SURVEY.md §8(d)'s numbers; the grid axes are the
    reference's (utils/plots.py:440-489) and the explicit point list has numpy.meshgrid's order."""
    import numpy as np
    import bench
    from i2sdf_b200 import configs
    from i2sdf_b200.grid import grid_axes_from_points, grid_axes_uniform, grid_points
    from i2sdf_b200.synthetic import synthetic_rays
    from oracle import i2sdf_oracle as orc
    for tl in (False, True):
        a, b = synthetic_rays(33, seed=4, train_layout=tl), orc.synthetic_rays(33, seed=4, train_layout=tl)
        assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
    fm = bench.flop_model(configs.model_conf("synthetic"))
    assert fm["sdf_eval"] == 918016 and fm["ray_sample"] == 2506752
    assert bench.flop_model(configs.model_conf("synthetic_light_mask"))["ray_sample"] == 1917184
    assert abs(bench.step_flops(fm, 1024, 5, False) / 1024 - 830.7e6) < 0.1e6               # 830.7 MFLOP per ray, worst case
    x, y, z = grid_axes_uniform(5, (-2.0, 2.0))
    assert np.array_equal(x, np.linspace(-2.0, 2.0, 5)) and x is y and y is z
    pts = grid_points([0.0, 1.0], [10.0, 20.0, 30.0], [5.0])
    assert pts.shape == (6, 3) and pts.tolist()[:3] == [[0.0, 10.0, 5.0], [1.0, 10.0, 5.0], [0.0, 20.0, 5.0]]   # (j, i, k): y outer, x, z inner
    cloud = torch.tensor([[0.0, 0.0, 0.0], [1.0, 2.0, 4.0]])
    (ax, ay, az), length, sa = grid_axes_from_points(cloud, 11)
    assert sa == 0 and len(ax) == 11 and abs(length - 1.2) < 1e-6 and abs((ay[1] - ay[0]) - 0.12) < 1e-6 and abs(ay[0] + 0.1) < 1e-6 and az[-1] >= 4.1 - 1e-6
