"""CPU: the oracle (oracle/i2sdf_oracle.py) re-checked against reference outputs stored in tests/golden."""
import pytest
import torch

from golden_util import EVAL_CASES, TRAIN_CASES, Case, relerr
from oracle import i2sdf_oracle as orc

torch.set_num_threads(8)


@pytest.mark.parametrize("name", EVAL_CASES)
def test_oracle_eval_matches_reference(name):
    c = Case(name)
    trace = {}
    with torch.no_grad():
        out = orc.render(c.spec, c.params, c.inputs, training=False, trace=trace)
    # sampler: discrete decisions and z values identical to the reference's
    assert trace["n_rounds"] == int(c.trace["n_rounds"])
    assert torch.equal(trace["z"], c.mid["z_all"])
    for i, r in enumerate(trace["rounds"]):
        assert torch.equal(r["inds"], c.trace[f"round{i}_inds"])
        if "perm" in r:
            assert torch.equal(torch.gather(torch.cat([r["z"], r["samples"]], -1), 1, r["perm"]),
                               c.trace[f"round{i}_z_merged"])
    assert set(out) == set(c.ref)
    for k, v in c.ref.items():
        assert out[k].shape == v.shape, k
        assert relerr(out[k], v) < 2e-5, k
    assert relerr(trace["sdf"], c.mid["sdf"]) < 1e-6
    assert relerr(trace["grad"], c.mid["grad"]) < 5e-6
    n = c.mid["feat_head"].shape[0]
    assert relerr(trace["feat"][:n], c.mid["feat_head"]) < 1e-6


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_oracle_train_matches_reference(name):
    c = Case(name)
    P = {k: v.clone().requires_grad_(True) for k, v in c.params.items()}
    trace = {}
    out = orc.render(c.spec, P, c.inputs, training=True, tape=c.tape, trace=trace)
    assert torch.equal(trace["z"], c.mid["z_all"])
    assert set(out) == set(c.ref)
    for k, v in c.ref.items():
        assert out[k].shape == v.shape, k
        assert relerr(out[k], v) < 5e-5, k
    loss = orc.recon_loss(out, c.gt, **c.loss_kwargs())
    assert abs(loss.item() - c.ref_loss) < 1e-5 * abs(c.ref_loss)
    loss.backward()
    for k, g in c.refgrads().items():
        mine = P[k].grad
        assert mine is not None, k
        if "full" in g:
            assert relerr(mine, g["full"]) < 1e-4, k
        else:
            assert abs(mine.norm().item() - g["norm"].item()) < 1e-4 * g["norm"].item() + 1e-9, k
            sub = mine.flatten()[::97]
            assert (sub - g["sub"]).abs().max() <= 1e-4 * g["sub"].abs().max() + 1e-9, k


def test_uniform_sampler_config_c1():
    """BASELINE config 1 (the reference's own CPU-runnable case): 64 uniform samples on [0, 6] -> 63 composited points/ray.
    Fixture: the unmodified reference with its own UniformSampler swapped in (tests/golden/make_golden.py: uniform_case)."""
    c = Case("eval_uniform_c1")
    R = c.inputs["uv"].shape[1]
    z = c.mid["z_all"]
    assert z.shape == (R, 64) and torch.equal(z, (torch.linspace(0, 1, 64) * c.spec.far)[None].repeat(R, 1))
    with torch.no_grad():
        out = orc.render(c.spec, c.params, c.inputs, training=False, z_override=z)
    assert set(out) == set(c.ref)
    for k, v in c.ref.items():
        assert out[k].shape == v.shape, k
        assert relerr(out[k], v) < 2e-6, k
    assert (out["weight_sum"] <= 1 + 1e-5).all() and float(out["weight_sum"].max()) > 0.5     # the rays do hit the surface


def test_loss_all_terms_matches_reference():
    """Every term of the reference's I2SDFLoss switched on (tests/golden/loss_all_terms.npz: values at two steps and the
    reference's autograd gradients w.r.t. every model output).  Pins (1) the oracle's recon_loss and (2) the PyTorch restatement
    the CUDA loss kernel is checked against on the GPU (i2sdf_b200.network.I2SDFLoss._forward_torch), values and gradients."""
    import os
    import numpy as np
    from golden_util import GOLDEN
    from i2sdf_b200.network import I2SDFLoss
    d = np.load(os.path.join(GOLDEN, "loss_all_terms.npz"))
    t = lambda a: torch.from_numpy(np.asarray(a).copy())          # noqa: E731
    out = {k[4:]: t(d[k]) for k in d.files if k.startswith("out_")}
    gt = {k[3:]: t(d[k]) for k in d.files if k.startswith("gt_")}
    kw = dict(eikonal_weight=0.1, smooth_weight=0.01, mask_weight=0.2, depth_weight=0.1, normal_weight=0.05, angular_weight=0.05,
              bubble_weight=0.5, light_mask_weight=0.5)
    for step in (5, 100):
        ref = {k[len(f"ref{step}_"):]: float(d[k]) for k in d.files if k.startswith(f"ref{step}_")}
        refgrad = {k[len(f"refgrad{step}_"):]: t(d[k]) for k in d.files if k.startswith(f"refgrad{step}_")}
        assert len(ref) == 10 and (ref["smooth_loss"] > 0) == (step > 10)
        o_loss = orc.recon_loss(out, gt, smooth_active=step > 10, **kw)
        assert abs(float(o_loss) - ref["loss"]) < 1e-6 * abs(ref["loss"])
        leaves = {k: v.clone().requires_grad_(True) for k, v in out.items()}
        mine = I2SDFLoss(smooth_iter=10, **kw)._forward_torch(leaves, gt, step)
        assert set(mine) == set(ref)
        for k, v in ref.items():
            assert abs(float(mine[k]) - v) <= 1e-6 * max(abs(v), 1e-3), (step, k, float(mine[k]), v)
        mine["loss"].backward()
        for k, g in refgrad.items():
            got = leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(leaves[k])
            assert relerr(got, g) < 1e-5 or float(g.abs().max()) == 0.0 == float(got.abs().max()), (step, k)


@pytest.mark.parametrize("name", ["eval_synthetic_soft", "eval_light_sharp"])
def test_oracle_predict_only_returns_the_early_dict(name):
    """forward(input, predict_only=True) returns before the normals / training extras (network/__init__.py:156-173): rgb, depth,
    weight_sum (+ light_mask) only, same values as the full call."""
    c = Case(name)
    with torch.no_grad():
        out = orc.render(c.spec, c.params, c.inputs, training=False, predict_only=True)
    want = {"rgb_values", "depth_values", "weight_sum"} | ({"light_mask"} if c.spec.light_dims else set())
    assert set(out) == want
    for k in want:
        assert relerr(out[k], c.ref[k]) < 2e-5, k
