"""GPU parity, direct stage tests the round-1 review asked for (rows a1, a3, a4 in training mode, §8(f)-2 at full resolution).

Named to sort last: the suite runs with `-x`, and these are the newest tests.  Same metric as tests/test_parity_gpu.py
(max|a-b| / max|b| per tensor); every test prints what it measured.
"""
import pytest
import torch

from golden_util import TRAIN_CASES, Case, general_cameras, relerr
from oracle import i2sdf_oracle as orc

pytestmark = pytest.mark.gpu


def _model(case, training=False):
    from i2sdf_b200.network import I2SDFNetwork
    conf = dict(case.model_conf)
    conf["use_normal"] = training
    m = I2SDFNetwork(conf)
    m.load_state_dict({k: v for k, v in case.params.items()}, strict=True)
    m = m.cuda()
    m.train(training)
    return m


@pytest.mark.parametrize("layout", ["eval", "train"])
def test_camera_rays_match_the_oracle(layout):
    """Row a1: i2sdf_rays == get_camera_params + lift + the flatten / normalise of I2SDFNetwork.forward
    (utils/rend_util.py:92-147, model/network/__init__.py:86-93): origins exact, unit directions and their norms to fp32 rounding."""
    c = Case("eval_synthetic_soft")
    core = _model(c)._ready_core()
    g = torch.Generator().manual_seed(17)
    worst = 0.0
    for R in (1, 37, 1024, 4099):
        if layout == "eval":             # one camera, P = R pixels (dataset/eval_dataset.py:150-168)
            pose, K = general_cameras(1, g)
            uv = (torch.rand(1, R, 2, generator=g) * torch.tensor([320.0, 240.0]))
        else:                            # one camera per ray, P = 1 (dataset/train_dataset.py:169-192)
            pose, K = general_cameras(R, g)
            uv = (torch.rand(R, 1, 2, generator=g) * torch.tensor([320.0, 240.0]))
        o_ref, d_ref, n_ref = orc.flatten_rays(uv, pose, K)
        o, d, dn = core.rays(uv.cuda(), pose.cuda(), K.cuda())
        assert o.shape == (R, 3) and d.shape == (R, 3) and dn.shape == (R,)
        assert torch.equal(o.cpu(), o_ref)                                  # a copy of pose[:, :3, 3]
        e_d = float((d.cpu() - d_ref).abs().max())                          # unit vectors: absolute = relative
        e_n = float(((dn.cpu() - n_ref).abs() / n_ref).max())
        worst = max(worst, e_d, e_n)
        # the 4-term dot products of the pose multiply are summed in a different order than the CPU bmm: a few ulp (measured: 0 with one
        # camera, 4.5e-7 with a camera per ray)
        assert e_d < 2e-6 and e_n < 2e-6, (layout, R, e_d, e_n)
        assert float((d.norm(dim=-1) - 1).abs().max()) < 1e-6
    print(f"rays ({layout} layout): worst |d - d_ref| / dnorm rel err {worst:.2e}")


@pytest.mark.parametrize("name", ["train_synthetic", "eval_synthetic_sharp"])
def test_sampler_initial_samples_are_the_references_bit_for_bit(name):
    """Row a3: the initial 128 samples per ray - stratified jitter driven by the reference's recorded `torch.rand` tape in training
    (ray_sampler.py:33-41), the plain linspace in eval (:30-31) - against the z's the REFERENCE drew (fixture `round0_z`).
    The C ABI does not expose the workspace, so the samples are read where the algorithm itself exposes them: with a wide density
    (beta = 5) the error bound holds after ONE round, the final draw is taken from the initial z's, and the 32 extra samples
    (ray_sampler.py:221-226) are z_init[:, perm] verbatim - four permutations walk all 128 indices."""
    c = Case(name)
    m = _model(c, training=c.training)
    core = m._ready_core()
    o, d, _ = orc.flatten_rays(c.inputs["uv"], c.inputs["pose"], c.inputs["intrinsics"])
    R = o.shape[0]
    z0 = c.trace["round0_z"]
    assert z0.shape == (R, 128)
    beta = torch.tensor(5.0, device="cuda")
    seen = 0
    for k in range(4):
        perm = torch.arange(32 * k, 32 * k + 32)
        tape = {"extra_perm": perm}
        if c.training:
            tape.update(jitter=c.tape["jitter"], u_final=c.tape["u_final"])
        z, _, info = core.sample(o.cuda(), d.cuda(), beta, tape, want_info=True)
        assert int(info[0]) == 1 and int(info[1]) == 128, info          # converged in the first round: the extras index the initial z's
        zc = z.cpu()
        want = z0[:, perm]                                                # [R, 32]
        present = (zc[:, :, None] == want[:, None, :]).any(1)            # bitwise equality, per expected value
        assert bool(present.all()), (name, k, int((~present).sum()))
        seen += int(present.sum())
    assert seen == R * 128
    print(f"{name}: all {seen} initial samples ({'jittered' if c.training else 'linspace'}) equal the reference's bit for bit")


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_training_sampler_on_the_reference_tape(name):
    """Row a4 in training mode, by value (the full-pipeline training test only bounds rendered outputs): the sampler driven by the
    reference's recorded draws (jitter, u, randperm, randint) against the z's / eikonal sample the reference produced from them."""
    c = Case(name)
    m = _model(c, training=True)
    core = m._ready_core()
    o, d, _ = orc.flatten_rays(c.inputs["uv"], c.inputs["pose"], c.inputs["intrinsics"])
    tape = {"jitter": c.tape["jitter"], "u_final": c.tape["u_final"], "extra_perm": c.tape["extra_perm"], "eik_idx": c.tape["eik_idx"]}
    z, z_eik, info = core.sample(o.cuda(), d.cuda(), m.density.beta.detach(), tape, want_info=True)
    assert int(info[0]) == int(c.trace["n_rounds"]) and int(info[1]) == int(c.trace["n_final"])          # counts exact
    ref = c.mid["z_all"]
    zc = z.cpu()
    assert zc.shape == ref.shape
    assert torch.equal(torch.sort(zc, -1)[0], zc)
    assert (zc[:, 0] == c.spec.near).all() and (zc[:, -1] == c.spec.far).all()
    # the eikonal pick is an integer gather of the sampler's own output (ray_sampler.py:233-234): exact
    assert torch.equal(z_eik.cpu().reshape(-1), torch.gather(zc, 1, c.tape["eik_idx"].long()[:, None])[:, 0])
    hit = (c.mid["sdf"].reshape(ref.shape[0], -1).min(-1)[0] < 0)
    close_all = float(((zc - ref).abs() < 1e-3).float().mean())
    close_hit = float(((zc - ref).abs() < 1e-3)[hit].float().mean()) if hit.any() else 1.0
    print(f"{name}: training sampler on the reference tape: z within 1e-3 of the reference: {close_hit:.4f} of surface-hitting rays' samples, "
          f"{close_all:.4f} of all; rounds {int(info[0])}")
    # measured: 1.0000 of the surface-hitting rays' samples on both fixtures, 0.9866 / 0.9767 of all (rays that miss the surface have a
    # noise-dominated up-sampling pdf, see test_sampler_rounds_on_identical_inputs)
    assert close_hit > 0.97 and close_all > 0.93, (close_hit, close_all)


def test_whole_image_driver_at_full_resolution():
    """§8(f)-2 / BASELINE configs[2] (C3): one 480 x 640 view = 307 200 rays in 65 536-ray chunks (config/synthetic.yml split_n_pixels,
    model/eval/recon.py:161-172), the last chunk ragged (45 056).  The driver's image == the chunks rendered one by one."""
    from i2sdf_b200.render import pixel_grid, render_image
    c = Case("eval_synthetic_sharp")
    m = _model(c)
    H, W, chunk = 480, 640, 65536
    pose = torch.eye(4)
    pose[2, 3] = -1.5
    K = torch.eye(4)
    K[0, 0] = K[1, 1] = 600.0
    K[0, 2], K[1, 2] = W / 2, H / 2
    img = render_image(m, pose, K, (H, W), split_n_pixels=chunk)
    assert set(img) == {"rgb_values", "depth_values", "weight_sum", "normal_map"}
    assert img["rgb_values"].shape == (H * W, 3) and img["depth_values"].shape == (H * W, 1)
    for k, v in img.items():
        assert bool(torch.isfinite(v).all()), k
    ws = img["weight_sum"]
    assert float(ws.min()) >= 0.0 and float(ws.max()) <= 1.0 + 1e-5
    # a few pixels against the oracle (the sphere of the geometric init sits in the middle of the view: the centre pixel hits it, the
    # corner does not); W-sharp weights run all 5 rounds in any batch, so a ray renders the same alone and inside its chunk
    pix = torch.tensor([(H // 2) * W + W // 2, 0, (H // 2) * W + W // 2 + 40, (H // 2 - 30) * W + W // 2])
    uv_all = pixel_grid((H, W))
    with torch.no_grad():
        ref = orc.render(c.spec, c.params, {"uv": uv_all[pix][None], "pose": pose[None], "intrinsics": K[None]}, training=False)
    assert float(ref["weight_sum"][0]) > 0.99 and float(ref["weight_sum"][1]) < 1e-3
    for k in ("rgb_values", "depth_values", "weight_sum"):
        e = relerr(img[k][pix.cuda()].reshape(ref[k].shape), ref[k])
        assert e < 5e-3, (k, e)                                            # loose on purpose: C3 parity proper is tests/test_scale_gpu.py
    uv = pixel_grid((H, W), torch.device("cuda"))
    for lo in (0, 4 * chunk):                                              # the first chunk and the ragged last one
        hi = min(lo + chunk, H * W)
        part = m({"uv": uv[None, lo:hi], "pose": pose[None].cuda(), "intrinsics": K[None].cuda()})
        for k, v in part.items():
            assert torch.equal(img[k][lo:hi], v.reshape(hi - lo, -1)), (k, lo)
    print(f"full-resolution view: {H * W} rays in {-(-H * W // chunk)} chunks, coverage {float((ws > 0.5).float().mean()):.3f}")


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_training_mode_predict_only_is_the_early_dict(name):
    """`self.model.forward(data, True)` with the module in training mode and the one-camera-per-ray layout - the call the trainer's
    bubble-PDF initialisation makes (model/trainer/recon.py:185-194) - returns the dict of network/__init__.py:156-173 (rgb, depth,
    weight_sum, + light_mask) before any training extra is drawn; on the reference's z's the values are the full call's."""
    c = Case(name)
    m = _model(c, training=True)
    m._tape_override = {"z_all": c.mid["z_all"], "z_eik": c.mid["z_eik"]}
    inp = {k: v.cuda() for k, v in c.inputs.items()}
    state = torch.cuda.get_rng_state()
    out = m.forward(inp, True)
    assert torch.equal(state, torch.cuda.get_rng_state())                 # z's given: nothing random is drawn (no eikonal / neighbour points)
    want = {"rgb_values", "depth_values", "weight_sum"} | ({"light_mask"} if c.spec.light_dims is not None else set())
    assert set(out) == want
    worst = 0.0
    for k in want:
        assert out[k].shape == c.ref[k].shape, k
        e = relerr(out[k], c.ref[k])
        worst = max(worst, e)
        assert e < 1e-4, (k, e)                                           # measured 3.7e-6 / 2.8e-6
    out["rgb_values"].sum().backward()                                   # still differentiable, as in the reference (no torch.no_grad around it)
    assert m.rendering_network.lin0.weight_v.grad is not None and bool(torch.isfinite(m.rendering_network.lin0.weight_v.grad).all())
    # and through the module's own sampler (fresh device draws): same keys, finite
    m._tape_override = None
    out2 = m.forward(inp, True)
    assert set(out2) == want and all(bool(torch.isfinite(v).all()) for v in out2.values())
    print(f"{name}: training-mode predict_only on the reference z's: worst rel err {worst:.2e}")
