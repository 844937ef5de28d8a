"""Loading of tests/golden fixtures (written by tests/golden/make_golden.py from the reference)."""
import os

import numpy as np
import torch

from i2sdf_b200 import configs
from oracle import i2sdf_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EVAL_CASES = ["eval_synthetic_sharp", "eval_synthetic_soft", "eval_light_sharp"]
TRAIN_CASES = ["train_synthetic", "train_light"]


def _t(a):
    a = np.asarray(a)
    return torch.from_numpy(a.copy()) if a.shape else torch.tensor(a.item())


class Case:
    def __init__(self, name):
        self.name = name
        d = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.raw = d
        self.conf_name = str(d["meta_conf"])
        self.training = name.startswith("train")
        w = np.load(os.path.join(GOLDEN, f"weights_{str(d['meta_weights'])}.npz"))
        self.params = {k: _t(w[k]) for k in w.files}
        self.params["density.beta"] = torch.tensor(float(d["meta_beta"]))
        self.model_conf = configs.model_conf(self.conf_name)
        self.spec = orc.spec_from_model_conf(self.model_conf, use_normal=self.training)
        self.inputs = {k[3:]: _t(d[k]) for k in d.files if k.startswith("in_")}
        self.gt = {k[3:]: _t(d[k]) for k in d.files if k.startswith("gt_")}
        self.tape = {k[5:]: _t(d[k]) for k in d.files if k.startswith("tape_")}
        self.ref = {k[4:]: _t(d[k]) for k in d.files if k.startswith("ref_") and not k.startswith("ref_mid_")
                    and k != "ref_loss"}
        self.mid = {k[8:]: _t(d[k]) for k in d.files if k.startswith("ref_mid_")}
        self.trace = {k[6:]: _t(d[k]) for k in d.files if k.startswith("trace_")}
        self.ref_loss = float(d["ref_loss"]) if "ref_loss" in d.files else None
        self.loss_conf = dict(configs.LOSS_SYNTHETIC_LIGHT_MASK if self.spec.light_dims else configs.LOSS_SYNTHETIC)

    def loss_kwargs(self):
        keys = ("eikonal_weight", "smooth_weight", "depth_weight", "normal_weight", "bubble_weight",
                "light_mask_weight")
        return {k: v for k, v in self.loss_conf.items() if k in keys}

    def refgrads(self):
        d = self.raw
        out = {}
        for k in d.files:
            for kind in ("full", "norm", "sum", "sub"):
                p = f"refgrad_{kind}_"
                if k.startswith(p):
                    out.setdefault(k[len(p):], {})[kind] = _t(d[k])
        return out


def relerr(a, b):
    """max |a-b| / max |b|  — the parity metric used throughout (north_star: 1e-4 relative fp32)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def general_cameras(B, g):
    """B cameras with a proper rotation (not the identity of the synthetic batches), a translation, per-camera focal lengths,
    principal points and a NON-ZERO skew: every term of rend_util.py:143-144 is exercised.  -> pose [B,4,4], intrinsics [B,4,4]."""
    q = torch.nn.functional.normalize(torch.randn(B, 4, generator=g), dim=-1)
    w, x, y, z = q.unbind(-1)
    Rm = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                      2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                      2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(B, 3, 3)
    pose = torch.eye(4).repeat(B, 1, 1)
    pose[:, :3, :3] = Rm
    pose[:, :3, 3] = torch.randn(B, 3, generator=g) * 2.0
    K = torch.eye(4).repeat(B, 1, 1)
    K[:, 0, 0] = 250.0 + 100.0 * torch.rand(B, generator=g)
    K[:, 1, 1] = 250.0 + 100.0 * torch.rand(B, generator=g)
    K[:, 0, 2] = 150.0 + 20.0 * torch.rand(B, generator=g)
    K[:, 1, 2] = 110.0 + 20.0 * torch.rand(B, generator=g)
    K[:, 0, 1] = torch.randn(B, generator=g) * 0.5
    return pose, K
