"""Import the UNMODIFIED reference (jingsenzhu/i2-sdf) read-only from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (in the build container, where
/root/reference exists) to generate golden vectors.  Nothing on the product path imports this.

The reference imports a handful of third-party modules at import time that are absent from this
image and are not used by the hot-path arithmetic; they are replaced by empty
stub modules.  `model/__init__.py` (which pulls the Lightning trainer) is bypassed by registering
`model` as a namespace package.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("I2SDF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model", "network"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load():
    """Returns (ref_network_module, ref_utils_module)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    class _Dummy:  # stands in for RichProgressBar / KMeans
        def __init__(self, *a, **k):
            pass

    try:
        import pytorch_lightning  # noqa: F401
    except Exception:
        pl = _stub("pytorch_lightning")
        cb = _stub("pytorch_lightning.callbacks", RichProgressBar=_Dummy)
        pl.callbacks = cb
    for name in ("imageio", "skimage"):
        try:
            importlib.import_module(name)
        except Exception:
            _stub(name)
    try:
        import fast_pytorch_kmeans  # noqa: F401
    except Exception:
        _stub("fast_pytorch_kmeans", KMeans=_Dummy)

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    ref_utils = importlib.import_module("utils")
    if "model" not in sys.modules:
        pkg = types.ModuleType("model")
        pkg.__path__ = [os.path.join(REF_ROOT, "model")]
        sys.modules["model"] = pkg
    ref_net = importlib.import_module("model.network")
    return ref_net, ref_utils


def load_conf(name="synthetic.yml"):
    """The `model:` node of a reference config as the reference's own CfgNode."""
    import yaml
    _, ref_utils = load()
    with open(os.path.join(REF_ROOT, "config", name)) as f:
        d = yaml.safe_load(f)
    return ref_utils.CfgNode(d)
