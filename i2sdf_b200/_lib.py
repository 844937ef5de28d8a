"""ctypes binding of libi2sdf_b200.so (C ABI in include/i2sdf_b200.h).  No torch types cross this boundary."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libi2sdf_b200.so")
# test infrastructure: the same ABI with the fp32 SIMT kernel + layer-by-layer backward as a selectable cross-check backend
# (I2SDF_SIMT=1 / I2SDF_SIMT_MAIN=1 / I2SDF_FUSED_BWD=0 at RenderCore creation); the product library ignores those switches
CHECK_LIB_PATH = os.path.join(_HERE, "libi2sdf_b200_check.so")
ABI_VERSION = 1


class Desc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "hidden", "feature_size", "n_sdf_layers", "sdf_skip_layer", "multires_x", "n_color_layers",
        "multires_d", "n_light_layers", "light_hidden", "n_samples", "n_samples_eval", "n_samples_extra",
        "beta_iters", "max_total_iters")] + [(n, C.c_float) for n in (
        "near_", "far_", "eps", "add_tiny", "beta_min", "lemma2_coeff")] + [
        ("u_up", C.POINTER(C.c_float)), ("u_final", C.POINTER(C.c_float)), ("t_init", C.POINTER(C.c_float)),
        ("extra_idx", C.POINTER(C.c_int32))]


class LossArgs(C.Structure):
    """i2sdf_loss_args (include/i2sdf_b200.h)."""
    _fields_ = [("R", C.c_int64), ("n_eik", C.c_int64), ("n_bubble", C.c_int64)] + [(n, C.c_void_p) for n in (
        "rgb", "rgb_gt", "grad_theta", "diff_norm", "weight_sum", "mask_gt", "depth", "depth_gt", "depth_mask", "normal",
        "normal_gt", "normal_mask", "surface_sdf", "light", "light_gt")] + [(n, C.c_float) for n in (
        "w_eik", "w_smooth", "w_mask", "w_depth", "w_normal", "w_angular", "w_bubble", "w_light")] + [(n, C.c_void_p) for n in (
        "terms", "g_rgb", "g_grad_theta", "g_diff_norm", "g_weight_sum", "g_depth", "g_normal", "g_surface_sdf", "g_light", "denom")]


WNORM_MAX_JOBS = 28


class WnormJob(C.Structure):
    """i2sdf_wnorm_job (include/i2sdf_b200.h)."""
    _fields_ = [(n, C.c_void_p) for n in ("g", "v", "W", "norm", "dW", "dg", "dv")] + [("rows", C.c_int32), ("cols", C.c_int32)]


class WnormBatch(C.Structure):
    _fields_ = [("n", C.c_int32), ("pad_", C.c_int32), ("jobs", WnormJob * WNORM_MAX_JOBS)]


ADAM_MAX_JOBS = 64


class AdamJob(C.Structure):
    """i2sdf_adam_job (include/i2sdf_b200.h)."""
    _fields_ = [(n, C.c_void_p) for n in ("param", "grad", "exp_avg", "exp_avg_sq")] + [("numel", C.c_int64)]


class AdamBatch(C.Structure):
    _fields_ = [("n", C.c_int32)] + [(n, C.c_float) for n in ("beta1", "beta2", "eps", "step_size", "bias_correction2_sqrt",
                                                                "one_minus_beta1", "one_minus_beta2")] + [("jobs", AdamJob * ADAM_MAX_JOBS)]


# every symbol include/i2sdf_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "i2sdf_abi_version": (C.c_int, []),
    "i2sdf_last_error": (C.c_char_p, []),
    "i2sdf_create": (C.c_int, [C.POINTER(Desc), C.c_int, C.POINTER(_P)]),
    "i2sdf_destroy": (C.c_int, [_P]),
    "i2sdf_num_layers": (C.c_int, [_P]),
    "i2sdf_uses_tensor_cores": (C.c_int, [_P]),
    "i2sdf_pack_weights": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), _P]),
    "i2sdf_workspace_bytes": (C.c_size_t, [_P, C.c_int64, C.c_int]),
    "i2sdf_rays": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "i2sdf_sdf_forward": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "i2sdf_sdf_grid": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "i2sdf_sampler_rounds": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, _P, C.c_size_t, _P]),
    "i2sdf_sampler_step": (C.c_int, [_P, _P, _P, C.c_int64, _P, _P, _P, C.c_int, C.c_int, _P, C.c_size_t, _P]),
    "i2sdf_sampler_beta_max": (C.c_void_p, [_P, C.c_int64, _P]),
    "i2sdf_sampler_finalize": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "i2sdf_sampler_finalize_candidates": (C.c_int, [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "i2sdf_sampler_info": (C.c_int, [_P, C.c_int64, _P, _P, _P, C.c_size_t, _P]),
    "i2sdf_sampler_round_debug": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int, _P, _P, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P]),
    "i2sdf_render_forward": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int, _P] + [_P] * 10 + [_P, C.c_size_t, _P, C.c_size_t, _P]),
    "i2sdf_saved_bytes": (C.c_size_t, [_P, C.c_int64, C.c_int]),
    "i2sdf_backward_workspace_bytes": (C.c_size_t, [_P, C.c_int64]),
    "i2sdf_points_forward": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int] + [_P] * 6 + [_P, C.c_size_t, _P]),
    "i2sdf_composite_forward": (C.c_int, [_P] + [_P] * 7 + [C.c_int64, C.c_int] + [_P] * 6 + [_P]),
    "i2sdf_composite_backward": (C.c_int, [_P] + [_P] * 7 + [C.c_int64, C.c_int] + [_P] * 10 + [_P]),
    "i2sdf_color_backward": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), _P, C.c_int, _P, _P, _P, C.c_int64, C.POINTER(_P), C.POINTER(_P), _P, _P, C.c_size_t, _P]),
    "i2sdf_light_backward": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), _P, _P, _P, _P, C.c_int64, C.POINTER(_P), C.POINTER(_P), _P, C.c_size_t, _P]),
    "i2sdf_sdf_backward": (C.c_int, [_P, C.POINTER(_P), _P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, _P, _P, _P, C.c_int, _P, C.POINTER(_P), C.POINTER(_P), _P, C.c_size_t, _P]),
    "i2sdf_profile_enable": (C.c_int, [_P, C.c_int]),
    "i2sdf_profile_read": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "i2sdf_profile_read_n": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "i2sdf_debug_bwd_timeline": (C.c_int64, [C.POINTER(C.c_int64), C.c_int64]),
    "i2sdf_saved_format": (C.c_int, [_P, C.c_int]),
    "i2sdf_sdf_saved_bytes": (C.c_size_t, [_P, C.c_int64]),
    "i2sdf_points_forward_ex": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int, _P, C.c_int64] + [_P] * 6 + [_P, C.c_size_t, _P]),
    "i2sdf_saved_bytes_points": (C.c_size_t, [_P, C.c_int64]),
    "i2sdf_fused_backward": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.POINTER(_P), C.POINTER(_P),
                                       C.POINTER(_P), C.POINTER(_P), _P, C.c_size_t, _P]),
    "i2sdf_fused_backward_ex": (C.c_int, [_P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P, _P, _P, _P, C.c_int64, _P, C.POINTER(_P), C.POINTER(_P),
                                          C.POINTER(_P), C.POINTER(_P), _P, C.c_size_t, _P]),
    "i2sdf_loss_forward": (C.c_int, [C.POINTER(LossArgs), _P]),
    "i2sdf_weight_norm": (C.c_int, [C.POINTER(WnormBatch), C.c_int, _P]),
    "i2sdf_adam_step": (C.c_int, [C.POINTER(AdamBatch), _P]),
    "i2sdf_adam_step_dev": (C.c_int, [C.POINTER(AdamBatch), _P, _P]),
    "i2sdf_planes_slot_bytes": (C.c_size_t, [C.c_int64, C.c_int]),
    "i2sdf_planes_pack": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int64, C.c_int, _P, _P]),
    "i2sdf_planes_unpack": (C.c_int, [_P, _P, C.c_int, C.c_int64, _P, C.c_int, C.c_int, _P]),
    "i2sdf_planes_wgrad": (C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(_P), C.c_int, C.c_int64, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
}

_lib = None
_check_lib = None


def wants_check_backend() -> bool:
    return os.environ.get("I2SDF_SIMT") == "1" or os.environ.get("I2SDF_SIMT_MAIN") == "1" or os.environ.get("I2SDF_FUSED_BWD") == "0"


class I2SDFError(RuntimeError):
    pass


def _open(path):
    if not os.path.exists(path):
        raise I2SDFError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or i2sdf_b200/csrc/build.sh).  i2sdf_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.i2sdf_abi_version() != ABI_VERSION:
        raise I2SDFError(f"ABI mismatch: library {lib.i2sdf_abi_version()} vs binding {ABI_VERSION}")
    return lib


def load():
    """Load the CUDA library.  Fails loudly: there is no CPU fallback."""
    global _lib
    if _lib is None:
        _lib = _open(LIB_PATH)
    return _lib


def load_check():
    """The check build (tests / tools only): product kernels + the fp32 cross-check backend behind the I2SDF_SIMT* switches."""
    global _check_lib
    if _check_lib is None:
        _check_lib = _open(CHECK_LIB_PATH)
    return _check_lib


def check(rc, what=""):
    if rc != 0:
        msg = load().i2sdf_last_error().decode(errors="replace")
        if not msg and _check_lib is not None:
            msg = _check_lib.i2sdf_last_error().decode(errors="replace")
        raise I2SDFError(f"{what} failed (code {rc}): {msg}")
