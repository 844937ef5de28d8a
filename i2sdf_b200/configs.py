"""The `model:` / `loss:` nodes of the reference's two shipped configs, as plain dicts.

Values are the ones in jingsenzhu/i2-sdf config/synthetic.yml:13-23,30-74 and
config/synthetic_light_mask.yml (same lines; differences: 6x256 SDF stack with skip_in [3], 3x256 radiance
stack, a [128] light-mask head and light_mask_weight 0.5).  They pin every size on the hot path
 and are what bench.py / the tests build networks from when the reference tree is absent.
"""
import copy

_SAMPLER = dict(near=0.0, N_samples=64, N_samples_eval=128, N_samples_extra=32, eps=0.1, beta_iters=10,
                max_total_iters=5, N_samples_inverse_sphere=32, add_tiny=1.0e-6)

SYNTHETIC = dict(
    feature_vector_size=256,
    scene_bounding_sphere=3.0,
    implicit_network=dict(d_in=3, d_out=1, dims=[256] * 8, geometric_init=True, bias=0.6, skip_in=[4],
                          weight_norm=True, embed_type="positional", multires=6),
    rendering_network=dict(mode="nerf", d_in=3, d_out=3, dims=[256] * 4, weight_norm=True,
                           embed_type="positional", multires=4),
    density=dict(params_init=dict(beta=0.1), beta_min=0.0001),
    ray_sampler=dict(_SAMPLER),
)

SYNTHETIC_LIGHT_MASK = dict(
    feature_vector_size=256,
    scene_bounding_sphere=3.0,
    implicit_network=dict(d_in=3, d_out=1, dims=[256] * 6, geometric_init=True, bias=0.6, skip_in=[3],
                          weight_norm=True, embed_type="positional", multires=6),
    rendering_network=dict(mode="nerf", d_in=3, d_out=3, dims=[256] * 3, weight_norm=True,
                           embed_type="positional", multires=4),
    light_network=dict(dims=[128], weight_norm=True),
    density=dict(params_init=dict(beta=0.1), beta_min=0.0001),
    ray_sampler=dict(_SAMPLER),
)

LOSS_SYNTHETIC = dict(eikonal_weight=0.1, smooth_weight=0.01, smooth_iter=150000, depth_weight=0.1,
                      normal_weight=0.05, bubble_weight=0.5, min_bubble_iter=50000, max_bubble_iter=150000)
LOSS_SYNTHETIC_LIGHT_MASK = dict(LOSS_SYNTHETIC, light_mask_weight=0.5)

MODEL_CONFIGS = {"synthetic": SYNTHETIC, "synthetic_light_mask": SYNTHETIC_LIGHT_MASK}


def model_conf(name: str) -> dict:
    return copy.deepcopy(MODEL_CONFIGS[name])
