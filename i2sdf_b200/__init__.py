"""i2sdf_b200 — B200-native (sm_100a) volume-rendering core for the I2-SDF per-ray hot path.

Drop-in for `model.network` of jingsenzhu/i2-sdf:
    from i2sdf_b200.network import I2SDFNetwork, I2SDFLoss
The compute runs in hand-written CUDA kernels behind a C-ABI shared library (include/i2sdf_b200.h);
there is no CPU fallback — calling into the renderer without the built library / a GPU raises.
"""
__version__ = "0.1.0"
