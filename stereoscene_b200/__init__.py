"""stereoscene_b200 -- B200-native volumetric hot path of StereoScene / BRGScene.

Layout: csrc/ (CUDA kernels + the C ABI of include/stereoscene_b200.h), cabi.py (ctypes binding),
ops.py (tensor-level wrappers), plugin/ (the reference's registered modules re-hosted on the
kernels), config.py / registry.py (mmcv-compatible config + registry surface), synth.py
(seeded synthetic inputs).
"""
__version__ = "0.1.0"
