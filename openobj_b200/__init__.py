"""openobj_b200: B200-native (sm_100a) implementation of OpenObj's vectorised per-object NeRF training path.

The arithmetic lives in hand-written CUDA kernels behind a C ABI (include/openobj_b200.h, built by
`python -m openobj_b200.build`); the Python modules mirror the reference's objnerf/ call surface.
There is no CPU / eager-PyTorch fallback: operations raise if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"
