"""
tamm_b200 -- B200-native (sm_100a) layer-stack transfer-matrix solver behind the API of nicholasmr/tamm.

Host code is Python; all arithmetic of the hot path runs in hand-written CUDA kernels reached through the C ABI
of `libtamm_b200.so` (include/tamm_b200.h) with PyTorch tensors used only as device buffers.  There is no CPU
fallback: importing works anywhere, computing requires the built library and a B200.
"""
from . import packing  # noqa: F401
from .engine import Engine, StackResult, get_engine  # noqa: F401
from .ensemble import EnsembleStack, shard_range  # noqa: F401

__version__ = '0.1.0'
