"""radtts_b200 -- B200-native (sm_100a) implementation of the RADTTS alignment + decoder-flow hot path.

Host code is Python/PyTorch mirroring the reference module API; compute runs in hand-written CUDA
kernels behind the C ABI in include/radtts_b200.h (loaded with ctypes, see _lib.py).
"""
__version__ = "0.1.0"
