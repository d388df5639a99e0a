"""ctypes binding of the C-ABI library (include/radtts_b200.h).

There is NO CPU fallback: if the CUDA library is missing or a call fails, this module raises.
"""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libradtts_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "radtts_b200.h")

_lib = None


class RadttsB200Error(RuntimeError):
    pass


def declared_symbols():
    """Names of every function the public header declares (used by the symbol-export test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return sorted(set(re.findall(r"RADTTS_API[^;(]*?\b(radtts_[a-z0-9_]+)\s*\(", text)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RadttsB200Error(
                "CUDA extension %s is missing: run `python -m radtts_b200.build` (there is no CPU fallback)"
                % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.radtts_error_string.restype = ctypes.c_char_p
        _lib.radtts_error_string.argtypes = [ctypes.c_int]
        _lib.radtts_launch_count.restype = ctypes.c_longlong
        _lib.radtts_abi_version.restype = ctypes.c_int
        for name in declared_symbols():
            fn = getattr(_lib, name)
            if name.endswith("_bytes"):
                fn.restype = ctypes.c_size_t
    return _lib


def check(code, what=""):
    if code != 0:
        msg = lib().radtts_error_string(int(code)).decode()
        raise RadttsB200Error("%s failed: %s (code %d)" % (what or "radtts_b200 call", msg, code))


def launch_count():
    return int(lib().radtts_launch_count())


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RadttsB200Error("radtts_b200 ops run on CUDA tensors only (got %s); no CPU fallback" % t.device)


def scratch(device, nbytes):
    """Per-call scratch buffer (uint8) from torch's stream-ordered caching allocator.  Every user gets its OWN
    allocation: kernels on different streams (MAS on the main stream, the prefetched CTC on a side stream) never share
    one, and a CUDA-graph capture keeps its scratch alive in the graph's private pool."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
