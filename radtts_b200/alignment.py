"""Monotonic Alignment Search on the GPU -- host mirror of reference alignment.py / radtts.py:320-334.

`binarize_attention` keeps the reference method's argument meaning (attn: B x 1 x T1 x T2 soft attention,
in_lens, out_lens) and returns the dense hard map; the D2H copy, the serial Numba loop and the
per-utterance H2D of the reference are replaced by one CUDA kernel launch (csrc/mas.cu).
"""
import ctypes

import torch

from . import _lib


def _lens(x, device):
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    return x.to(device=device, dtype=torch.int64, non_blocking=True).contiguous()


def mas_forward(attn, in_lens, out_lens, is_prob=True, return_indices=False):
    """attn (B,1,T1,T2) float32 CUDA.  Returns attn_hard (and frame_to_token (B,T1) int32,
    durations (B,T2) int32 when return_indices)."""
    _lib.require_cuda(attn)
    if attn.dim() != 4 or attn.shape[1] != 1 or attn.dtype != torch.float32:
        raise ValueError("attn must be (B,1,T1,T2) float32, got %s %s" % (tuple(attn.shape), attn.dtype))
    attn = attn.detach().contiguous()
    B, _, T1, T2 = attn.shape
    dev = attn.device
    in_lens = _lens(in_lens, dev)
    out_lens = _lens(out_lens, dev)
    hard = torch.empty_like(attn)
    f2t = torch.empty((B, T1), dtype=torch.int32, device=dev) if return_indices else None
    dur = torch.empty((B, T2), dtype=torch.int32, device=dev) if return_indices else None
    L = _lib.lib()
    nws = L.radtts_mas_workspace_bytes(B, T1, T2, int(bool(is_prob)))
    ws = _lib.scratch(dev, nws)
    rc = L.radtts_mas_forward(_lib.ptr(attn), ctypes.c_int(int(bool(is_prob))), _lib.ptr(in_lens),
                              _lib.ptr(out_lens), B, T1, T2, _lib.ptr(hard), _lib.ptr(f2t), _lib.ptr(dur),
                              _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_of(attn))
    _lib.check(rc, "radtts_mas_forward")
    if return_indices:
        return hard, f2t, dur
    return hard


def mas_width1(attn_map):
    """Single utterance (T1,T2) probabilities -> hard map; mirrors reference alignment.mas_width1."""
    t1, t2 = attn_map.shape
    hard = mas_forward(attn_map.reshape(1, 1, t1, t2), [t2], [t1], is_prob=True)
    return hard.reshape(t1, t2)


def binarize_attention(attn, in_lens, out_lens):
    """Mirror of RADTTS.binarize_attention (reference radtts.py:320-334): no gradient flows."""
    with torch.no_grad():
        return mas_forward(attn, in_lens, out_lens, is_prob=True)
