"""Bidirectional LSTM over variable-length sequences through the persistent CUDA recurrence (csrc/lstm.cu).

Drop-in for `pad_packed_sequence(lstm(pack_padded_sequence(x, lens)))` on an nn.LSTM(num_layers=1,
bidirectional=True, batch_first=True) module (reference radtts.py:284-293, common.py:359-371): same parameters
(including the spectral-norm re-parameterisation hooks), same zero outputs beyond each length, same gradients."""
import ctypes

import torch

from . import _lib

MAX_B = 32
MAX_H = 592


class _rnn_matmul_precision:
    """The GEMMs around the recurrence (input projection, weight / input gradients) replace work cuDNN's RNN does
    internally, so in fp32 they follow cuDNN's TF32 switch (torch.backends.cudnn.allow_tf32, default True) rather than
    the matmul one (default False, which sends them to SIMT sgemm kernels)."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(torch.backends.cudnn.allow_tf32)

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


def supported(lstm, x):
    return (x.is_cuda and isinstance(lstm, torch.nn.LSTM) and lstm.bidirectional and lstm.num_layers == 1
            and lstm.hidden_size <= MAX_H and lstm.proj_size == 0)


def _effective_weights(lstm):
    """Runs the module's forward pre-hooks (old-style spectral / weight norm recompute weight_hh_l0*) and returns
    stacked (w_ih (2,4H,In), w_hh (2,4H,H), bias (2,4H)) with autograd links to the underlying parameters."""
    for hook in lstm._forward_pre_hooks.values():
        hook(lstm, ())
    w_ih = torch.stack((lstm.weight_ih_l0, lstm.weight_ih_l0_reverse))
    w_hh = torch.stack((lstm.weight_hh_l0, lstm.weight_hh_l0_reverse))
    bias = torch.stack((lstm.bias_ih_l0 + lstm.bias_hh_l0, lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse))
    return w_ih, w_hh, bias


class _BiLSTMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_tm, lens, w_ih, w_hh, bias):
        T, B, _ = x_tm.shape
        H = w_hh.shape[2]
        dev = x_tm.device
        need_bwd = any(ctx.needs_input_grad)
        # input projection for every step: one GEMM per direction (plain library GEMM; autocast-aware)
        # (bias added in the GEMM's epilogue: a separate add is a 0.4 GB pass over gx at the context LSTM's size)
        with _rnn_matmul_precision():
            xin = x_tm.reshape(1, T * B, -1).expand(2, -1, -1)
            gx = torch.baddbmm(bias[:, None, :].to(x_tm.dtype), xin, w_ih.transpose(1, 2).to(x_tm.dtype))
            gx = gx.float().reshape(2, T, B, 4 * H).contiguous()
        whh = w_hh.detach().float().contiguous()
        h_all = torch.empty((T, B, 2 * H), dtype=torch.float32, device=dev)
        gates = torch.empty((2, T, B, 4 * H), dtype=torch.float32, device=dev) if need_bwd else None
        cs = torch.empty((2, T, B, H), dtype=torch.float32, device=dev) if need_bwd else None
        L = _lib.lib()
        nws = int(L.radtts_lstm_workspace_bytes(B, H))
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        # under autocast (where cuDNN would run the LSTM in half precision) the recurrent product uses bf16 operands
        # on tensor cores; state, gates and outputs stay fp32.  The backward pass follows the forward's precision.
        low_prec = torch.is_autocast_enabled()
        # recurrence precision: bf16 operands under autocast; in fp32, where cuDNN's RNN would be allowed TF32
        # (torch.backends.cudnn.allow_tf32, the default), split-bf16 operands (16 mantissa bits) on tensor cores;
        # the exact fp32 FMA kernel otherwise
        prec = 1 if low_prec else (2 if torch.backends.cudnn.allow_tf32 else 0)
        _lib.check(L.radtts_lstm_forward(_lib.ptr(gx), _lib.ptr(whh), _lib.ptr(lens), T, B, H, _lib.ptr(h_all),
                                         _lib.ptr(gates), _lib.ptr(cs), _lib.ptr(ws), ctypes.c_size_t(nws),
                                         prec, _lib.stream_of(x_tm)), "radtts_lstm_forward")
        if need_bwd:
            ctx.save_for_backward(x_tm, lens, w_ih, whh, gates, cs, h_all)
            ctx.low_prec = low_prec
            ctx.prec = prec
        return h_all

    @staticmethod
    def backward(ctx, dh_all):
        x_tm, lens, w_ih, whh, gates, cs, h_all = ctx.saved_tensors
        T, B, _ = x_tm.shape
        H = whh.shape[2]
        dev = x_tm.device
        dh_all = dh_all.float().contiguous()
        dg = torch.empty((2, T, B, 4 * H), dtype=torch.float32, device=dev)
        L = _lib.lib()
        nws = int(L.radtts_lstm_workspace_bytes(B, H))
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        _lib.check(L.radtts_lstm_backward(_lib.ptr(dh_all), _lib.ptr(whh), _lib.ptr(lens), _lib.ptr(gates), _lib.ptr(cs),
                                          T, B, H, _lib.ptr(dg), _lib.ptr(ws), ctypes.c_size_t(nws),
                                          ctx.prec, _lib.stream_of(x_tm)), "radtts_lstm_backward")
        mm = torch.bfloat16 if ctx.low_prec else torch.float32   # plain library GEMMs; bf16 under autocast
        dg2 = dg.reshape(2, T * B, 4 * H)
        dgm = dg2.to(mm)
        xf = x_tm.reshape(T * B, -1).to(mm)
        # h_{t-1} in each direction's own time order: the first (forward) / last (reverse) step multiplies a zero state,
        # so those B rows are simply left out of the product -- no zero-padded copies of h
        h16 = h_all.to(mm)
        h_prev_f = h16[:-1, :, :H].reshape((T - 1) * B, H)
        h_prev_r = h16[1:, :, H:].reshape((T - 1) * B, H)
        with _rnn_matmul_precision():
            wm = w_ih.to(mm)
            d_x = torch.addmm(dgm[0] @ wm[0], dgm[1], wm[1]).reshape(x_tm.shape).to(x_tm.dtype)   # sum over directions
            d_w_ih = torch.matmul(dgm.transpose(1, 2), xf).float()
            d_w_hh = torch.stack((dgm[0][B:].t() @ h_prev_f, dgm[1][:(T - 1) * B].t() @ h_prev_r)).float()
        d_bias = dg2.sum(1)
        return d_x, None, d_w_ih.to(w_ih.dtype), d_w_hh, d_bias


def bilstm(lstm, x, lens):
    """x (B, T, In) batch-first, lens (B,) -> (B, T, 2H) with zeros beyond each length."""
    w_ih, w_hh, bias = _effective_weights(lstm)
    lens32 = lens.to(device=x.device, dtype=torch.int32).contiguous()
    if x.shape[0] <= MAX_B:
        x_tm = x.transpose(0, 1).contiguous()
        return _BiLSTMFn.apply(x_tm, lens32, w_ih, w_hh, bias).transpose(0, 1)
    # the kernel keeps <= 32 utterances per launch (one MMA N-tile): larger batches run in chunks, utterances being
    # independent (weight gradients accumulate across the chunks through autograd)
    outs = []
    for lo in range(0, x.shape[0], MAX_B):
        x_tm = x[lo:lo + MAX_B].transpose(0, 1).contiguous()
        outs.append(_BiLSTMFn.apply(x_tm, lens32[lo:lo + MAX_B].contiguous(), w_ih, w_hh, bias).transpose(0, 1))
    return torch.cat(outs, 0)
