"""Bidirectional LSTM over variable-length sequences through the persistent CUDA recurrence (csrc/lstm.cu).

Drop-in for `pad_packed_sequence(lstm(pack_padded_sequence(x, lens)))` on an nn.LSTM(num_layers=1,
bidirectional=True, batch_first=True) module (reference radtts.py:284-293, common.py:359-371): same parameters
(including the spectral-norm re-parameterisation hooks), same zero outputs beyond each length, same gradients."""
import ctypes
import os

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


_weight_streams = {}


def _spectral_hooks(lstm):
    """{weight name: hook} when every forward pre-hook of the module is an old-style torch.nn.utils.spectral_norm hook over
    dim 0 (what the reference puts on weight_hh_l0 / weight_hh_l0_reverse), else None."""
    from torch.nn.utils.spectral_norm import SpectralNorm
    hooks = {}
    for h in lstm._forward_pre_hooks.values():
        if not isinstance(h, SpectralNorm) or h.dim != 0:
            return None
        hooks[h.name] = h
    return hooks or None


class _SpectralWeight(torch.autograd.Function):
    """weight = W / sigma, sigma = u^T W v with u, v constants (torch.nn.utils.spectral_norm.compute_weight) -- the value
    comes precomputed from prefetch_weights (side stream, no autograd); this node only carries the gradient:
    dL/dW = G / sigma - <G, W> / sigma^2 * u v^T.  One node on the consuming stream instead of ~25 small autograd nodes."""

    @staticmethod
    def forward(ctx, w_orig, u, v, sigma, w_eff):
        ctx.save_for_backward(w_orig, u, v, sigma)
        return w_eff.view_as(w_eff)

    @staticmethod
    def backward(ctx, g):
        w, u, v, sigma = ctx.saved_tensors
        g = g.to(w.dtype)
        coef = (g * w).sum() / (sigma * sigma)
        return g / sigma - coef * torch.outer(u, v), None, None, None, None


def prefetch_weights(lstms, fp32=()):
    """Starts the spectral-norm work of `lstms` NOW, on a side stream, for the bilstm() calls of the forward pass that is
    starting: the power iteration (training), sigma and W / sigma -- ~30 tiny launches per LSTM that sat on the critical
    path right in front of each recurrence (0.15 ms each in the step's timeline) although they depend on nothing but the
    parameters.  Everything on the side stream runs WITHOUT autograd (nodes recorded on a second stream raced with the
    main stream's backward in testing); the gradient comes from _SpectralWeight at the point of use.  `fp32`: modules
    whose owner runs them with autocast disabled (the text encoder).  The power iteration still runs exactly once per
    forward, as in the reference.  RADTTS.forward calls this first (training on CUDA)."""
    if os.environ.get("RADTTS_NO_LSTM_PREFETCH"):
        return
    todo = []
    for m in lstms:
        if isinstance(m, torch.nn.LSTM) and m.weight_ih_l0.is_cuda and m._forward_pre_hooks:
            hooks = _spectral_hooks(m)
            if hooks is not None:
                todo.append((m, hooks))
    if not todo:
        return
    dev = todo[0][0].weight_ih_l0.device
    side = _weight_streams.get(dev.index)
    if side is None:
        side = _weight_streams[dev.index] = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    F = torch.nn.functional
    with torch.cuda.stream(side), torch.no_grad():
        for m, hooks in todo:
            with torch.autocast("cuda", enabled=torch.is_autocast_enabled() and not any(m is x for x in fp32)):
                state = {}
                for name, h in hooks.items():
                    W = getattr(m, name + "_orig")
                    u, v = getattr(m, name + "_u"), getattr(m, name + "_v")
                    if m.training:
                        for _ in range(h.n_power_iterations):
                            v = F.normalize(torch.mv(W.t(), u), dim=0, eps=h.eps, out=v)
                            u = F.normalize(torch.mv(W, v), dim=0, eps=h.eps, out=u)
                        if h.n_power_iterations > 0:
                            u, v = u.clone(), v.clone()
                    sigma = torch.dot(u, torch.mv(W, v))
                    state[name] = (u, v, sigma, W / sigma)
            ev = torch.cuda.Event()
            ev.record(side)
            m._rb_prefetched = (state, ev)


def drop_prefetched(lstms):
    """Forgets prefetched weights no bilstm() call consumed (a forward that skipped the module); the current stream still
    waits for the side stream's in-place updates of the u / v buffers."""
    for m in lstms:
        pre = getattr(m, "_rb_prefetched", None) if m is not None else None
        if pre is not None:
            torch.cuda.current_stream(m.weight_ih_l0.device).wait_event(pre[1])
            m._rb_prefetched = None


def _effective_weights(lstm):
    """Stacked (w_ih (2,4H,In), w_hh (2,4H,H), bias (2,4H)) with autograd links to the underlying parameters.  The module's
    forward pre-hooks (old-style spectral / weight norm recompute weight_hh_l0*) either run here, or -- spectral norm
    prefetched by prefetch_weights -- are replaced by one _SpectralWeight node per weight."""
    pre = getattr(lstm, "_rb_prefetched", None)
    if pre is not None:
        lstm._rb_prefetched = None
        state, ev = pre
        torch.cuda.current_stream(lstm.weight_ih_l0.device).wait_event(ev)
        for name, (u, v, sigma, w_eff) in state.items():
            setattr(lstm, name, _SpectralWeight.apply(getattr(lstm, name + "_orig"), u, v, sigma, w_eff))
    else:
        for hook in lstm._forward_pre_hooks.values():
            hook(lstm, ())
    w_ih = torch.stack((lstm.weight_ih_l0, lstm.weight_ih_l0_reverse))
    w_hh = torch.stack((lstm.weight_hh_l0, lstm.weight_hh_l0_reverse))
    bias = torch.stack((lstm.bias_ih_l0 + lstm.bias_hh_l0, lstm.bias_ih_l0_reverse + lstm.bias_hh_l0_reverse))
    return w_ih, w_hh, bias


_out_dtype_ok = [True]


def _input_projection(x_tm, w_ih, bias):
    """x (T, B, In) @ w_ih^T + bias for both directions -> (2, T * B, 4H) float32.  With bf16 operands the GEMM writes fp32
    itself (aten::baddbmm.dtype: fp32 bias in the epilogue, no separate up-cast pass over the 0.2 GB result); a torch build
    without that overload falls back to bf16 output + cast, once and for all."""
    T, B, _ = x_tm.shape
    xin = x_tm.reshape(1, T * B, -1).expand(2, -1, -1)
    wt = w_ih.transpose(1, 2).to(x_tm.dtype)
    if x_tm.dtype in (torch.bfloat16, torch.float16) and _out_dtype_ok[0]:
        try:
            with torch.autocast("cuda", enabled=False):
                return torch.baddbmm(bias.float()[:, None, :].expand(-1, T * B, -1), xin, wt, out_dtype=torch.float32)
        except (RuntimeError, TypeError):
            _out_dtype_ok[0] = False
    return torch.baddbmm(bias[:, None, :].to(x_tm.dtype), xin, wt).float()


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
            gx = _input_projection(x_tm, w_ih, bias).reshape(2, T, B, 4 * H).contiguous()
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

    def time_major(xb):
        # ONE pass: batch-first -> time-major and, under autocast, fp32 -> bf16 (the input projection would cast it
        # anyway; done here the transposed copy moves half the bytes and the GEMM's own cast disappears)
        dt = torch.bfloat16 if (xb.is_cuda and torch.is_autocast_enabled() and xb.dtype == torch.float32) else xb.dtype
        if os.environ.get("RADTTS_NO_LSTM_FUSED_CAST"):
            return xb.transpose(0, 1).contiguous()
        return torch.empty((xb.shape[1], xb.shape[0], xb.shape[2]), dtype=dt, device=xb.device).copy_(xb.transpose(0, 1))

    if x.shape[0] <= MAX_B:
        x_tm = time_major(x)
        return _BiLSTMFn.apply(x_tm, lens32, w_ih, w_hh, bias).transpose(0, 1)
    # the kernel keeps <= 32 utterances per launch (one MMA N-tile): larger batches run in chunks, utterances being
    # independent (weight gradients accumulate across the chunks through autograd)
    outs = []
    for lo in range(0, x.shape[0], MAX_B):
        x_tm = time_major(x[lo:lo + MAX_B])
        outs.append(_BiLSTMFn.apply(x_tm, lens32[lo:lo + MAX_B].contiguous(), w_ih, w_hh, bias).transpose(0, 1))
    return torch.cat(outs, 0)
