"""Host-side mirror of the reference layer API (reference common.py / partialconv1d.py).

Module names, constructor arguments, parameter names and shapes follow the reference so that reference
checkpoints load unchanged.  The layers on the hot path (ConvAttention, Invertible1x1ConvLUS, WN,
AffineTransformationLayer, SimpleConvNet, SplineTransformationLayer) do their math in the CUDA library
(see ops.py); the remaining layers (text Encoder, ConvLSTMLinear, LengthRegulator ...) are out of the hot-path
scope and are ordinary PyTorch.
"""
import ast

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from . import lstm_ops, ops


def update_params(config, params):
    """`-p a.b.c=value` overrides (reference common.py:65-83)."""
    for param in params:
        key, _, raw = param.partition("=")
        try:
            value = ast.literal_eval(raw)
        except Exception:
            value = raw
        node = config
        parts = key.split(".")
        for p in parts[:-1]:
            node = node[p]
        if parts[-1] in node:
            node[parts[-1]] = value
        else:
            print("%s, %s params not updated" % (key, value))


def get_mask_from_lengths(lengths, max_len=None):
    """(B,) lengths -> (B, max_len) bool mask, True inside the sequence (reference common.py:86-97).
    Works on any device; pass max_len to avoid the device->host sync the reference always pays."""
    if max_len is None:
        max_len = int(lengths.max())
    ids = torch.arange(max_len, device=lengths.device)
    return ids[None, :] < lengths[:, None]


class ExponentialClass(nn.Module):
    def forward(self, x):
        return torch.exp(x)


class LinearNorm(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, w_init_gain="linear"):
        super().__init__()
        self.linear_layer = nn.Linear(in_dim, out_dim, bias=bias)
        nn.init.xavier_uniform_(self.linear_layer.weight, gain=nn.init.calculate_gain(w_init_gain))

    def forward(self, x):
        return self.linear_layer(x)


class PartialConv1d(nn.Conv1d):
    """Mask-renormalised conv (reference partialconv1d.py:20-71), PyTorch version used by the layers that
    are outside the hot path (text encoder).  The decoder's WN in_layers share these parameter names but
    run through the fused CUDA kernels instead."""

    def forward(self, x, mask_in=None):
        k = self.kernel_size[0]
        if mask_in is None:
            mask_in = torch.ones(1, 1, x.shape[2], dtype=x.dtype, device=x.device)
            xin = x
        else:
            xin = x * mask_in
        with torch.no_grad():
            ones = torch.ones(1, 1, k, dtype=x.dtype, device=x.device)
            count = F.conv1d(mask_in, ones, None, self.stride, self.padding, self.dilation)
            seen = count.clamp(0, 1)
            ratio = (k / (count + 1e-6)) * seen
        raw = F.conv1d(xin, self.weight, self.bias, self.stride, self.padding, self.dilation)
        if self.bias is None:
            return raw * ratio
        b = self.bias.view(1, -1, 1)
        return ((raw - b) * ratio + b) * seen


class ConvNorm(nn.Module):
    """Conv1d wrapper with the reference's attribute layout (`.conv`, optional weight norm) -- common.py:121-154."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear", use_partial_padding=False, use_weight_norm=False):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = int(dilation * (kernel_size - 1) / 2)
        self.kernel_size = kernel_size
        self.dilation = dilation
        self.use_partial_padding = use_partial_padding
        self.use_weight_norm = use_weight_norm
        cls = PartialConv1d if use_partial_padding else nn.Conv1d
        self.conv = cls(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding,
                        dilation=dilation, bias=bias)
        nn.init.xavier_uniform_(self.conv.weight, gain=nn.init.calculate_gain(w_init_gain))
        if use_weight_norm:
            self.conv = nn.utils.weight_norm(self.conv)

    def forward(self, signal, mask=None):
        y = self.conv(signal, mask) if self.use_partial_padding else self.conv(signal)
        if mask is not None:
            y = y * mask
        return y


class DenseLayer(nn.Module):
    def __init__(self, in_dim=1024, sizes=(1024, 1024)):
        super().__init__()
        dims = [in_dim] + list(sizes)
        self.layers = nn.ModuleList([LinearNorm(a, b, bias=True) for a, b in zip(dims[:-1], dims[1:])])

    def forward(self, x):
        for layer in self.layers:
            x = torch.tanh(layer(x))
        return x


class LengthRegulator(nn.Module):
    """Token -> frame expansion (reference common.py:171-200).  The reference loops over tokens in Python; here frame t
    of utterance b takes token searchsorted(cumsum(dur[b]), t) -- one batched index computation and one gather, with a
    single host read (the longest output, which fixes the result's shape)."""

    def forward(self, x, dur):
        # x (B, N, C), dur (B, N); expanded_len = int(dur + 0.5)
        reps = (dur.float() + 0.5).floor().long().clamp(min=0)
        ends = torch.cumsum(reps, 1)                                   # (B, N) exclusive end frame of every token
        total = ends[:, -1]
        max_len = int(total.max())
        t = torch.arange(max_len, device=x.device)[None, :].expand(x.shape[0], -1).contiguous()
        idx = torch.searchsorted(ends, t, right=True).clamp(max=x.shape[1] - 1)   # (B, max_len) token of every frame
        out = torch.gather(x, 1, idx[:, :, None].expand(-1, -1, x.shape[2]))
        return out * (t < total[:, None])[:, :, None].to(out.dtype)


def _apply_lstm_norm(lstm, kind):
    if kind is None:
        return lstm
    if "spectral" in kind:
        fn = nn.utils.spectral_norm
    elif "weight" in kind:
        fn = nn.utils.weight_norm
    else:
        return lstm
    lstm = fn(lstm, "weight_hh_l0")
    return fn(lstm, "weight_hh_l0_reverse")


class ConvLSTMLinear(nn.Module):
    """Deterministic attribute predictor trunk (reference common.py:203-302); outside the hot path."""

    def __init__(self, in_dim, out_dim, n_layers=2, n_channels=256, kernel_size=3, p_dropout=0.1,
                 lstm_type="bilstm", use_linear=True):
        super().__init__()
        self.out_dim = out_dim
        self.lstm_type = lstm_type
        self.use_linear = use_linear
        self.dropout = nn.Dropout(p=p_dropout)
        convs = []
        for i in range(n_layers):
            layer = ConvNorm(in_dim if i == 0 else n_channels, n_channels, kernel_size=kernel_size, stride=1,
                             padding=int((kernel_size - 1) / 2), dilation=1, w_init_gain="relu")
            convs.append(nn.utils.weight_norm(layer.conv, name="weight"))
        self.convolutions = nn.ModuleList(convs)
        if not use_linear:
            n_channels = out_dim
        if lstm_type != "":
            bi = lstm_type == "bilstm"
            hidden = n_channels // 2 if bi else n_channels
            self.bilstm = nn.LSTM(n_channels, hidden, 1, batch_first=True, bidirectional=bi)
            self.bilstm = nn.utils.spectral_norm(self.bilstm, "weight_hh_l0")
            if bi:
                self.bilstm = nn.utils.spectral_norm(self.bilstm, "weight_hh_l0_reverse")
        if use_linear:
            self.dense = nn.Linear(n_channels, out_dim)

    def _convs(self, x):
        for conv in self.convolutions:
            x = self.dropout(F.relu(conv(x)))
        return x

    def _convs_masked(self, x, mask):
        """All utterances at once: the reference crops every utterance to its length and runs the (zero-padded) conv
        stack on it alone (common.py:246-255).  Zeroing the input beyond each length in front of every conv is the same
        arithmetic on the valid frames -- no Python loop over the batch, no `int(lens[b])` host read."""
        for conv in self.convolutions:
            x = self.dropout(F.relu(conv(x * mask)))
        return x * mask

    def forward(self, context, lens):
        """context (B, C, T), lens (B,) or None (every utterance has T frames).  Output (B, out_dim, T) -- the reference
        returns max(lens) frames for B > 1, which equals T for every batch its DataCollate builds."""
        B, _, T = context.shape
        if B > 1 and lens is not None:
            mask = get_mask_from_lengths(lens.to(context.device), T)[:, None].to(context.dtype)
            context = self._convs_masked(context, mask)
        else:
            context = self._convs(context)
        if self.lstm_type != "":
            x = context.transpose(1, 2)
            full = lens if lens is not None else torch.full((B,), T, dtype=torch.int64, device=context.device)
            if lstm_ops.supported(self.bilstm, x):
                x = lstm_ops.bilstm(self.bilstm, x, full)
            elif lens is not None:
                x = _cudnn_packed_lstm(self.bilstm, x, lens)
            else:
                x = self.bilstm(x)[0]
            context = x.transpose(1, 2)
        if self.use_linear:
            context = self.dense(context.transpose(1, 2)).transpose(1, 2)
        return context


def _cudnn_packed_lstm(lstm, x, lens):
    """The reference's packed-sequence cuDNN call (common.py:257-275), for LSTMs the persistent kernel does not cover
    (unidirectional, hidden size > 592).  Costs a device->host read of the lengths."""
    packed = nn.utils.rnn.pack_padded_sequence(x, lens.long().cpu(), batch_first=True, enforce_sorted=False)
    return nn.utils.rnn.pad_packed_sequence(lstm(packed)[0], batch_first=True, total_length=x.shape[1])[0]


def run_bilstm(lstm, x, lens):
    """pad_packed(lstm(pack_padded(x, lens))) for a batch-first BiLSTM on the persistent CUDA recurrence (csrc/lstm.cu;
    batches larger than its 32-utterance limit run in chunks -- utterances are independent).  Only a shape the kernel
    cannot take at all (hidden size > 592, CPU tensors) goes to the cuDNN packed-sequence path the reference uses."""
    if lstm_ops.supported(lstm, x):
        return lstm_ops.bilstm(lstm, x, lens)
    return _cudnn_packed_lstm(lstm, x, lens)


class Encoder(nn.Module):
    """Text encoder: 3 x (partial conv k5 + InstanceNorm + ReLU + dropout) + BiLSTM (reference
    common.py:305-384); outside the hot path, kept as PyTorch / cuDNN."""

    def __init__(self, encoder_n_convolutions=3, encoder_embedding_dim=512, encoder_kernel_size=5,
                 norm_fn=nn.BatchNorm1d, lstm_norm_fn=None):
        super().__init__()
        self.convolutions = nn.ModuleList([
            nn.Sequential(
                ConvNorm(encoder_embedding_dim, encoder_embedding_dim, kernel_size=encoder_kernel_size, stride=1,
                         padding=int((encoder_kernel_size - 1) / 2), dilation=1, w_init_gain="relu",
                         use_partial_padding=True),
                norm_fn(encoder_embedding_dim, affine=True))
            for _ in range(encoder_n_convolutions)])
        self.lstm = nn.LSTM(encoder_embedding_dim, int(encoder_embedding_dim / 2), 1, batch_first=True,
                            bidirectional=True)
        self.lstm = _apply_lstm_norm(self.lstm, lstm_norm_fn)

    def _convs(self, x):
        for conv in self.convolutions:
            x = F.dropout(F.relu(conv(x)), 0.5, self.training)
        return x

    def _convs_masked(self, x, in_lens):
        """All utterances at once: the reference crops every utterance and runs the conv stack on it alone
        (common.py:348-356); a length mask through the partial convs and the instance norms is the same math
        without the Python loop over the batch.  On CUDA everything after each k-tap convolution (partial-conv
        renormalisation, masked instance norm, ReLU, dropout, mask) is ONE fused kernel (ops.encoder_conv_block)."""
        mask = get_mask_from_lengths(in_lens, x.shape[2])[:, None].to(x.dtype)
        if x.is_cuda and all(isinstance(b[1], nn.InstanceNorm1d) and isinstance(b[0].conv, PartialConv1d)
                             and not hasattr(b[0].conv, "weight_v") for b in self.convolutions):
            x = x * mask
            for block in self.convolutions:
                x = ops.encoder_conv_block(block[0].conv, block[1], x, in_lens, 0.5, self.training)
            return x
        n = in_lens.to(x.dtype).clamp(min=1)[:, None, None]
        for block in self.convolutions:
            conv, norm = block[0], block[1]
            y = conv(x, mask)
            mean = (y * mask).sum(2, keepdim=True) / n
            var = (((y - mean) * mask) ** 2).sum(2, keepdim=True) / n
            y = (y - mean) * torch.rsqrt(var + norm.eps)
            if norm.weight is not None:
                y = y * norm.weight[None, :, None] + norm.bias[None, :, None]
            x = F.dropout(F.relu(y), 0.5, self.training) * mask
        return x

    def forward(self, x, in_lens):
        with torch.autocast(device_type=x.device.type, enabled=False):
            x = x.float()
            if x.shape[0] > 1:
                x = self._convs_masked(x, in_lens).transpose(1, 2)
            else:
                x = self._convs(x).transpose(1, 2)
            out = run_bilstm(self.lstm, x, in_lens)
        return out

    def infer(self, x):
        """reference common.py:375-384: the whole (padded) batch goes through the LSTM unpacked, i.e. every utterance at
        the full padded length -- the persistent recurrence kernel with lens = T for all."""
        with torch.autocast(device_type=x.device.type, enabled=False):
            x = self._convs(x.float()).transpose(1, 2)
            if lstm_ops.supported(self.lstm, x):
                full = torch.full((x.shape[0],), x.shape[1], dtype=torch.int32, device=x.device)
                return lstm_ops.bilstm(self.lstm, x, full)
            return self.lstm(x)[0]


# --------------------------------------------------------------------------------------------------------
# hot-path layers
# --------------------------------------------------------------------------------------------------------
def _random_rotation(c):
    w = torch.linalg.qr(torch.randn(c, c))[0]
    if torch.det(w) < 0:
        w[:, 0] = -w[:, 0]
    return w


class Invertible1x1ConvLUS(nn.Module):
    """LU-parameterised invertible 1x1 conv (reference common.py:387-428).  Parameters `lower`, `upper`,
    `upper_diag`, buffers `p`, `lower_diag` as in the reference.  Always fp32."""

    def __init__(self, c, cache_inverse=False):
        super().__init__()
        p, lower, upper = torch.linalg.lu(_random_rotation(c))
        self.register_buffer("p", p)
        self.register_buffer("lower_diag", torch.ones(c))
        self.lower = nn.Parameter(torch.tril(lower, -1))
        self.upper_diag = nn.Parameter(torch.diag(upper).clone())
        self.upper = nn.Parameter(torch.triu(upper, 1))
        self.cache_inverse = cache_inverse

    def weight(self):
        with torch.autocast(device_type=self.lower.device.type, enabled=False):
            up = torch.triu(self.upper, 1) + torch.diag(self.upper_diag)
            lo = torch.tril(self.lower, -1) + torch.diag(self.lower_diag)
            return self.p @ (lo @ up)

    def log_det(self):
        return torch.sum(torch.log(torch.abs(self.upper_diag)))

    def inverse_weight(self):
        if hasattr(self, "W_inverse"):
            return self.W_inverse
        w_inv = torch.linalg.inv(self.weight().float())
        if self.cache_inverse:
            self.W_inverse = w_inv
        return w_inv

    def forward(self, z, inverse=False):
        if inverse:
            return ops.pointwise_conv(z.float(), self.inverse_weight())
        return ops.pointwise_conv(z.float(), self.weight()), self.log_det()


class Invertible1x1Conv(nn.Module):
    """Plain invertible 1x1 conv (reference common.py:431-472), used by the BGAP attribute flows."""

    def __init__(self, c, cache_inverse=False):
        super().__init__()
        self.conv = nn.Conv1d(c, c, kernel_size=1, stride=1, padding=0, bias=False)
        self.conv.weight.data = _random_rotation(c).view(c, c, 1)
        self.cache_inverse = cache_inverse

    def inverse_weight(self):
        if hasattr(self, "W_inverse"):
            return self.W_inverse
        w_inv = torch.linalg.inv(self.conv.weight.squeeze(-1).float())
        if self.cache_inverse:
            self.W_inverse = w_inv
        return w_inv

    def forward(self, z, inverse=False):
        w = self.conv.weight.squeeze(-1)
        if inverse:
            return ops.pointwise_conv(z.float(), self.inverse_weight())
        return ops.pointwise_conv(z.float(), w.float()), torch.logdet(w).clone()


class SimpleConvNet(nn.Module):
    """ConvNorm+ReLU stack with a 1x1 head (reference common.py:475-515)."""

    def __init__(self, n_mel_channels, n_context_dim, final_out_channels, n_layers=2, kernel_size=5,
                 with_dilation=True, max_channels=1024, zero_init=True, use_partial_padding=True):
        super().__init__()
        self.layers = nn.ModuleList()
        self.n_layers = n_layers
        self.kernel_size = kernel_size
        self.with_dilation = with_dilation
        self.use_partial_padding = use_partial_padding
        in_channels = n_mel_channels + n_context_dim
        out_channels = -1
        for i in range(n_layers):
            dilation = 2 ** i if with_dilation else 1
            padding = int((kernel_size * dilation - dilation) / 2)
            out_channels = min(max_channels, in_channels * 2)
            self.layers.append(ConvNorm(in_channels, out_channels, kernel_size=kernel_size, stride=1,
                                        padding=padding, dilation=dilation, bias=True, w_init_gain="relu",
                                        use_partial_padding=use_partial_padding))
            in_channels = out_channels
        self.last_layer = nn.Conv1d(out_channels, final_out_channels, kernel_size=1)
        if zero_init:
            self.last_layer.weight.data *= 0
            self.last_layer.bias.data *= 0

    def forward(self, z_w_context, seq_lens=None):
        return ops.simple_conv_net(self, z_w_context, seq_lens)


class WN(nn.Module):
    """Dilated-conv parameter network of the affine coupling (reference common.py:518-578): weight-normed
    1x1 `start`, n_layers x [partial conv k5 dilation 2^i -> softplus -> 1x1 res_skip -> softplus -> sum],
    zero-initialised 1x1 `end`.  The math runs in the fused CUDA flow-step kernels (ops.flow_step); this
    module only owns the parameters."""

    def __init__(self, n_in_channels, n_context_dim, n_layers, n_channels, kernel_size=5,
                 affine_activation="softplus", use_partial_padding=True):
        super().__init__()
        assert kernel_size % 2 == 1 and n_channels % 2 == 0
        self.n_layers = n_layers
        self.n_channels = n_channels
        self.n_in_channels = n_in_channels
        self.n_context_dim = n_context_dim
        self.kernel_size = kernel_size
        self.affine_activation = affine_activation
        self.use_partial_padding = use_partial_padding
        self.in_layers = nn.ModuleList()
        self.res_skip_layers = nn.ModuleList()
        self.start = nn.utils.weight_norm(nn.Conv1d(n_in_channels + n_context_dim, n_channels, 1), name="weight")
        self.softplus = nn.Softplus()
        end = nn.Conv1d(n_channels, 2 * n_in_channels, 1)
        end.weight.data.zero_()
        end.bias.data.zero_()
        self.end = end
        for i in range(n_layers):
            dilation = 2 ** i
            padding = int((kernel_size * dilation - dilation) / 2)
            self.in_layers.append(ConvNorm(n_channels, n_channels, kernel_size=kernel_size, dilation=dilation,
                                           padding=padding, use_partial_padding=use_partial_padding,
                                           use_weight_norm=True))
            self.res_skip_layers.append(nn.utils.weight_norm(nn.Conv1d(n_channels, n_channels, 1)))

    def forward(self, forward_input, seq_lens=None):
        z, context = forward_input
        return ops.wn_forward(self, z, context, seq_lens)


_SCALING_FNS = ("translate", "exp", "tanh", "sigmoid")


class AffineTransformationLayer(nn.Module):
    """Affine coupling (reference common.py:746-832)."""

    def __init__(self, n_mel_channels, n_context_dim, n_layers, affine_model="simple_conv", with_dilation=True,
                 kernel_size=5, scaling_fn="exp", affine_activation="softplus", n_channels=1024,
                 use_partial_padding=False):
        super().__init__()
        if affine_model not in ("wavenet", "simple_conv"):
            raise Exception("{} affine model not supported".format(affine_model))
        fns = scaling_fn if isinstance(scaling_fn, list) else [scaling_fn]
        if not all(f in _SCALING_FNS for f in fns):
            raise Exception("{} scaling fn not supported".format(scaling_fn))
        self.affine_model = affine_model
        self.scaling_fn = scaling_fn
        if affine_model == "wavenet":
            self.affine_param_predictor = WN(int(n_mel_channels / 2), n_context_dim, n_layers=n_layers,
                                             n_channels=n_channels, affine_activation=affine_activation,
                                             use_partial_padding=use_partial_padding)
        else:
            self.affine_param_predictor = SimpleConvNet(int(n_mel_channels / 2), n_context_dim, n_mel_channels,
                                                        n_layers, with_dilation=with_dilation,
                                                        kernel_size=kernel_size,
                                                        use_partial_padding=use_partial_padding)
        self.n_mel_channels = n_mel_channels

    def forward(self, z, context, inverse=False, seq_lens=None):
        return ops.affine_coupling(self, z, context, inverse, seq_lens)


class SplineTransformationLayer(nn.Module):
    """Spline coupling (reference common.py:663-743); the rational-quadratic (use_quadratic=True) variant used
    by the shipped BGAP configs runs in the CUDA library."""

    def __init__(self, n_mel_channels, n_context_dim, n_layers, with_dilation=True, kernel_size=5,
                 scaling_fn="exp", affine_activation="softplus", n_channels=1024, n_bins=8, left=-4, right=4,
                 bottom=-4, top=4, use_quadratic=False):
        super().__init__()
        self.n_mel_channels = n_mel_channels
        self.half_mel_channels = int(n_mel_channels / 2)
        self.left, self.right, self.bottom, self.top = left, right, bottom, top
        self.n_bins = n_bins
        self.use_quadratic = use_quadratic
        if not use_quadratic:
            raise NotImplementedError("piecewise-linear spline coupling (use_quadratic=False) is outside the "
                                      "B200 hot-path scope; no shipped config uses it")
        self.n_bins = 2 * self.n_bins + 1
        self.param_predictor = SimpleConvNet(self.half_mel_channels, n_context_dim,
                                             self.half_mel_channels * self.n_bins, n_layers,
                                             with_dilation=with_dilation, kernel_size=kernel_size, zero_init=False)

    def forward(self, z, context, inverse=False, seq_lens=None):
        return ops.spline_coupling(self, z, context, inverse, seq_lens)


class ConvAttention(nn.Module):
    """Text<->mel soft alignment (reference common.py:835-924): key/query conv projections, isotropic
    Gaussian log-likelihood, log-softmax + prior, masked softmax -- one fused CUDA path (ops.conv_attention)."""

    def __init__(self, n_mel_channels=80, n_text_channels=512, n_att_channels=80, temperature=1.0):
        super().__init__()
        self.temperature = temperature
        self.key_proj = nn.Sequential(
            ConvNorm(n_text_channels, n_text_channels * 2, kernel_size=3, bias=True, w_init_gain="relu"),
            nn.ReLU(),
            ConvNorm(n_text_channels * 2, n_att_channels, kernel_size=1, bias=True))
        self.query_proj = nn.Sequential(
            ConvNorm(n_mel_channels, n_mel_channels * 2, kernel_size=3, bias=True, w_init_gain="relu"),
            nn.ReLU(),
            ConvNorm(n_mel_channels * 2, n_mel_channels, kernel_size=1, bias=True),
            nn.ReLU(),
            ConvNorm(n_mel_channels, n_att_channels, kernel_size=1, bias=True))

    def forward(self, queries, keys, query_lens, mask=None, key_lens=None, attn_prior=None):
        return ops.conv_attention(self, queries, keys, mask, key_lens, attn_prior)
