"""Attribute predictors (reference attribute_prediction_model.py): DAP (deterministic, PyTorch, outside the
hot path) and BGAP (bi-partite flow: affine x2 + RQ-spline x4 couplings, CUDA hot path for config 4).
AGAP (LSTM autoregressive flow) is outside the scope of this build."""
import torch
from torch import nn

from . import ops
from .common import (AffineTransformationLayer, ConvLSTMLinear, ConvNorm, Invertible1x1Conv,
                     SplineTransformationLayer)


def get_attribute_prediction_model(config):
    name, hparams = config["name"], config["hparams"]
    if name == "dap":
        return DAP(**hparams)
    if name == "bgap":
        return BGAP(**hparams)
    if name == "agap":
        raise NotImplementedError("AGAP (autoregressive LSTM flow) is outside the radtts_b200 hot-path scope")
    raise Exception("{} model is not supported".format(name))


class AttributeProcessing:
    def __init__(self, take_log_of_input=False):
        self.take_log_of_input = take_log_of_input

    def normalize(self, x):
        return torch.log(x + 1) if self.take_log_of_input else x

    def denormalize(self, x):
        return torch.exp(x) - 1 if self.take_log_of_input else x


class BottleneckLayerLayer(nn.Module):
    """Channel-reducing conv in front of every attribute predictor (reference :61-85)."""

    def __init__(self, in_dim, reduction_factor, norm="weightnorm", non_linearity="relu", kernel_size=3,
                 use_partial_padding=False):
        super().__init__()
        self.reduction_factor = reduction_factor
        self.out_dim = int(in_dim / reduction_factor)
        if reduction_factor > 1:
            fn = ConvNorm(in_dim, self.out_dim, kernel_size=kernel_size, use_weight_norm=(norm == "weightnorm"))
            if norm == "instancenorm":
                fn = nn.Sequential(fn, nn.InstanceNorm1d(self.out_dim, affine=True))
            self.projection_fn = fn
            self.non_linearity = nn.LeakyReLU() if non_linearity == "leakyrelu" else nn.ReLU()

    def forward(self, x):
        if self.reduction_factor > 1:
            x = self.non_linearity(self.projection_fn(x))
        return x


class DAP(nn.Module):
    def __init__(self, n_speaker_dim, bottleneck_hparams, take_log_of_input, arch_hparams, use_transformer=False):
        super().__init__()
        if use_transformer:
            raise NotImplementedError("FFTransformer predictor is unused by the shipped configs")
        self.attribute_processing = AttributeProcessing(take_log_of_input)
        self.bottleneck_layer = BottleneckLayerLayer(**bottleneck_hparams)
        arch_hparams = dict(arch_hparams)
        arch_hparams["in_dim"] = self.bottleneck_layer.out_dim + n_speaker_dim
        self.feat_pred_fn = ConvLSTMLinear(**arch_hparams)

    def forward(self, txt_enc, spk_emb, x, lens):
        if x is not None:
            x = self.attribute_processing.normalize(x)
        txt_enc = self.bottleneck_layer(txt_enc)
        spk = spk_emb[..., None].expand(-1, -1, txt_enc.shape[2])
        x_hat = self.feat_pred_fn(torch.cat((txt_enc, spk), 1), lens)
        return {"x_hat": x_hat, "x": x}

    def infer(self, z, txt_enc, spk_emb, lens=None):
        x_hat = self.forward(txt_enc, spk_emb, x=None, lens=lens)["x_hat"]
        return self.attribute_processing.denormalize(x_hat)


class BGAP(nn.Module):
    """Bi-partite generative attribute predictor (reference :120-224).  Note the order inside a step is
    transform THEN 1x1 conv -- the opposite of the decoder's FlowStep."""

    def __init__(self, n_in_dim, n_speaker_dim, bottleneck_hparams, n_flows, n_group_size, n_layers,
                 with_dilation, kernel_size, scaling_fn, take_log_of_input=False, n_channels=1024,
                 use_quadratic=False, n_bins=8, n_spline_steps=2):
        super().__init__()
        self.n_flows = n_flows
        self.n_group_size = n_group_size
        self.transforms = nn.ModuleList()
        self.convinv = nn.ModuleList()
        self.n_speaker_dim = n_speaker_dim
        self.scaling_fn = scaling_fn
        self.attribute_processing = AttributeProcessing(take_log_of_input)
        self.n_spline_steps = n_spline_steps
        self.bottleneck_layer = BottleneckLayerLayer(**bottleneck_hparams)
        context_dim = self.bottleneck_layer.out_dim * n_group_size + n_speaker_dim
        for k in range(n_flows):
            self.convinv.append(Invertible1x1Conv(n_in_dim * n_group_size))
            if k >= n_flows - n_spline_steps:
                self.transforms.append(SplineTransformationLayer(
                    n_in_dim * n_group_size, context_dim, n_layers, with_dilation=with_dilation,
                    kernel_size=kernel_size, scaling_fn=scaling_fn, n_channels=n_channels, top=3, bottom=-3,
                    left=-3, right=3, use_quadratic=use_quadratic, n_bins=n_bins))
            else:
                self.transforms.append(AffineTransformationLayer(
                    n_in_dim * n_group_size, context_dim, n_layers, with_dilation=with_dilation,
                    kernel_size=kernel_size, scaling_fn=scaling_fn, affine_model="simple_conv",
                    n_channels=n_channels))

    def unfold(self, x4):
        return ops.squeeze_time(x4.squeeze(-1), self.n_group_size) if self.n_group_size > 1 else x4.squeeze(-1)

    def fold(self, data):
        return ops.unsqueeze_time(data, self.n_group_size) if self.n_group_size > 1 else data

    def preprocess_context(self, txt_emb, speaker_vecs, std_scale=None):
        if self.n_group_size > 1:
            txt_emb = ops.squeeze_time(txt_emb, self.n_group_size)
        spk = speaker_vecs[..., None].expand(-1, -1, txt_emb.shape[2])
        return torch.cat((txt_emb, spk), 1)

    def forward(self, txt_enc, spk_emb, x, lens):
        assert txt_enc.size(2) >= x.size(1)
        if x.dim() == 2:
            x = x[:, None]
        txt_enc = self.bottleneck_layer(txt_enc)
        lens_grouped = (lens // self.n_group_size).long()
        context = self.preprocess_context(txt_enc, spk_emb)
        x = ops.squeeze_time(x, self.n_group_size)
        log_s_list, log_det_W_list = [], []
        for k in range(self.n_flows):
            x, log_s = self.transforms[k](x, context, seq_lens=lens_grouped)
            x, log_det_W = self.convinv[k](x)
            log_det_W_list.append(log_det_W)
            log_s_list.append(log_s)
        return {"z": x, "log_det_W_list": log_det_W_list, "log_s_list": log_s_list}

    def infer(self, z, txt_enc, spk_emb, seq_lens):
        txt_enc = self.bottleneck_layer(txt_enc)
        context = self.preprocess_context(txt_enc, spk_emb)
        lens_grouped = (seq_lens // self.n_group_size).long()
        z = ops.squeeze_time(z, self.n_group_size)
        for k in reversed(range(self.n_flows)):
            z = self.convinv[k](z, inverse=True)
            z = self.transforms[k].forward(z, context, inverse=True, seq_lens=lens_grouped)
        return self.fold(z)
