"""RADTTS model with the reference's constructor / forward / infer API (reference radtts.py), built on the
B200 CUDA hot path: ConvAttention + MAS (kernels 3 and 1) and the decoder flow stack (kernel 2).

`RADTTS(**model_config)` accepts every key of the reference configs/*.json `model_config` section; the
state dict uses the reference names, so reference checkpoints load with load_state_dict().
"""
import torch
from torch import nn

from . import lstm_ops, alignment, ops
from .attribute_prediction_model import get_attribute_prediction_model
from .common import (AffineTransformationLayer, ConvAttention, Encoder, ExponentialClass, Invertible1x1Conv,
                     Invertible1x1ConvLUS, LengthRegulator, LinearNorm, _apply_lstm_norm, get_mask_from_lengths,
                     run_bilstm)


def _noise(shape, device):
    """Standard-normal draw for RADTTS.infer (reference radtts.py:559,607,622,652 use torch.cuda.FloatTensor.normal_);
    a module-level function so that parity tests can substitute the reference's recorded noise."""
    return torch.randn(*shape, device=device)


class FlowStep(nn.Module):
    """One decoder flow: invertible 1x1 conv then affine coupling (reference radtts.py:31-59)."""

    def __init__(self, n_mel_channels, n_context_dim, n_layers, affine_model="simple_conv", scaling_fn="exp",
                 matrix_decomposition="", affine_activation="softplus", use_partial_padding=False,
                 cache_inverse=False):
        super().__init__()
        conv_cls = Invertible1x1ConvLUS if matrix_decomposition == "LUS" else Invertible1x1Conv
        self.invtbl_conv = conv_cls(n_mel_channels, cache_inverse=cache_inverse)
        self.affine_tfn = AffineTransformationLayer(n_mel_channels, n_context_dim, n_layers,
                                                    affine_model=affine_model, scaling_fn=scaling_fn,
                                                    affine_activation=affine_activation,
                                                    use_partial_padding=use_partial_padding)

    def enable_inverse_cache(self):
        self.invtbl_conv.cache_inverse = True

    def forward(self, z, context, inverse=False, seq_lens=None):
        return ops.flow_step(self, z, context, inverse, seq_lens)


class RADTTS(nn.Module):
    def __init__(self, n_speakers, n_speaker_dim, n_text, n_text_dim, n_flows, n_conv_layers_per_step,
                 n_mel_channels, n_hidden, mel_encoder_n_hidden, dummy_speaker_embedding, n_early_size,
                 n_early_every, n_group_size, affine_model, dur_model_config, f0_model_config,
                 energy_model_config, v_model_config=None, include_modules="dec", scaling_fn="exp",
                 matrix_decomposition="", learn_alignments=False, affine_activation="softplus",
                 attn_use_CTC=True, use_speaker_emb_for_alignment=False, use_context_lstm=False,
                 context_lstm_norm=None, text_encoder_lstm_norm=None, n_f0_dims=0, n_energy_avg_dims=0,
                 context_lstm_w_f0_and_energy=True, use_first_order_features=False, unvoiced_bias_activation="",
                 ap_pred_log_f0=False, **kwargs):
        super().__init__()
        assert n_early_size % 2 == 0
        assert n_speaker_dim % 2 == 0
        self.do_mel_descaling = kwargs.get("do_mel_descaling", True)
        self.n_mel_channels = n_mel_channels
        self.n_f0_dims = n_f0_dims
        self.n_energy_avg_dims = n_energy_avg_dims
        self.decoder_use_partial_padding = kwargs.get("decoder_use_partial_padding", True)
        self.n_speaker_dim = n_speaker_dim
        self.speaker_embedding = nn.Embedding(n_speakers, n_speaker_dim)
        self.embedding = nn.Embedding(n_text, n_text_dim)
        self.flows = nn.ModuleList()
        self.encoder = Encoder(encoder_embedding_dim=n_text_dim, norm_fn=nn.InstanceNorm1d,
                               lstm_norm_fn=text_encoder_lstm_norm)
        self.dummy_speaker_embedding = dummy_speaker_embedding
        self.learn_alignments = learn_alignments
        self.affine_activation = affine_activation
        self.include_modules = include_modules
        self.attn_use_CTC = bool(attn_use_CTC)
        self.use_speaker_emb_for_alignment = use_speaker_emb_for_alignment
        self.use_context_lstm = bool(use_context_lstm)
        self.context_lstm_norm = context_lstm_norm
        self.context_lstm_w_f0_and_energy = context_lstm_w_f0_and_energy
        self.length_regulator = LengthRegulator()
        self.use_first_order_features = bool(use_first_order_features)
        self.decoder_use_unvoiced_bias = kwargs.get("decoder_use_unvoiced_bias", True)
        self.ap_pred_log_f0 = ap_pred_log_f0
        self.ap_use_unvoiced_bias = kwargs.get("ap_use_unvoiced_bias", True)
        self.attn_straight_through_estimator = kwargs.get("attn_straight_through_estimator", False)

        if "atn" in include_modules or "dec" in include_modules:
            if learn_alignments:
                key_dim = n_text_dim + (n_speaker_dim if use_speaker_emb_for_alignment else 0)
                self.attention = ConvAttention(n_mel_channels, key_dim)
            self.n_flows = n_flows
            self.n_group_size = n_group_size
            n_cond = n_speaker_dim + (n_text_dim + n_f0_dims + n_energy_avg_dims) * n_group_size
            if self.use_context_lstm:
                n_in = n_speaker_dim + n_text_dim * n_group_size
                n_hid = int(n_in / 2)
                if context_lstm_w_f0_and_energy:
                    n_in = (n_f0_dims + n_energy_avg_dims + n_text_dim) * n_group_size + n_speaker_dim
                    n_cond = n_speaker_dim + n_text_dim * n_group_size
                self.context_lstm = nn.LSTM(input_size=n_in, hidden_size=n_hid, num_layers=1, batch_first=True,
                                            bidirectional=True)
                self.context_lstm = _apply_lstm_norm(self.context_lstm, context_lstm_norm)
            if n_group_size > 1:
                self.unfold_params = {"kernel_size": (n_group_size, 1), "stride": n_group_size, "padding": 0,
                                      "dilation": 1}
            self.exit_steps = []
            self.n_early_size = n_early_size
            c = n_mel_channels * n_group_size
            for i in range(n_flows):
                if i > 0 and i % n_early_every == 0:
                    c -= n_early_size
                    self.exit_steps.append(i)
                self.flows.append(FlowStep(c, n_cond, n_conv_layers_per_step, affine_model, scaling_fn,
                                           matrix_decomposition, affine_activation=affine_activation,
                                           use_partial_padding=self.decoder_use_partial_padding))

        if "dpm" in include_modules:
            dur_model_config["hparams"]["n_speaker_dim"] = n_speaker_dim
            self.dur_pred_layer = get_attribute_prediction_model(dur_model_config)

        self.use_unvoiced_bias = False
        self.use_vpred_module = False
        self.ap_use_voiced_embeddings = kwargs.get("ap_use_voiced_embeddings", True)
        if self.decoder_use_unvoiced_bias or self.ap_use_unvoiced_bias:
            assert unvoiced_bias_activation in {"relu", "exp"}
            self.use_unvoiced_bias = True
            nonlin = nn.ReLU() if unvoiced_bias_activation == "relu" else ExponentialClass()
            self.unvoiced_bias_module = nn.Sequential(LinearNorm(n_text_dim, 1), nonlin)
        if self.ap_use_voiced_embeddings or self.use_unvoiced_bias or "vpred" in include_modules:
            self.use_vpred_module = True
        if self.use_vpred_module:
            v_model_config["hparams"]["n_speaker_dim"] = n_speaker_dim
            self.v_pred_module = get_attribute_prediction_model(v_model_config)
            if self.ap_use_voiced_embeddings:
                self.v_embeddings = nn.Embedding(4, n_text_dim)

        if "apm" in include_modules:
            for cfg in (f0_model_config, energy_model_config):
                hp = cfg["hparams"]
                hp["n_speaker_dim"] = n_speaker_dim
                if self.use_first_order_features:
                    hp["n_in_dim"] = 2
                if hp.get("spline_flow_params") is not None:
                    hp["spline_flow_params"]["n_in_channels"] = hp["n_in_dim"]
            self.f0_pred_module = get_attribute_prediction_model(f0_model_config)
            self.energy_pred_module = get_attribute_prediction_model(energy_model_config)

    # ------------------------------------------------------------------------------------------------
    def is_attribute_unconditional(self):
        return self.n_f0_dims == 0 and self.n_energy_avg_dims == 0

    def encode_speaker(self, spk_ids):
        if self.dummy_speaker_embedding:
            spk_ids = spk_ids * 0
        return self.speaker_embedding(spk_ids)

    def encode_text(self, text, in_lens):
        text_embeddings = self.embedding(text).transpose(1, 2)
        if in_lens is None:
            text_enc = self.encoder.infer(text_embeddings).transpose(1, 2)
        else:
            text_enc = self.encoder(text_embeddings, in_lens).transpose(1, 2)
        return text_enc, text_embeddings

    def unfold(self, x4):
        """nn.Unfold((g,1), stride g) on a (B,C,T,1) tensor (reference radtts.py:165-169)."""
        return ops.squeeze_time(x4.squeeze(-1), self.n_group_size)

    def fold(self, mel):
        return ops.unsqueeze_time(mel, self.n_group_size)

    def preprocess_context(self, context, speaker_vecs, out_lens=None, f0=None, energy_avg=None):
        """reference radtts.py:262-302."""
        g = self.n_group_size
        if g > 1:
            context = ops.squeeze_time(context, g)
            if f0 is not None:
                f0 = ops.squeeze_time(f0[:, None], g)
            if energy_avg is not None:
                energy_avg = ops.squeeze_time(energy_avg[:, None], g)
        elif f0 is not None or energy_avg is not None:
            f0 = None if f0 is None else f0[:, None]
            energy_avg = None if energy_avg is None else energy_avg[:, None]
        spk = speaker_vecs[..., None].expand(-1, -1, context.shape[2])
        ctx = torch.cat((context, spk), 1)
        extras = [t for t in (f0, energy_avg) if t is not None]
        if self.use_context_lstm:
            if self.context_lstm_w_f0_and_energy and extras:
                ctx = torch.cat([ctx] + extras, 1)
            ctx = run_bilstm(self.context_lstm, ctx.transpose(1, 2), out_lens // g).transpose(1, 2)
        if not self.context_lstm_w_f0_and_energy and extras:
            ctx = torch.cat([ctx] + extras, 1)
        return ctx

    def enable_inverse_cache(self):
        for flow_step in self.flows:
            flow_step.enable_inverse_cache()

    def binarize_attention(self, attn, in_lens, out_lens):
        """MAS on the GPU (kernel 1); replaces the CPU/Numba loop of reference radtts.py:320-334."""
        return alignment.binarize_attention(attn, in_lens, out_lens)

    def get_first_order_features(self, feats, out_lens, dilation=1):
        pad = torch.zeros_like(feats[:, :dilation])
        right = torch.cat((feats, pad), 1)[:, dilation:] - feats
        left = feats - torch.cat((pad, feats), 1)[:, :-dilation]
        return (right + left) * 0.5

    def apply_voice_mask_to_text(self, text_enc, voiced_mask):
        vm = voiced_mask.unsqueeze(1)
        w = self.v_embeddings.weight
        scale = torch.sigmoid(w[0:1, :, None] * vm + w[1:2, :, None] * (1 - vm))
        bias = 0.1 * torch.tanh(w[2:3, :, None] * vm + w[3:4, :, None] * (1 - vm))
        return text_enc * scale + bias

    # ------------------------------------------------------------------------------------------------
    def forward(self, mel, speaker_ids, text, in_lens, out_lens, binarize_attention=False, attn_prior=None,
                f0=None, energy_avg=None, voiced_mask=None, p_voiced=None):
        prep = None
        if "dec" in self.include_modules and mel.is_cuda and torch.is_grad_enabled():
            # weight norm / LU composition / re-layout of all 8 flows start now, on a side stream, and finish underneath
            # the text encoder, the attention and the context LSTM
            prep = ops.begin_decoder_prep(self)
        if mel.is_cuda and torch.is_grad_enabled():
            # spectral-norm power iterations of the two BiLSTMs: off the critical path, on a side stream, now
            enc_lstm = getattr(self.encoder, "lstm", None)
            lstm_ops.prefetch_weights([enc_lstm, getattr(self, "context_lstm", None)], fp32=(enc_lstm,))
        speaker_vecs = self.encode_speaker(speaker_ids)
        text_enc, text_embeddings = self.encode_text(text, in_lens)
        log_s_list, log_det_W_list, z_mel = [], [], []
        attn = attn_soft = attn_hard = attn_logprob = None
        context = None
        if "atn" in self.include_modules or "dec" in self.include_modules:
            attn_mask = ~get_mask_from_lengths(in_lens, text.shape[1])[..., None]
            keys = text_embeddings
            if self.use_speaker_emb_for_alignment:
                spk = speaker_vecs[:, :, None].expand(-1, -1, text_embeddings.shape[2])
                keys = torch.cat((keys, spk.detach()), 1)
            attn_soft, attn_logprob = self.attention(mel, keys, out_lens, attn_mask, key_lens=in_lens,
                                                     attn_prior=attn_prior)
            if binarize_attention and type(self).binarize_attention is RADTTS.binarize_attention \
                    and attn_soft.is_cuda and attn_soft.shape[2] * 4 <= 48 * 1024:
                # the MAS kernel also returns each frame's token: the context "bmm with a one-hot matrix" of the
                # reference (radtts.py:399) becomes a gather (forward) / per-token segment sum (backward)
                with torch.no_grad():
                    attn, f2t, _ = alignment.mas_forward(attn_soft, in_lens, out_lens, is_prob=True, return_indices=True)
                attn_hard = attn
                if self.attn_straight_through_estimator:
                    attn_hard = attn_soft + (attn_hard - attn_soft).detach()
                context = ops.hard_attention_context(text_enc, f2t)
            else:
                if binarize_attention:
                    attn = self.binarize_attention(attn_soft, in_lens, out_lens)
                    attn_hard = attn
                    if self.attn_straight_through_estimator:
                        attn_hard = attn_soft + (attn_hard - attn_soft).detach()
                else:
                    attn = attn_soft
                context = torch.bmm(text_enc, attn.squeeze(1).transpose(1, 2))

        f0_bias = 0
        if self.use_unvoiced_bias:
            f0_bias = -self.unvoiced_bias_module(context.permute(0, 2, 1))[..., 0]
            f0_bias = f0_bias * (~voiced_mask.bool()).float()

        if "dec" in self.include_modules:
            if f0 is None:
                f0_aug = None
            elif self.decoder_use_unvoiced_bias:
                f0_aug = f0 * voiced_mask + f0_bias
            else:
                f0_aug = f0 * voiced_mask
            context_w_spkvec = self.preprocess_context(context, speaker_vecs, out_lens, f0_aug, energy_avg)
            z_mel, log_det_W_list, log_s_list = ops.decoder_forward(self, mel, context_w_spkvec, out_lens, prep=prep)

        duration_model_outputs = None
        if "dpm" in self.include_modules:
            if attn_hard is None:
                attn_hard = self.binarize_attention(attn_soft, in_lens, out_lens)
            durations = attn_hard.sum(2)[:, 0, :]
            duration_model_outputs = self.dur_pred_layer(text_enc.detach(), speaker_vecs.detach(),
                                                         durations.float().detach(), in_lens)

        f0_model_outputs = energy_model_outputs = vpred_model_outputs = None
        if "apm" in self.include_modules:
            if attn_hard is None:
                attn_hard = self.binarize_attention(attn_soft, in_lens, out_lens)
            if binarize_attention:
                text_enc_time_expanded = context.clone()
            else:
                text_enc_time_expanded = torch.bmm(text_enc, attn_hard.squeeze(1).transpose(1, 2))
            if self.use_vpred_module:
                vpred_model_outputs = self.v_pred_module(text_enc_time_expanded.detach(), speaker_vecs.detach(),
                                                         voiced_mask.detach(), out_lens)
                if self.ap_use_voiced_embeddings:
                    text_enc_time_expanded = self.apply_voice_mask_to_text(text_enc_time_expanded, voiced_mask)
            f0_target = f0.clone()
            if self.ap_use_unvoiced_bias:
                f0_target = (f0_target * voiced_mask + f0_bias).detach()
            else:
                f0_target = f0_target.detach()
            vb = voiced_mask.bool()
            f0_target[vb] = torch.log(f0_target[vb])
            f0_target = f0_target / 6
            energy_avg = energy_avg * 2 - 1
            if self.use_first_order_features:
                df0 = self.get_first_order_features(f0_target, out_lens)
                de = self.get_first_order_features(energy_avg, out_lens)
                f0_voiced = torch.cat((f0_target[:, None], df0[:, None]), 1) * 3
                energy_avg = torch.cat((energy_avg[:, None], de[:, None]), 1) * 3
            else:
                f0_voiced = f0_target * 2
                energy_avg = energy_avg * 1.4
            f0_model_outputs = self.f0_pred_module(text_enc_time_expanded, speaker_vecs.detach(), f0_voiced,
                                                   out_lens)
            energy_model_outputs = self.energy_pred_module(text_enc_time_expanded, speaker_vecs.detach(),
                                                           energy_avg, out_lens)

        lstm_ops.drop_prefetched([getattr(self.encoder, "lstm", None), getattr(self, "context_lstm", None)])
        return {"z_mel": z_mel, "log_det_W_list": log_det_W_list, "log_s_list": log_s_list,
                "duration_model_outputs": duration_model_outputs, "f0_model_outputs": f0_model_outputs,
                "energy_model_outputs": energy_model_outputs, "vpred_model_outputs": vpred_model_outputs,
                "attn_soft": attn_soft, "attn": attn, "text_embeddings": text_embeddings,
                "attn_logprob": attn_logprob}

    # ------------------------------------------------------------------------------------------------
    def infer(self, speaker_id, text, sigma, sigma_dur=0.8, sigma_f0=0.8, sigma_energy=0.8,
              token_dur_scaling=1.0, token_duration_max=100, speaker_id_text=None, speaker_id_attributes=None,
              dur=None, f0=None, energy_avg=None, voiced_mask=None, f0_mean=0.0, f0_std=0.0, energy_mean=0.0,
              energy_std=0.0, in_lens=None):
        """reference radtts.py:541-684.  `in_lens` (B,) is an extension for BATCHED synthesis of padded text (the
        reference's predictors cannot run with batch > 1): the text encoder and the duration
        predictor then respect each utterance's length and padded tokens get duration 0.  None = reference behaviour."""
        batch_size, n_tokens = text.shape[0], text.shape[1]
        dev = text.device
        spk_vec = self.encode_speaker(speaker_id)
        spk_vec_text = spk_vec if speaker_id_text is None else self.encode_speaker(speaker_id_text)
        spk_vec_attributes = spk_vec if speaker_id_attributes is None else self.encode_speaker(
            speaker_id_attributes)
        txt_enc, _ = self.encode_text(text, in_lens)

        if dur is None:
            z_dur = _noise((batch_size, 1, n_tokens), dev) * sigma_dur
            dur = self.dur_pred_layer.infer(z_dur, txt_enc, spk_vec_text, lens=in_lens)
            if dur.shape[-1] < txt_enc.shape[-1]:
                dur = nn.functional.pad(dur, (0, txt_enc.shape[-1] - dur.shape[2]), mode="replicate")
            dur = dur[:, 0].clamp(0, token_duration_max)
            if token_dur_scaling > 0:
                dur = dur * token_dur_scaling
            dur = (dur + 0.5).floor().int()
            if in_lens is not None:
                dur = dur * get_mask_from_lengths(in_lens.to(dev), n_tokens).to(dur.dtype)

        out_lens = dur.sum(1).long().to(dev)
        max_n_frames = int(out_lens.max())
        txt_enc_time_expanded = self.length_regulator(txt_enc.transpose(1, 2), dur).transpose(1, 2)

        if not self.is_attribute_unconditional():
            if voiced_mask is None and self.use_vpred_module:
                logits = self.v_pred_module.infer(None, txt_enc_time_expanded, spk_vec_attributes,
                                                  lens=out_lens if batch_size > 1 else None)
                voiced_mask = (torch.sigmoid(logits[:, 0]) > 0.5).float()
            ap_txt = txt_enc_time_expanded
            if self.ap_use_voiced_embeddings:
                ap_txt = self.apply_voice_mask_to_text(txt_enc_time_expanded, voiced_mask)
            f0_bias = 0
            if self.use_unvoiced_bias:
                f0_bias = -self.unvoiced_bias_module(txt_enc_time_expanded.permute(0, 2, 1))[..., 0]
                f0_bias = f0_bias * (~voiced_mask.bool()).float()
            if f0 is None:
                n_ch = 2 if self.use_first_order_features else 1
                z_f0 = _noise((batch_size, n_ch, max_n_frames), dev) * sigma_f0
                if batch_size > 1:      # batched extension: the padded region must look like the zero padding a
                    z_f0 = z_f0 * get_mask_from_lengths(out_lens, max_n_frames)[:, None].to(z_f0.dtype)   # B=1 run sees
                f0 = self.infer_f0(z_f0, ap_txt, spk_vec_attributes, voiced_mask, out_lens)[:, 0]
            if f0_mean > 0.0:
                vb = voiced_mask.bool()
                mu, sd = f0[vb].mean(), f0[vb].std()
                f0[vb] = (f0[vb] - mu) / sd
                f0[vb] = f0[vb] * (f0_std if f0_std > 0 else sd) + f0_mean
            if energy_avg is None:
                n_ch = 2 if self.use_first_order_features else 1
                z_e = _noise((batch_size, n_ch, max_n_frames), dev) * sigma_energy
                if batch_size > 1:
                    z_e = z_e * get_mask_from_lengths(out_lens, max_n_frames)[:, None].to(z_e.dtype)
                energy_avg = self.infer_energy(z_e, ap_txt, spk_vec, out_lens)[:, 0]
            n0 = int(out_lens[0])
            if energy_avg.shape[1] < n0:  # reference radtts.py:629-637 (both padded by the energy deficit)
                pad = n0 - energy_avg.shape[1]
                f0 = nn.functional.pad(f0[None], (0, pad), mode="replicate")[0]
                energy_avg = nn.functional.pad(energy_avg[None], (0, pad), mode="replicate")[0]
            if f0.shape[1] < n0:
                f0 = nn.functional.pad(f0[None], (0, n0 - f0.shape[1]), mode="replicate")[0]
            f0_in = f0 * voiced_mask + f0_bias if self.decoder_use_unvoiced_bias else f0 * voiced_mask
            context_w_spkvec = self.preprocess_context(txt_enc_time_expanded, spk_vec, out_lens, f0_in,
                                                       energy_avg)
        else:
            context_w_spkvec = self.preprocess_context(txt_enc_time_expanded, spk_vec, out_lens, None, None)

        residual = _noise((batch_size, 80 * self.n_group_size, max_n_frames // self.n_group_size), dev) * sigma
        mel = ops.decoder_inverse(self, residual, context_w_spkvec, out_lens)
        if self.do_mel_descaling:
            mel = mel * 2 - 5.5
        return {"mel": mel, "dur": dur, "f0": f0, "energy_avg": energy_avg, "voiced_mask": voiced_mask}

    def infer_f0(self, residual, txt_enc_time_expanded, spk_vec, voiced_mask=None, lens=None):
        f0 = self.f0_pred_module.infer(residual, txt_enc_time_expanded, spk_vec, lens)
        if voiced_mask is not None and voiced_mask.dim() == 2:
            voiced_mask = voiced_mask[:, None]
        if self.ap_pred_log_f0:
            f0 = (f0[:, 0:1, :] / 3 if self.use_first_order_features else f0 / 2) * 6
        else:
            f0 = f0 / 6 / 640
        voiced_mask = (f0 > 0.0) if voiced_mask is None else voiced_mask.bool()
        voiced_mask = voiced_mask[:, :, :f0.shape[-1]]
        if self.ap_pred_log_f0:
            f0[voiced_mask] = torch.exp(f0[voiced_mask])
        f0[~voiced_mask] = 0.0
        return f0

    def infer_energy(self, residual, txt_enc_time_expanded, spk_vec, lens):
        energy = self.energy_pred_module.infer(residual, txt_enc_time_expanded, spk_vec, lens)
        energy = energy / 3 if self.use_first_order_features else energy / 1.4
        return (energy + 1) / 2

    def remove_norms(self):
        """Strip weight/spectral norm re-parameterisations before inference (reference radtts.py:732-750)."""
        for name, module in self.named_modules():
            for fn, kw in ((nn.utils.remove_spectral_norm, {"name": "weight_hh_l0"}),
                           (nn.utils.remove_spectral_norm, {"name": "weight_hh_l0_reverse"}),
                           (nn.utils.remove_weight_norm, {})):
                try:
                    fn(module, **kw)
                    print("Removed norm from {}".format(name))
                except Exception:
                    pass
