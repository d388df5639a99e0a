"""`model_config` dictionaries equivalent to the reference's shipped configs (reference configs/*.json),
expressed as a base + per-variant overrides.  `RADTTS(**model_config("radtts"))` builds the same
architecture, with the same state-dict names, as the reference does from config_ljs_radtts.json.
A user's own JSON (json.load(...)["model_config"]) works just as well -- the constructor takes the same keys.
"""
import copy


def _bottleneck(non_linearity="relu", **extra):
    d = {"in_dim": 512, "reduction_factor": 16, "norm": "weightnorm", "non_linearity": non_linearity}
    d.update(extra)
    return d


def _dap(kernel_size=3, p_dropout=0.25, take_log_of_input=False, use_transformer=None, **arch_extra):
    arch = {"out_dim": 1, "n_layers": 2, "n_channels": 256, "kernel_size": kernel_size, "p_dropout": p_dropout}
    arch.update(arch_extra)
    hp = {"n_speaker_dim": 16, "bottleneck_hparams": _bottleneck(), "take_log_of_input": take_log_of_input}
    if use_transformer is not None:      # only config_ljs_dap.json spells the key out
        hp["use_transformer"] = use_transformer
    hp["arch_hparams"] = arch
    return {"name": "dap", "hparams": hp}


def _bgap(n_group_size):
    return {"name": "bgap", "hparams": {
        "n_in_dim": 2, "take_log_of_input": False, "n_speaker_dim": 16, "n_flows": 6, "n_group_size": n_group_size,
        "n_layers": 4, "kernel_size": 5, "scaling_fn": "tanh", "with_dilation": True,
        "bottleneck_hparams": _bottleneck("leakyrelu", use_partial_padding=True, kernel_size=1),
        "n_bins": 16, "use_quadratic": True, "n_spline_steps": 4}}


_BASE = {
    "n_speakers": 1, "n_speaker_dim": 16, "n_text": 185, "n_text_dim": 512, "n_flows": 8,
    "n_conv_layers_per_step": 4, "n_mel_channels": 80, "n_hidden": 1024, "mel_encoder_n_hidden": 512,
    "dummy_speaker_embedding": False, "n_early_size": 2, "n_early_every": 2, "n_group_size": 2,
    "affine_model": "wavenet", "include_modules": "decatn", "scaling_fn": "tanh", "matrix_decomposition": "LUS",
    "learn_alignments": True, "use_speaker_emb_for_alignment": False, "attn_straight_through_estimator": True,
    "use_context_lstm": True, "context_lstm_norm": "spectral", "context_lstm_w_f0_and_energy": False,
    "text_encoder_lstm_norm": "spectral", "n_f0_dims": 0, "n_energy_avg_dims": 0,
    "use_first_order_features": False, "unvoiced_bias_activation": "relu", "decoder_use_partial_padding": True,
    "decoder_use_unvoiced_bias": False, "ap_pred_log_f0": False, "ap_use_unvoiced_bias": False,
    "ap_use_voiced_embeddings": False, "dur_model_config": _dap(take_log_of_input=True),
    "f0_model_config": None, "energy_model_config": None, "v_model_config": None,
}

_VPRED = _dap(p_dropout=0.5, lstm_type="", use_linear=1)
_ATTR = {"n_f0_dims": 1, "n_energy_avg_dims": 1, "context_lstm_w_f0_and_energy": True,
         "decoder_use_unvoiced_bias": True, "ap_pred_log_f0": True, "ap_use_voiced_embeddings": True,
         "v_model_config": _VPRED}

_VARIANTS = {
    "radtts": {},
    "decoder": dict(_ATTR, include_modules="decatnvpred", ap_use_unvoiced_bias=True, dur_model_config=None),
    "bgap": dict(_ATTR, include_modules="decatndpmvpredapm", use_first_order_features=True,
                 f0_model_config=_bgap(2), energy_model_config=_bgap(4)),
    "dap": dict(_ATTR, include_modules="decatndpmvpredapm",
                f0_model_config=_dap(kernel_size=11, p_dropout=0.5, use_transformer=False),
                energy_model_config=_dap(use_transformer=False)),
}

LOSS_WEIGHTS = {"blank_logprob": -1, "ctc_loss_weight": 0.1, "binarization_loss_weight": 1.0, "dur_loss_weight": 1.0,
                "f0_loss_weight": 1.0, "energy_loss_weight": 1.0, "vpred_loss_weight": 1.0}


def model_config(name="radtts"):
    """name in {'radtts', 'decoder', 'bgap', 'dap'} <-> reference configs/config_ljs_<name>.json."""
    cfg = copy.deepcopy(_BASE)
    cfg.update(copy.deepcopy(_VARIANTS[name]))
    return cfg
