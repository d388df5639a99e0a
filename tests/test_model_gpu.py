"""GPU parity of the full drop-in RADTTS.forward (+ RADTTSLoss) against goldens produced by running the unmodified
reference (tests/golden/radtts_forward.npz): ConvAttention (kernel 3), MAS (kernel 1), decoder flows (kernel 2)."""
import os

import numpy as np
import pytest
import torch

from radtts_b200 import configs, loss as rloss, ops, synth
from radtts_b200.radtts import RADTTS
from _gradcheck import check_param_grads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "radtts_forward.npz"))


@pytest.fixture(scope="module")
def model(cuda_lib):
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("radtts")).eval()
    synth.load_synth(m, seed=1234)
    return m.cuda()


def _batch():
    b = synth.synth_batch(3, 70, 24, seed=1234)
    return {k: v.cuda() for k, v in b.items()}


def _valid(x, lens):
    m = (torch.arange(x.shape[-1], device=x.device)[None, :] < lens.to(x.device)[:, None]).to(x.dtype)
    return x * m[:, None]


def test_forward_fp32_matches_reference(model, gold):
    ops.set_precision("fp32")
    # the out-of-scope PyTorch parts (text encoder convs, cuDNN LSTMs, the context bmm) default to TF32 on a GPU,
    # whose ~1e-3 noise is the size of the parity bar (trap 8b): run them in true fp32 here
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        b = _batch()
        with torch.no_grad():
            out = model(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True,
                        attn_prior=b["attn_prior"])
        # kernel 3
        assert torch.allclose(out["attn_soft"].cpu(), torch.from_numpy(gold["attn_soft"]), rtol=1e-3, atol=1e-6)
        lp, lp_ref = out["attn_logprob"].cpu(), torch.from_numpy(gold["attn_logprob"])
        assert torch.allclose(lp, lp_ref, rtol=1e-3, atol=1e-3)
        # kernel 1 inside the model: probabilities in, device log; report path agreement with the reference
        hard, hard_ref = out["attn"].cpu().numpy(), gold["attn"]
        frames_diff = int((hard != hard_ref).any(axis=3).sum())
        assert frames_diff <= 1, frames_diff
        # kernel 2
        lens = b["out_lens"] // 2
        if frames_diff == 0:
            assert torch.allclose(_valid(out["z_mel"], lens).cpu(), _valid(torch.from_numpy(gold["z_mel"]), lens.cpu()),
                                  rtol=1e-3, atol=2e-4)
            crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
            ld = crit(out, b["in_lens"], b["out_lens"])
            for k in ("loss_mel", "loss_prior_mel", "loss_ctc"):
                assert abs(float(ld[k][0]) - float(gold[k])) < 1e-3 * abs(float(gold[k])) + 1e-6, k
            lb = rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"])
            assert abs(float(lb) - float(gold["loss_binarization"])) < 1e-3 * abs(float(gold["loss_binarization"]))
        assert out["attn"].requires_grad is False
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)


def test_train_step_runs_and_all_hot_path_params_get_grads(model):
    """every flows.* / attention.* parameter receives a gradient; `attn` (hard) has none."""
    ops.set_precision("bf16")
    try:
        model.train()
        model.zero_grad(set_to_none=True)
        b = _batch()
        out = model(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True,
                    attn_prior=b["attn_prior"])
        crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
        ld = crit(out, b["in_lens"], b["out_lens"])
        loss = sum(v * w for v, w in ld.values()) + rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"])
        loss.backward()
        assert torch.isfinite(loss)
        missing = [n for n, p in model.named_parameters()
                   if (n.startswith("flows.") or n.startswith("attention.")) and (p.grad is None or not torch.isfinite(p.grad).all())]
        assert not missing, missing[:5]
        assert not out["attn"].requires_grad
    finally:
        model.eval()
        ops.set_precision(None)


def test_decoder_config_forward_matches_reference(golden_dir, cuda_lib):
    """config_ljs_decoder (cfg3): decoder conditioned on F0 / energy / voicing -- full forward vs reference golden."""
    g = np.load(os.path.join(golden_dir, "decoder_cfg_forward.npz"))
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("decoder")).eval()
    synth.load_synth(m, seed=1234)
    m = m.cuda()
    b = {k: v.cuda() for k, v in synth.synth_batch(2, 60, 20, seed=4321, with_attributes=True).items()}
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            out = m(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True,
                    attn_prior=b["attn_prior"], f0=b["f0"], energy_avg=b["energy_avg"], voiced_mask=b["voiced_mask"],
                    p_voiced=b["p_voiced"])
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
    assert torch.allclose(out["attn_soft"].cpu(), torch.from_numpy(g["attn_soft"]), rtol=1e-3, atol=1e-6)
    assert np.array_equal(out["attn"].cpu().numpy(), g["attn"])
    lens = b["out_lens"] // 2
    assert torch.allclose(_valid(out["z_mel"], lens).cpu(), _valid(torch.from_numpy(g["z_mel"]), lens.cpu()), rtol=1e-3,
                          atol=3e-4)
    assert np.allclose(np.array([float(x) for x in out["log_det_W_list"]]), g["log_det_W"], rtol=1e-4, atol=1e-5)
    for i in (0, 7):
        assert torch.allclose(_valid(out["log_s_list"][i], lens).cpu(),
                              _valid(torch.from_numpy(g["log_s_%d" % i]), lens.cpu()), rtol=1e-3, atol=1e-5)


def test_fused_attention_ctc_matches_torch(cuda_lib):
    """Fused CTC kernel vs the batched PyTorch formulation (itself checked against the reference's
    per-utterance loop on CPU): value and gradient."""
    torch.manual_seed(3)
    for (B, T1, T2) in [(4, 50, 17), (3, 120, 40), (2, 33, 33)]:
        lp = (torch.randn(B, 1, T1, T2, device="cuda") * 2 - 3)
        in_lens = torch.randint(max(1, T2 // 2), T2 + 1, (B,), device="cuda")
        out_lens = torch.randint(T2, T1 + 1, (B,), device="cuda")
        in_lens[0], out_lens[0] = T2, T1
        crit = rloss.AttentionCTCLoss()
        a = lp.clone().requires_grad_(True)
        la = crit(a, in_lens, out_lens)
        (ga,) = torch.autograd.grad(la * 1.7, a)
        r = lp.clone().requires_grad_(True)
        lr = crit.forward_torch(r, in_lens, out_lens)
        (gr,) = torch.autograd.grad(lr * 1.7, r)
        assert abs(float(la) - float(lr)) < 1e-4 * abs(float(lr)) + 1e-6, (float(la), float(lr))
        assert torch.allclose(ga, gr, rtol=1e-3, atol=1e-6), float((ga - gr).abs().max())


def test_bgap_config_trains_end_to_end(cuda_lib):
    """config_ljs_bgap (decoder + duration / voicing predictors + BGAP F0 and energy flows): forward, RADTTSLoss with
    the attribute losses, backward -- every trainable parameter of the attribute flows receives a finite gradient."""
    torch.manual_seed(0)
    cfg = configs.model_config("bgap")
    m = RADTTS(**cfg).train()
    synth.load_synth(m, seed=1234)
    m = m.cuda()
    b = {k: v.cuda() for k, v in synth.synth_batch(2, 64, 18, seed=77, with_attributes=True).items()}
    crit = rloss.RADTTSLoss(1.0, cfg["n_group_size"], cfg["dur_model_config"], cfg["f0_model_config"],
                            cfg["energy_model_config"], cfg["v_model_config"], configs.LOSS_WEIGHTS)
    ops.set_precision("fp32")
    try:
        out = m(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True,
                attn_prior=b["attn_prior"], f0=b["f0"], energy_avg=b["energy_avg"], voiced_mask=b["voiced_mask"],
                p_voiced=b["p_voiced"])
        losses = crit(out, b["in_lens"], b["out_lens"])
        total = sum(v * w for v, w in losses.values() if w > 0)
        total.backward()
    finally:
        ops.set_precision(None)
    assert torch.isfinite(total)
    assert {"loss_f0", "loss_energy"} <= set(losses) or any("f0" in k for k in losses), sorted(losses)
    for prefix in ("f0_pred_module.", "energy_pred_module."):
        got = [(n, p.grad) for n, p in m.named_parameters() if n.startswith(prefix) and p.requires_grad]
        assert got
        missing = [n for n, g in got if g is None]
        assert not missing, missing[:5]
        assert all(bool(torch.isfinite(g).all()) for _, g in got)
        assert sum(float(g.abs().sum()) for _, g in got) > 0


def test_soft_attention_training_direction_matches_reference(golden_dir, cuda_lib):
    """binarize_attention=False (the reference's regime before binarization_start_iter): context = bmm(text, attn_soft),
    so the flow loss reaches the attention and the text encoder.  Forward values, loss and EVERY parameter gradient
    against the reference's autograd (golden: oracle/make_golden.py::gen_radtts_forward_soft), fp32."""
    g = np.load(os.path.join(golden_dir, "radtts_forward_soft.npz"))
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("radtts")).eval()
    synth.load_synth(m, seed=1234)
    m = m.cuda()
    b = {k: v.cuda() for k, v in synth.synth_batch(2, 44, 15, seed=2468).items()}
    crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = m(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=False,
                attn_prior=b["attn_prior"])
        losses = crit(out, b["in_lens"], b["out_lens"])
        total = sum(v * w for v, w in losses.values() if w > 0)
        total.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
    lens = b["out_lens"] // 2
    assert torch.allclose(out["attn"].detach().cpu(), torch.from_numpy(g["attn"]), rtol=1e-3, atol=1e-6)
    assert torch.allclose(_valid(out["z_mel"].detach(), lens).cpu(), _valid(torch.from_numpy(g["z_mel"]), lens.cpu()),
                          rtol=1e-3, atol=3e-4)
    assert abs(float(losses["loss_mel"][0]) - float(g["loss_mel"])) < 1e-3 * abs(float(g["loss_mel"]))
    assert abs(float(losses["loss_ctc"][0]) - float(g["loss_ctc"])) < 1e-3 * abs(float(g["loss_ctc"]))
    assert abs(float(total) - float(g["total"])) < 1e-3 * abs(float(g["total"]))
    bad = check_param_grads(dict(m.named_parameters()), g)     # norm AND 64 strided values per parameter
    assert not bad, bad[:8]


def test_decoder_config_training_gradients_match_reference(golden_dir, cuda_lib):
    """config_ljs_decoder (F0 / energy / voicing conditioned decoder) in the binarized training regime: loss terms and
    every parameter gradient against the reference's autograd (golden: gen_decoder_cfg_train), fp32."""
    g = np.load(os.path.join(golden_dir, "decoder_cfg_train.npz"))
    torch.manual_seed(0)
    cfg = configs.model_config("decoder")
    m = RADTTS(**cfg).eval()
    synth.load_synth(m, seed=1234)
    m = m.cuda()
    b = {k: v.cuda() for k, v in synth.synth_batch(2, 48, 16, seed=1357, with_attributes=True).items()}
    crit = rloss.RADTTSLoss(1.0, cfg["n_group_size"], cfg["dur_model_config"], cfg["f0_model_config"],
                            cfg["energy_model_config"], cfg["v_model_config"], configs.LOSS_WEIGHTS)
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = m(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True,
                attn_prior=b["attn_prior"], f0=b["f0"], energy_avg=b["energy_avg"], voiced_mask=b["voiced_mask"],
                p_voiced=b["p_voiced"])
        losses = crit(out, b["in_lens"], b["out_lens"])
        total = sum(v * w for v, w in losses.values() if w > 0)
        total = total + rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"]) * \
            configs.LOSS_WEIGHTS["binarization_loss_weight"]
        total.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
    assert np.array_equal(out["attn"].detach().cpu().numpy(), g["attn"])
    for name, want in zip(g["loss_names"], g["loss_values"]):
        got = float(losses[str(name)][0])
        assert abs(got - float(want)) < 1e-3 * abs(float(want)) + 1e-6, (str(name), got, float(want))
    assert abs(float(total) - float(g["total"])) < 1e-3 * abs(float(g["total"]))
    bad = check_param_grads(dict(m.named_parameters()), g)     # norm AND 64 strided values per parameter
    assert not bad, bad[:8]


def test_benchmarked_training_regime_gradients_match_reference(golden_dir, cuda_lib):
    """config_ljs_radtts exactly as bench.py steps it (binarize_attention=True; flow + CTC + binarization losses), in
    fp32: hard map bit-identical, loss terms and all parameter gradients against the reference's autograd (golden:
    gen_radtts_train).  Covers the fused CTC gradient, the context gather's segment-sum backward, the packed flow loss."""
    g = np.load(os.path.join(golden_dir, "radtts_train.npz"))
    torch.manual_seed(0)
    model = RADTTS(**configs.model_config("radtts")).eval()   # a fresh instance: the shared fixture has been through a
    synth.load_synth(model, seed=1234)                        # train-mode forward (spectral-norm power iteration) by now
    model = model.cuda()
    b = {k: v.cuda() for k, v in synth.synth_batch(3, 52, 17, seed=97531).items()}
    crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
    model.zero_grad()
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = model(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True,
                    attn_prior=b["attn_prior"])
        losses = crit(out, b["in_lens"], b["out_lens"])
        total = sum(v * w for v, w in losses.values() if w > 0)
        bin_loss = rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"])
        total = total + bin_loss * configs.LOSS_WEIGHTS["binarization_loss_weight"]
        total.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
    assert np.array_equal(out["attn"].detach().cpu().numpy(), g["attn"])
    for key, got in (("loss_mel", losses["loss_mel"][0]), ("loss_ctc", losses["loss_ctc"][0]), ("loss_bin", bin_loss),
                     ("total", total)):
        assert abs(float(got) - float(g[key])) < 1e-3 * abs(float(g[key])), (key, float(got), float(g[key]))
    bad = check_param_grads(dict(model.named_parameters()), g)     # norm AND 64 strided values per parameter
    assert not bad, bad[:8]
