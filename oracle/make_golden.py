"""Generates the committed golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(/root/reference, via oracle/ref_shim.py) on CPU in this build container.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py [mas] [...]

The GPU box has no /root/reference; tests there compare against these files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_shim  # noqa: E402


def synth_attention(rng, t1, t2, sharp=4.0):
    """Soft-attention-like rows: positive, each row sums to 1, diagonal-ish ridge plus noise."""
    i = np.arange(t1, dtype=np.float32)[:, None] / max(t1 - 1, 1)
    j = np.arange(t2, dtype=np.float32)[None, :] / max(t2 - 1, 1)
    logits = -sharp * t2 * (i - j) ** 2 + rng.normal(0, 1.0, (t1, t2)).astype(np.float32)
    logits -= logits.max(1, keepdims=True)
    p = np.exp(logits).astype(np.float32)
    return (p / p.sum(1, keepdims=True)).astype(np.float32)


def gen_mas(ns):
    from oracle import mas as omas
    rng = np.random.default_rng(1234)
    cases = {}
    shapes = [(1, 1), (1, 3), (2, 4), (6, 3), (5, 5), (17, 9), (40, 33), (64, 32), (33, 64), (97, 35), (120, 41),
              (200, 50), (160, 97)]
    for n, (t1, t2) in enumerate(shapes):
        p = synth_attention(rng, t1, t2)
        if n == 3:
            p[:] = 1.0 / t2  # uniform: every comparison is a tie
        if n == 6:
            p[7] = 0.0  # a zero-probability row (log -> -inf)
        if n == 8:
            p[:, 5] = 0.0  # a zero-probability column
        ref = ns.alignment.mas_width1(p.copy())
        cases["p%d" % n] = p
        cases["logp%d" % n] = omas.libm_logf(p)
        cases["hard%d" % n] = ref.astype(np.float32)
    # a padded batch through the reference's own binarize_attention loop (radtts.py:320-334)
    import torch
    B, T1, T2 = 5, 48, 21
    out_lens = np.array([48, 40, 31, 20, 9], dtype=np.int64)
    in_lens = np.array([21, 17, 12, 20, 3], dtype=np.int64)
    attn = np.zeros((B, 1, T1, T2), dtype=np.float32)
    for b in range(B):
        attn[b, 0, :out_lens[b], :in_lens[b]] = synth_attention(rng, int(out_lens[b]), int(in_lens[b]))
    attn[attn == 0] = 1e-30  # the reference never sees exact zeros in the padded area either way
    dummy = type("D", (), {})()
    hard = ns.radtts.RADTTS.binarize_attention(dummy, torch.from_numpy(attn), torch.from_numpy(in_lens),
                                               torch.from_numpy(out_lens)).numpy()
    cases["batch_attn"] = attn
    cases["batch_logp"] = omas.libm_logf(attn)
    cases["batch_in_lens"] = in_lens
    cases["batch_out_lens"] = out_lens
    cases["batch_hard"] = hard
    cases["n_single"] = np.array(len(shapes))
    np.savez_compressed(os.path.join(GOLD, "mas_cases.npz"), **cases)
    print("wrote mas_cases.npz", os.path.getsize(os.path.join(GOLD, "mas_cases.npz")), "bytes")


def _ref_model(ns, config_name):
    import json
    import torch
    from radtts_b200 import synth
    cfg = json.load(open(os.path.join(ns.root, "configs", config_name)))
    torch.manual_seed(0)
    model = ns.radtts.RADTTS(**cfg["model_config"]).eval()
    sd = synth.synth_state_dict([(k, v.shape) for k, v in model.state_dict().items()], seed=1234)
    model.load_state_dict(sd, strict=True)   # strict: names and shapes of the product model == reference
    return model, cfg, sd


def _param_summary(t):
    f = t.detach().double().flatten()
    return np.array([f.sum().item(), f.norm().item()], dtype=np.float64), f[::1009][:64].float().numpy()


def grad_samples(t):
    """64 strided values of a gradient tensor (the whole tensor when it has <= 64 elements), zero-padded to 64: the
    fixture side of the elementwise gradient checks in tests/_gradcheck.py (same indexing there)."""
    f = t.detach().flatten()
    sp = f[::max(1, f.numel() // 64)][:64].float().numpy()
    return np.pad(sp, (0, 64 - len(sp)))


def gen_radtts_forward(ns):
    """config_ljs_radtts: full RADTTS.forward (eval mode) with binarize_attention=True, the RADTTSLoss terms,
    then -- on the captured decoder inputs -- the decoder loop's gradients and the sampling direction."""
    import torch
    from radtts_b200 import synth
    model, cfg, sd = _ref_model(ns, "config_ljs_radtts.json")
    B, T1, T2 = 3, 70, 24
    batch = synth.synth_batch(B, T1, T2, seed=1234)
    captured = {}
    def grab(module, inp, out):
        captured.setdefault("context", inp[1].detach().clone())  # hooks must return None to leave `out` alone

    h = model.flows[0].register_forward_hook(grab)
    with torch.no_grad():
        out = model(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                    binarize_attention=True, attn_prior=batch["attn_prior"])
    h.remove()
    g = {"z_mel": out["z_mel"].numpy(), "attn_soft": out["attn_soft"].numpy(), "attn": out["attn"].numpy(),
         "attn_logprob": out["attn_logprob"].numpy(), "text_embeddings": out["text_embeddings"].numpy(),
         "context": captured["context"].numpy(),
         "log_det_W": np.array([float(x) for x in out["log_det_W_list"]], dtype=np.float32)}
    for i, ls in enumerate(out["log_s_list"]):
        g["log_s_%d" % i] = ls.numpy()
    # losses (loss.py:147-203); compute_flow_loss mutates log_det_W_list[0] in place -> saved above first
    crit = ns.loss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=cfg["train_config"]["loss_weights"])
    with torch.no_grad():
        ld = crit(out, batch["in_lens"], batch["out_lens"])
        g["loss_mel"] = np.float32(ld["loss_mel"][0])
        g["loss_prior_mel"] = np.float32(ld["loss_prior_mel"][0])
        g["loss_ctc"] = np.float32(ld["loss_ctc"][0])
        g["loss_binarization"] = np.float32(ns.loss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"]))

    # ---- decoder sub-path with gradients (context and mel as leaves; reference radtts.py:414,431-444) ----
    model.zero_grad()
    mel = batch["mel"].clone().requires_grad_(True)
    ctx = captured["context"].clone().requires_grad_(True)
    z = model.unfold(mel.unsqueeze(-1))
    lens = batch["out_lens"] // 2
    z_out, log_s_list, logdets = [], [], []
    for i, flow in enumerate(model.flows):
        if i in model.exit_steps:
            z_out.append(z[:, :2])
            z = z[:, 2:]
        z, ld_w, ls = flow(z, ctx, seq_lens=lens)
        log_s_list.append(ls)
        logdets.append(ld_w)
    z_out.append(z)
    z_mel = torch.cat(z_out, 1)
    n_el = batch["out_lens"].sum() // 2
    mask = ns.common.get_mask_from_lengths(lens)[:, None].float()
    loss, _ = ns.loss.compute_flow_loss(z_mel, [x.clone() for x in logdets], log_s_list, n_el, 160, mask, 1.0)
    loss.backward()
    assert np.allclose(z_mel.detach().numpy(), g["z_mel"], atol=1e-6)
    g["dec_loss"] = np.float32(loss.item())
    g["g_mel"] = mel.grad.numpy()
    g["g_context"] = ctx.grad.numpy()
    names, sums, samples = [], [], []
    for k, p in model.named_parameters():
        if k.startswith("flows.") and p.grad is not None:
            sm, sp = _param_summary(p.grad)
            names.append(k); sums.append(sm); samples.append(np.pad(sp, (0, 64 - len(sp))))
    g["grad_names"] = np.array(names)
    g["grad_sums"] = np.stack(sums)
    g["grad_samples"] = np.stack(samples)

    # ---- sampling direction (reference radtts.py:652-677) on the same context ----
    rng = np.random.default_rng(4321)
    residual = torch.from_numpy(rng.standard_normal((B, 160, T1 // 2), dtype=np.float32) * 0.8)
    with torch.no_grad():
        stack = model.exit_steps.copy()
        x = residual[:, len(stack) * 2:]
        rest = residual[:, :len(stack) * 2]
        for i, flow in enumerate(reversed(model.flows)):
            cur = len(model.flows) - i - 1
            x = flow(x, ctx.detach(), inverse=True, seq_lens=lens)
            if stack and cur == stack[-1]:
                stack.pop()
                x = torch.cat((rest[:, len(stack) * 2:], x), 1)
                rest = rest[:, :len(stack) * 2]
        mel_inf = model.fold(x)
    g["residual"] = residual.numpy()
    g["mel_inferred"] = mel_inf.numpy()
    np.savez_compressed(os.path.join(GOLD, "radtts_forward.npz"), **g)
    print("wrote radtts_forward.npz", os.path.getsize(os.path.join(GOLD, "radtts_forward.npz")), "bytes;",
          "loss_mel %.5f ctc %.5f bin %.5f" % (g["loss_mel"], g["loss_ctc"], g["loss_binarization"]))


def gen_radtts_forward_soft(ns):
    """config_ljs_radtts with binarize_attention=False (what the reference trains with before binarization_start_iter,
    train.py:389-392): the context is bmm(text_enc, attn_soft^T), so the flow loss back-propagates into the attention
    and the text encoder.  Full forward + RADTTSLoss + backward in eval mode (no dropout); gradients of EVERY parameter."""
    import torch
    from radtts_b200 import synth
    model, cfg, sd = _ref_model(ns, "config_ljs_radtts.json")
    B, T1, T2 = 2, 44, 15
    batch = synth.synth_batch(B, T1, T2, seed=2468)
    model.zero_grad()
    out = model(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                binarize_attention=False, attn_prior=batch["attn_prior"])
    g = {"z_mel": out["z_mel"].detach().numpy(), "attn": out["attn"].detach().numpy(),
         "log_det_W": np.array([float(x) for x in out["log_det_W_list"]], dtype=np.float32)}
    crit = ns.loss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=cfg["train_config"]["loss_weights"])
    ld = crit(out, batch["in_lens"], batch["out_lens"])
    total = sum(v * w for v, w in ld.values() if w > 0)
    total.backward()
    g["loss_mel"] = np.float32(ld["loss_mel"][0].item())
    g["loss_ctc"] = np.float32(ld["loss_ctc"][0].item())
    g["total"] = np.float32(total.item())
    names, sums, samples = [], [], []
    for k, p in model.named_parameters():
        if p.grad is not None:
            sm, sp = _param_summary(p.grad)
            names.append(k); sums.append(sm); samples.append(np.pad(sp, (0, 64 - len(sp))))
    g["grad_names"] = np.array(names)
    g["grad_sums"] = np.stack(sums)
    g["grad_samples"] = np.stack(samples)
    g["grad_strided"] = np.stack([grad_samples(p.grad) for k, p in model.named_parameters() if p.grad is not None])
    np.savez_compressed(os.path.join(GOLD, "radtts_forward_soft.npz"), **g)
    print("wrote radtts_forward_soft.npz", os.path.getsize(os.path.join(GOLD, "radtts_forward_soft.npz")), "bytes;",
          "total %.5f, %d parameter gradients" % (g["total"], len(names)))


def gen_decoder_cfg_train(ns):
    """config_ljs_decoder (decoder conditioned on F0 / energy / voicing, voicing predictor, unvoiced bias) in the usual
    training regime (binarize_attention=True): RADTTSLoss incl. the vpred loss + binarization loss, backward, gradients
    of every parameter.  eval mode (no dropout)."""
    import torch
    from radtts_b200 import synth
    model, cfg, sd = _ref_model(ns, "config_ljs_decoder.json")
    mc = cfg["model_config"]
    B, T1, T2 = 2, 48, 16
    batch = synth.synth_batch(B, T1, T2, seed=1357, with_attributes=True)
    model.zero_grad()
    out = model(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                binarize_attention=True, attn_prior=batch["attn_prior"], f0=batch["f0"],
                energy_avg=batch["energy_avg"], voiced_mask=batch["voiced_mask"], p_voiced=batch["p_voiced"])
    lw = cfg["train_config"]["loss_weights"]
    crit = ns.loss.RADTTSLoss(1.0, mc["n_group_size"], mc["dur_model_config"], mc["f0_model_config"],
                              mc["energy_model_config"], mc["v_model_config"], lw)
    ld = crit(out, batch["in_lens"], batch["out_lens"])
    total = sum(v * w for v, w in ld.values() if w > 0)
    total = total + ns.loss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"]) * lw["binarization_loss_weight"]
    total.backward()
    g = {"attn": out["attn"].detach().numpy(), "total": np.float32(total.item()),
         "loss_names": np.array(sorted(ld)), "loss_values": np.array([float(ld[k][0]) for k in sorted(ld)], dtype=np.float32)}
    names, sums, samples = [], [], []
    for k, p in model.named_parameters():
        if p.grad is not None:
            names.append(k); sums.append(_param_summary(p.grad)[0]); samples.append(grad_samples(p.grad))
    g["grad_names"] = np.array(names)
    g["grad_sums"] = np.stack(sums)
    g["grad_strided"] = np.stack(samples)
    np.savez_compressed(os.path.join(GOLD, "decoder_cfg_train.npz"), **g)
    print("wrote decoder_cfg_train.npz", os.path.getsize(os.path.join(GOLD, "decoder_cfg_train.npz")), "bytes;",
          "total %.5f, losses %s, %d gradients" % (g["total"], dict(zip(g["loss_names"], g["loss_values"])), len(names)))


def gen_radtts_train(ns):
    """config_ljs_radtts in the regime bench.py times (binarize_attention=True, flow + CTC + binarization losses): the
    reference's full forward + RADTTSLoss + backward in eval mode; loss terms and gradients of every parameter."""
    import torch
    from radtts_b200 import synth
    model, cfg, sd = _ref_model(ns, "config_ljs_radtts.json")
    B, T1, T2 = 3, 52, 17
    batch = synth.synth_batch(B, T1, T2, seed=97531)
    model.zero_grad()
    out = model(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                binarize_attention=True, attn_prior=batch["attn_prior"])
    lw = cfg["train_config"]["loss_weights"]
    crit = ns.loss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=lw)
    ld = crit(out, batch["in_lens"], batch["out_lens"])
    total = sum(v * w for v, w in ld.values() if w > 0)
    bin_loss = ns.loss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"])
    total = total + bin_loss * lw["binarization_loss_weight"]
    total.backward()
    g = {"attn": out["attn"].detach().numpy(), "total": np.float32(total.item()), "loss_bin": np.float32(bin_loss.item()),
         "loss_mel": np.float32(ld["loss_mel"][0].item()), "loss_ctc": np.float32(ld["loss_ctc"][0].item())}
    names, sums, samples = [], [], []
    for k, p in model.named_parameters():
        if p.grad is not None:
            names.append(k); sums.append(_param_summary(p.grad)[0]); samples.append(grad_samples(p.grad))
    g["grad_names"] = np.array(names)
    g["grad_sums"] = np.stack(sums)
    g["grad_strided"] = np.stack(samples)
    np.savez_compressed(os.path.join(GOLD, "radtts_train.npz"), **g)
    print("wrote radtts_train.npz", os.path.getsize(os.path.join(GOLD, "radtts_train.npz")), "bytes; total %.5f, %d gradients"
          % (g["total"], len(names)))


def gen_cfg1(ns):
    """BASELINE.json configs[0]: config_ljs_radtts, batch 4 x text 100 x 400 mel frames, in the regime bench.py steps
    (binarize_attention=True; flow + CTC + binarization losses).  The unmodified reference's forward, RADTTSLoss and
    backward in eval mode (fp32, CPU): outputs, loss terms, and for EVERY parameter gradient its (sum, norm) and 64
    strided values."""
    import torch
    from radtts_b200 import synth
    model, cfg, sd = _ref_model(ns, "config_ljs_radtts.json")
    B, T1, T2 = 4, 400, 100
    batch = synth.synth_batch(B, T1, T2, seed=20261)
    model.zero_grad()
    out = model(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                binarize_attention=True, attn_prior=batch["attn_prior"])
    g = {"z_mel": out["z_mel"].detach().numpy(),
         "attn_packed": np.packbits(out["attn"].detach().numpy().astype(np.uint8), axis=None),
         "attn_soft": out["attn_soft"].detach().numpy().astype(np.float16),      # checked at 2e-3 relative only
         "attn_soft_sample": out["attn_soft"].detach().numpy()[:, :, ::7, ::3].copy(),
         "attn_logprob_sample": out["attn_logprob"].detach().numpy()[:, :, ::7, ::3].copy(),
         "log_det_W": np.array([float(x) for x in out["log_det_W_list"]], dtype=np.float32)}
    for i in (0, 3, 7):
        g["log_s_%d" % i] = out["log_s_list"][i].detach().numpy()
    lw = cfg["train_config"]["loss_weights"]
    crit = ns.loss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=lw)
    ld = crit(out, batch["in_lens"], batch["out_lens"])
    total = sum(v * w for v, w in ld.values() if w > 0)
    bin_loss = ns.loss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"])
    total = total + bin_loss * lw["binarization_loss_weight"]
    total.backward()
    g.update(total=np.float32(total.item()), loss_bin=np.float32(bin_loss.item()),
             loss_mel=np.float32(ld["loss_mel"][0].item()), loss_prior_mel=np.float32(ld["loss_prior_mel"][0].item()),
             loss_ctc=np.float32(ld["loss_ctc"][0].item()))
    names, sums, samples = [], [], []
    for k, p in model.named_parameters():
        if p.grad is not None:
            names.append(k); sums.append(_param_summary(p.grad)[0]); samples.append(grad_samples(p.grad))
    g["grad_names"] = np.array(names)
    g["grad_sums"] = np.stack(sums)
    g["grad_strided"] = np.stack(samples)
    np.savez_compressed(os.path.join(GOLD, "cfg1_train.npz"), **g)
    print("wrote cfg1_train.npz", os.path.getsize(os.path.join(GOLD, "cfg1_train.npz")), "bytes; total %.5f mel %.5f ctc "
          "%.5f bin %.5f, %d gradients" % (g["total"], g["loss_mel"], g["loss_ctc"], g["loss_bin"], len(names)))


def gen_bgap(ns):
    """config_ljs_bgap: the two BGAP attribute flows (F0: group 2, energy: group 4), sampling (infer) and training
    (forward) directions, straight through the reference modules (attribute_prediction_model.py:187-224)."""
    import torch
    model, cfg, sd = _ref_model(ns, "config_ljs_bgap.json")
    rng = np.random.default_rng(777)
    B, T = 3, 48
    lens = torch.tensor([48, 36, 28])
    txt = rng.standard_normal((B, 512, T), dtype=np.float32) * 0.5
    for b in range(B):
        txt[b, :, int(lens[b]):] = 0
    txt = torch.from_numpy(txt)
    spk = torch.from_numpy(rng.standard_normal((B, 16), dtype=np.float32) * 0.3)
    g = {"txt": txt.numpy(), "spk": spk.numpy(), "lens": lens.numpy()}
    for name, mod in (("f0", model.f0_pred_module), ("energy", model.energy_pred_module)):
        z = torch.from_numpy(rng.standard_normal((B, 2, T), dtype=np.float32) * 0.8)
        x = torch.from_numpy(rng.standard_normal((B, 2, T), dtype=np.float32) * 0.9)
        with torch.no_grad():
            x_hat = mod.infer(z, txt, spk, lens)
            out = mod(txt, spk, x, lens)
        g[name + "_z_in"] = z.numpy()
        g[name + "_x_hat"] = x_hat.numpy()
        g[name + "_x_in"] = x.numpy()
        g[name + "_z_out"] = out["z"].numpy()
        g[name + "_log_det_W"] = np.array([float(v) for v in out["log_det_W_list"]], dtype=np.float32)
        for i, ls in enumerate(out["log_s_list"]):
            g[name + "_log_s_%d" % i] = ls.numpy()
        # training direction: gradients of a flow-loss-like scalar through the reference's autograd
        mod.zero_grad()
        xg = x.clone().requires_grad_(True)
        tg = txt.clone().requires_grad_(True)
        out = mod(tg, spk, xg, lens)
        loss = 0.5 * (out["z"] ** 2).sum() - sum(ls.sum() for ls in out["log_s_list"]) \
            - 7.0 * sum(out["log_det_W_list"])
        loss.backward()
        g[name + "_train_loss"] = np.array(float(loss), dtype=np.float64)
        g[name + "_g_x"] = xg.grad.numpy()
        g[name + "_g_txt_summary"], g[name + "_g_txt_sample"] = _param_summary(tg.grad)
        names = []
        for pn, prm in mod.named_parameters():
            if prm.grad is None:
                continue
            names.append(pn)
            g["%s_gp_%d_summary" % (name, len(names) - 1)], g["%s_gp_%d_sample" % (name, len(names) - 1)] = \
                _param_summary(prm.grad)
        g[name + "_gp_names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "bgap.npz"), **g)
    print("wrote bgap.npz", os.path.getsize(os.path.join(GOLD, "bgap.npz")), "bytes")


def gen_decoder_cfg_forward(ns):
    """config_ljs_decoder (RADTTS++ decoder conditioned on F0 / energy / voicing): full forward in eval mode."""
    import torch
    from radtts_b200 import synth
    model, cfg, sd = _ref_model(ns, "config_ljs_decoder.json")
    B, T1, T2 = 2, 60, 20
    batch = synth.synth_batch(B, T1, T2, seed=4321, with_attributes=True)
    with torch.no_grad():
        out = model(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                    binarize_attention=True, attn_prior=batch["attn_prior"], f0=batch["f0"],
                    energy_avg=batch["energy_avg"], voiced_mask=batch["voiced_mask"], p_voiced=batch["p_voiced"])
    g = {"z_mel": out["z_mel"].numpy(), "attn": out["attn"].numpy(), "attn_soft": out["attn_soft"].numpy(),
         "log_det_W": np.array([float(x) for x in out["log_det_W_list"]], dtype=np.float32),
         "log_s_0": out["log_s_list"][0].numpy(), "log_s_7": out["log_s_list"][7].numpy()}
    np.savez_compressed(os.path.join(GOLD, "decoder_cfg_forward.npz"), **g)
    print("wrote decoder_cfg_forward.npz", os.path.getsize(os.path.join(GOLD, "decoder_cfg_forward.npz")), "bytes")


def gen_radtts_infer(ns):
    """RADTTS.infer (reference radtts.py:541-684), sampling direction end to end, with the noise the reference drew
    recorded (torch.Tensor.normal_ is wrapped) so that the CUDA path can be fed the same residuals.
      * config_ljs_radtts: B=2 ragged durations given (decoder-only model);
      * config_ljs_bgap:   B=1, durations given, voicing / F0 / energy predicted (DAP + BGAP flows), plus the raw output
        of the duration predictor on recorded noise (its rounding is too brittle on random weights to drive infer)."""
    import torch
    g = {}
    rec = []
    orig = torch.Tensor.normal_

    def normal_rec(self, *a, **k):
        out = orig(self, *a, **k)
        rec.append(out.detach().clone())
        return out

    for tag, cfgname, B in (("radtts", "config_ljs_radtts.json", 2), ("bgap", "config_ljs_bgap.json", 1)):
        model, cfg, sd = _ref_model(ns, cfgname)
        rng = np.random.default_rng(5 + B)
        T2 = 13
        text = torch.from_numpy(rng.integers(1, 185, (B, T2)).astype(np.int64))
        spk = torch.zeros(B, dtype=torch.long)
        dur = torch.from_numpy(rng.integers(2, 7, (B, T2)).astype(np.int64))
        for b in range(B):
            dur[b, 0] += (4 - int(dur[b].sum()) % 4) % 4      # total frames a multiple of 4
        torch.manual_seed(77)
        torch.Tensor.normal_ = normal_rec
        rec.clear()
        try:
            with torch.no_grad():
                out = model.infer(spk, text, 0.8, dur=dur)
        finally:
            torch.Tensor.normal_ = orig
        g[tag + "_text"] = text.numpy()
        g[tag + "_dur"] = dur.numpy()
        g[tag + "_n_noise"] = np.array(len(rec))
        for i, r in enumerate(rec):
            g[tag + "_noise_%d" % i] = r.numpy()
        for k in ("mel", "f0", "energy_avg", "voiced_mask"):
            if out.get(k) is not None:
                g[tag + "_" + k] = out[k].float().numpy()
        if tag == "bgap":
            with torch.no_grad():
                spk_vec = model.encode_speaker(spk)
                txt_enc, _ = model.encode_text(text, None)
                z_dur = torch.from_numpy(rng.standard_normal((B, 1, T2), dtype=np.float32) * 0.8)
                g["bgap_z_dur"] = z_dur.numpy()
                g["bgap_dur_raw"] = model.dur_pred_layer.infer(z_dur, txt_enc, spk_vec).numpy()
    np.savez_compressed(os.path.join(GOLD, "radtts_infer.npz"), **g)
    print("wrote radtts_infer.npz", os.path.getsize(os.path.join(GOLD, "radtts_infer.npz")), "bytes;",
          {k: v.shape for k, v in g.items() if k.endswith("mel")})


GENERATORS = {"cfg1": gen_cfg1, "mas": gen_mas, "radtts_infer": gen_radtts_infer, "radtts_forward_soft": gen_radtts_forward_soft, "decoder_cfg_train": gen_decoder_cfg_train, "radtts_train": gen_radtts_train, "radtts_forward": gen_radtts_forward, "bgap": gen_bgap,
              "decoder_cfg_forward": gen_decoder_cfg_forward}

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_shim.load()
    which = sys.argv[1:] or list(GENERATORS)
    for w in which:
        GENERATORS[w](ns)
