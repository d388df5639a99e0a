"""Parity at the BENCHMARKED shape (BASELINE.json configs[1]: batch 32 x <=800 mel frames x <=150 tokens).

At this size the packed row axis is ~10.9 k rows = 85 M-tiles x 4 N-tiles = 340 tiles over 148 persistent CTAs (3 waves,
both TMEM accumulators recycled, the TMA ring wraps across tiles), the weight-gradient kernel runs its ~850-tile problem
list, and dilation-8 taps cross real utterance gaps.  Three checks:

 1. the bf16 tcgen05 engine against the fp32 SIMT parity engine on the SAME batch (outputs, loss, every gradient),
    within the stated bf16 bounds;
 2. the fp32 engine against the CPU oracle (oracle/flow.py, pinned to the reference goldens) on 4 utterances of that
    batch, rtol 1e-3 -- and the 32-utterance packed run reproduces those 4 utterances (packing / gaps at scale);
 3. the trainer's side-stream CTC prefetch + MAS at this shape against the plain sequential path (the two
    used to share one scratch buffer)."""
import numpy as np
import pytest
import torch

from radtts_b200 import configs, loss as rloss, ops, synth
from radtts_b200.radtts import RADTTS

pytestmark = pytest.mark.gpu

B, T1, T2 = 32, 800, 150
SUB = [0, 7, 19, 31]


def _model():
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("radtts")).eval()
    sd = synth.load_synth(m, seed=1234)
    return m.cuda(), sd


def _decoder_inputs(idx=None):
    b = synth.synth_batch(B, T1, T2, seed=1000)           # bench.py's first batch
    rng = np.random.default_rng(5)
    ctx = torch.from_numpy(rng.standard_normal((B, 1040, T1 // 2), dtype=np.float32) * 0.5)
    mel, lens = b["mel"], b["out_lens"]
    if idx is not None:
        mel, ctx, lens = mel[idx], ctx[idx], lens[idx]
        tmax = int(lens.max())
        tmax -= tmax % 2
        mel, ctx = mel[:, :, :tmax].contiguous(), ctx[:, :, :tmax // 2].contiguous()
    return mel, ctx, lens


def _flow_loss(z, log_dets, log_s, lens):
    """reference loss.py:27-52 on (B, C, T') tensors."""
    tp = lens // 2
    mask = (torch.arange(z.shape[2], device=z.device)[None, :] < tp.to(z.device)[:, None])[:, None].float()
    n_el = lens.sum() // 2
    tot = 0.5 * ((z * mask) ** 2).sum()
    for ls in log_s:
        tot = tot - (ls * mask).sum()
    tot = tot - n_el * torch.stack([ld.reshape(()) for ld in log_dets]).sum()
    return tot / (n_el * z.shape[1])


def _run_gpu(model, mel, ctx, lens, prec):
    model.zero_grad(set_to_none=True)
    mel = mel.cuda().requires_grad_(True)
    ctx = ctx.cuda().requires_grad_(True)
    lens = lens.cuda()
    ops.set_precision(prec)
    try:
        z, log_dets, log_s = ops.decoder_forward(model, mel, ctx, lens)
        loss = _flow_loss(z, log_dets, log_s, lens)
        loss.backward()
    finally:
        ops.set_precision(None)
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return dict(z=z.detach(), log_s=[t.detach() for t in log_s], loss=float(loss), g_mel=mel.grad, g_ctx=ctx.grad,
                grads=grads, log_dets=[float(x) for x in log_dets])


def _valid(x, tp):
    m = (torch.arange(x.shape[-1], device=x.device)[None, :] < tp.to(x.device)[:, None]).to(x.dtype)
    return x * m[:, None]


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def full_fp32(cuda_lib):
    model, _ = _model()
    mel, ctx, lens = _decoder_inputs()
    return model, _run_gpu(model, mel, ctx, lens, "fp32"), (mel, ctx, lens)


def test_cfg2_bf16_tcgen05_engine_vs_fp32_engine(full_fp32):
    model, ref, (mel, ctx, lens) = full_fp32
    got = _run_gpu(model, mel, ctx, lens, "bf16")
    tp = lens.cuda() // 2
    assert float((_valid(got["z"], tp) - _valid(ref["z"], tp)).abs().max()) < 0.1           # z rms ~ 2-5
    for i, (a, r) in enumerate(zip(got["log_s"], ref["log_s"])):
        assert float((_valid(a, tp) - _valid(r, tp)).abs().max()) < 2e-2, i
    assert abs(got["loss"] - ref["loss"]) < 2e-3 * abs(ref["loss"]), (got["loss"], ref["loss"])
    assert _rel(got["g_mel"], ref["g_mel"]) < 8e-2
    assert _rel(got["g_ctx"], ref["g_ctx"]) < 8e-2
    bad = []
    for n, g_ref in ref["grads"].items():
        e = _rel(got["grads"][n], g_ref)
        if e > 1e-1:
            bad.append((n, e))
    assert not bad, (len(bad), bad[:8])


def test_cfg2_fp32_engine_vs_cpu_oracle_on_4_utterances(full_fp32):
    from oracle import flow as oflow
    model, full, _ = full_fp32
    _, sd = _model()
    mel, ctx, lens = _decoder_inputs(SUB)
    sub = _run_gpu(model, mel, ctx, lens, "fp32")
    # the 32-utterance packed run reproduces the 4-utterance run on those utterances (valid frames)
    tp = lens // 2
    tmax = sub["z"].shape[2]
    zf = full["z"][SUB][:, :, :tmax]
    assert torch.allclose(_valid(zf, tp.cuda()), _valid(sub["z"], tp.cuda()), rtol=1e-5, atol=1e-5)
    # CPU oracle with autograd on the same 4 utterances
    sdg = {k: (v.clone().requires_grad_(True) if k.startswith("flows.") and v.dtype.is_floating_point
               and not k.endswith((".p", "lower_diag")) else v) for k, v in sd.items()}
    mel_c, ctx_c = mel.clone().requires_grad_(True), ctx.clone().requires_grad_(True)
    z, log_dets, log_s = oflow.decoder_forward(sdg, mel_c, ctx_c, lens)
    loss = _flow_loss(z, log_dets, log_s, lens)
    loss.backward()
    assert torch.allclose(_valid(sub["z"].cpu(), tp), _valid(z.detach(), tp), rtol=1e-3, atol=3e-4)
    for i in range(8):
        assert torch.allclose(_valid(sub["log_s"][i].cpu(), tp), _valid(log_s[i].detach(), tp), rtol=1e-3, atol=1e-5), i
    assert abs(sub["loss"] - float(loss)) < 1e-3 * abs(float(loss))
    assert np.allclose(sub["log_dets"], [float(x) for x in log_dets], rtol=1e-4, atol=1e-5)
    assert _rel(sub["g_mel"].cpu(), mel_c.grad) < 2e-3
    assert _rel(sub["g_ctx"].cpu(), ctx_c.grad) < 2e-3
    bad = []
    for n, g in sub["grads"].items():
        if not n.startswith("flows."):
            continue
        e = _rel(g.cpu(), sdg[n].grad)
        if e > 2e-3:
            bad.append((n, e))
    assert not bad, (len(bad), bad[:8])


def test_cfg2_ctc_prefetch_and_mas_do_not_interfere(cuda_lib):
    """TrainStep launches the fused CTC kernel on a side stream from a ConvAttention forward hook while MAS runs on the
    main stream.  Against the plain sequential path at the bench shape: same hard map, same loss, same gradients of
    everything the CTC / binarization losses reach."""
    from radtts_b200.trainer import TrainStep
    batch = {k: v.cuda() for k, v in synth.synth_batch(B, T1, T2, seed=1000).items()}
    ref, _ = _model()
    crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = ref(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                  binarize_attention=True, attn_prior=batch["attn_prior"])
        ld = crit(out, batch["in_lens"], batch["out_lens"])
        total = sum(v * w for v, w in ld.values() if w > 0)
        total = total + rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"]) * \
            configs.LOSS_WEIGHTS["binarization_loss_weight"]
    total.backward()
    hard_ref = out["attn"].detach().clone()
    want = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None}
    want_loss = float(total)
    del out, ld, total
    try:
        m, _ = _model()
        ts = TrainStep(m, configs.LOSS_WEIGHTS, bf16=True, capturable=True)
        for _ in range(3):                                   # the race was timing dependent: a few repetitions
            ts.optimizer.zero_grad()
            tot, out = ts.forward_loss(batch)
            tot.backward()
            assert torch.equal(out["attn"], hard_ref)
            assert abs(float(tot) - want_loss) < 1e-4 * abs(want_loss), (float(tot), want_loss)
            bad = []
            for n, p in m.named_parameters():
                if n.startswith(("attention.", "embedding.", "encoder.")) and n in want:
                    e = _rel(p.grad, want[n])
                    if e > 2e-3:
                        bad.append((n, e))
            assert not bad, bad[:6]
    finally:
        ops.set_direct_grad_accumulation(False)
        ops.flow_backward_done = None
