"""BASELINE.json configs[0] at full size: config_ljs_radtts, batch 4 x text 100 x 400 mel frames, the training regime
bench.py steps (binarize_attention=True; flow + CTC + binarization losses), against the golden the UNMODIFIED reference
produced on CPU (oracle/make_golden.py::gen_cfg1 -> tests/golden/cfg1_train.npz).

fp32: outputs / losses / every parameter gradient (norm AND 64 strided values) within rtol 1e-3 .. 5e-3.
bf16 (tcgen05 path under autocast): the stated looser bounds, written out below."""
import os

import numpy as np
import pytest
import torch

from radtts_b200 import configs, loss as rloss, ops, synth
from radtts_b200.radtts import RADTTS
from _gradcheck import check_param_grads

pytestmark = pytest.mark.gpu

B, T1, T2 = 4, 400, 100


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "cfg1_train.npz"))


def _model():
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("radtts")).eval()    # eval: no dropout (the golden was made the same way)
    synth.load_synth(m, seed=1234)
    return m.cuda()


def _step(model, bf16):
    b = {k: v.cuda() for k, v in synth.synth_batch(B, T1, T2, seed=20261).items()}
    crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf16):
        out = model(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True,
                    attn_prior=b["attn_prior"])
        losses = crit(out, b["in_lens"], b["out_lens"])
        total = sum(v * w for v, w in losses.values() if w > 0)
        bin_loss = rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"])
        total = total + bin_loss * configs.LOSS_WEIGHTS["binarization_loss_weight"]
    total.backward()
    return b, out, losses, bin_loss, total


def _valid(x, lens):
    m = (torch.arange(x.shape[-1], device=x.device)[None, :] < lens.to(x.device)[:, None]).to(x.dtype)
    return x * m[:, None]


def _hard_ref(gold):
    return np.unpackbits(gold["attn_packed"])[:B * T1 * T2].reshape(B, 1, T1, T2).astype(np.float32)


def test_cfg1_fp32_matches_reference(gold, cuda_lib):
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False       # SURVEY 7 trap 8b: TF32 noise is the size of the parity bar
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        model = _model()
        b, out, losses, bin_loss, total = _step(model, bf16=False)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
    # kernel 3
    soft = out["attn_soft"].detach().cpu().numpy()
    assert np.allclose(soft[:, :, ::7, ::3], gold["attn_soft_sample"], rtol=1e-3, atol=1e-7)
    assert np.allclose(soft, gold["attn_soft"].astype(np.float32), rtol=2e-3, atol=1e-6)     # fp16-stored full map
    assert np.allclose(out["attn_logprob"].detach().cpu().numpy()[:, :, ::7, ::3], gold["attn_logprob_sample"],
                       rtol=1e-3, atol=1e-3)
    # kernel 1 inside the model (probabilities in, device logf): bit-identical hard map on this batch
    hard = out["attn"].detach().cpu().numpy()
    flips = int((hard != _hard_ref(gold)).any(axis=3).sum())
    assert flips == 0, "MAS path differs from the reference on %d of %d frames" % (flips, int(b["out_lens"].sum()))
    # kernel 2
    lens = b["out_lens"] // 2
    z, z_ref = _valid(out["z_mel"].detach(), lens).cpu(), _valid(torch.from_numpy(gold["z_mel"]), lens.cpu())
    assert torch.allclose(z, z_ref, rtol=1e-3, atol=3e-4), float((z - z_ref).abs().max())
    assert np.allclose(np.array([float(x) for x in out["log_det_W_list"]]), gold["log_det_W"], rtol=1e-4, atol=1e-5)
    for i in (0, 3, 7):
        ls = _valid(out["log_s_list"][i].detach(), lens).cpu()
        ls_ref = _valid(torch.from_numpy(gold["log_s_%d" % i]), lens.cpu())
        assert torch.allclose(ls, ls_ref, rtol=1e-3, atol=1e-5), (i, float((ls - ls_ref).abs().max()))
    for key, got in (("loss_mel", losses["loss_mel"][0]), ("loss_prior_mel", losses["loss_prior_mel"][0]),
                     ("loss_ctc", losses["loss_ctc"][0]), ("loss_bin", bin_loss), ("total", total)):
        assert abs(float(got) - float(gold[key])) < 1e-3 * abs(float(gold[key])), (key, float(got), float(gold[key]))
    bad = check_param_grads(dict(model.named_parameters()), gold, rtol_norm=5e-3, rtol_elem=5e-3)
    assert not bad, (len(bad), bad[:8])


def test_cfg1_bf16_within_stated_bounds(gold, cuda_lib):
    """The benchmarked precision: bf16 WN activations / weights on tcgen05 with fp32 accumulate, z / coupling / log_s /
    1x1 conv in fp32.  Bounds: z atol 0.1 (rms ~ 2-5), log_s atol 2e-2, losses rtol 3e-3,
    parameter-gradient norms rtol 5e-2, strided gradient values 1.2e-1 relative L2."""
    model = _model()
    b, out, losses, bin_loss, total = _step(model, bf16=True)
    lens = b["out_lens"] // 2
    hard = out["attn"].detach().cpu().numpy()
    flips = int((hard != _hard_ref(gold)).any(axis=3).sum())
    n_frames = int(b["out_lens"].sum())
    assert flips <= 0.01 * n_frames, (flips, n_frames)           # bf16 attention projections may move a few boundaries
    if flips == 0:
        z, z_ref = _valid(out["z_mel"].detach(), lens).cpu(), _valid(torch.from_numpy(gold["z_mel"]), lens.cpu())
        assert float((z - z_ref).abs().max()) < 0.1
        for i in (0, 3, 7):
            ls = _valid(out["log_s_list"][i].detach(), lens).cpu()
            ls_ref = _valid(torch.from_numpy(gold["log_s_%d" % i]), lens.cpu())
            assert float((ls - ls_ref).abs().max()) < 2e-2, i
    for key, got, tol in (("loss_mel", losses["loss_mel"][0], 3e-3), ("loss_ctc", losses["loss_ctc"][0], 1e-2),
                          ("loss_bin", bin_loss, 1e-2), ("total", total, 3e-3)):
        assert abs(float(got) - float(gold[key])) < tol * abs(float(gold[key])), (key, float(got), float(gold[key]))
    if flips == 0:
        bad = check_param_grads(dict(model.named_parameters()), gold, rtol_norm=5e-2, rtol_elem=1.2e-1)
        assert not bad, (len(bad), bad[:8])
