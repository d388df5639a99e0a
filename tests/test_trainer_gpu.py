"""TrainStep (radtts_b200/trainer.py): the flat-gradient-buffer / direct-accumulation / CTC-prefetch machinery must
compute the same gradients as the plain autograd path it replaces.  (The captured-graph step itself is exercised by
bench.py; capturing AFTER an eager step on the legacy default stream is a known limitation)"""
import pytest
import torch

from radtts_b200 import configs, loss as rloss, ops, synth
from radtts_b200.radtts import RADTTS
from radtts_b200.trainer import TrainStep

pytestmark = pytest.mark.gpu


def _model():
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("radtts")).eval()   # eval: no dropout, no spectral-norm power iteration
    synth.load_synth(m, seed=1234)
    return m.cuda()


def test_trainstep_gradients_equal_plain_autograd(cuda_lib):
    batch = {k: v.cuda() for k, v in synth.synth_batch(4, 96, 24, seed=4242).items()}
    # ---- plain path: autocast forward, RADTTSLoss + binarization loss, backward; gradients land in fresh .grad tensors
    ref = _model()
    crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = ref(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                  binarize_attention=True, attn_prior=batch["attn_prior"])
        ld = crit(out, batch["in_lens"], batch["out_lens"])
        total = sum(v * w for v, w in ld.values() if w > 0)
        total = total + rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"]) * \
            configs.LOSS_WEIGHTS["binarization_loss_weight"]
    total.backward()
    want = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None}
    want_loss = float(total)
    del ref, out, ld, total
    try:
        # ---- TrainStep: flat gradient buffer, direct accumulation from the flow stack, CTC on the side stream
        m = _model()
        ts = TrainStep(m, configs.LOSS_WEIGHTS, bf16=True, capturable=True)
        loss = ts._fwd_bwd(batch)
        assert abs(float(loss) - want_loss) < 2e-3 * abs(want_loss), (float(loss), want_loss)
        bad = []
        for n, p in m.named_parameters():
            if n not in want:
                continue
            ref_norm = float(want[n].double().norm())
            err = float((p.grad.double() - want[n].double()).norm())
            if err > 2e-2 * ref_norm + 1e-6:
                bad.append((n, err, ref_norm))
        assert not bad, bad[:6]
    finally:
        ops.set_direct_grad_accumulation(False)
