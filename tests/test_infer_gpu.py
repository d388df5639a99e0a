"""GPU parity of the drop-in RADTTS.infer (sampling direction end to end, reference radtts.py:541-684) against goldens
produced by running the unmodified reference with its noise draws recorded (tests/golden/radtts_infer.npz, generator:
oracle/make_golden.py::gen_radtts_infer)."""
import os

import numpy as np
import pytest
import torch

from radtts_b200 import configs, ops, synth
from radtts_b200 import radtts as rmod
from radtts_b200.radtts import RADTTS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "radtts_infer.npz"))


def _model(name):
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config(name)).eval()
    synth.load_synth(m, seed=1234)
    return m.cuda()


class _Replay:
    """Feeds RADTTS.infer the noise tensors the reference drew, in order, checking the requested shapes."""

    def __init__(self, gold, tag):
        self.items = [torch.from_numpy(gold["%s_noise_%d" % (tag, i)]) for i in range(int(gold[tag + "_n_noise"]))]
        self.i = 0

    def __call__(self, shape, device):
        t = self.items[self.i]
        self.i += 1
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t.to(device)


def _run(model, gold, tag, monkeypatch, **kw):
    replay = _Replay(gold, tag)
    monkeypatch.setattr(rmod, "_noise", replay)
    text = torch.from_numpy(gold[tag + "_text"]).cuda()
    dur = torch.from_numpy(gold[tag + "_dur"]).cuda()
    spk = torch.zeros(text.shape[0], dtype=torch.long, device="cuda")
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            out = model.infer(spk, text, 0.8, dur=dur, **kw)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
    assert replay.i == len(replay.items), "infer drew %d noise tensors, the reference %d" % (replay.i, len(replay.items))
    return out, dur


def _valid(x, lens):
    m = (torch.arange(x.shape[-1])[None, :] < lens.cpu()[:, None]).to(x.dtype)
    return x.cpu() * (m[:, None] if x.dim() == 3 else m)


def test_infer_decoder_only_ragged_batch_matches_reference(gold, cuda_lib, monkeypatch):
    """config_ljs_radtts, B=2 with different total durations: mel on the valid frames, rtol 1e-3."""
    out, dur = _run(_model("radtts"), gold, "radtts", monkeypatch)
    lens = dur.sum(1)
    ref = torch.from_numpy(gold["radtts_mel"])
    assert out["mel"].shape == ref.shape
    got, want = _valid(out["mel"], lens), _valid(ref, lens)
    assert torch.allclose(got, want, rtol=1e-3, atol=2e-3), float((got - want).abs().max())


def test_infer_with_attribute_predictors_matches_reference(gold, cuda_lib, monkeypatch):
    """config_ljs_bgap, B=1: voicing (DAP), F0 and energy (BGAP spline flows) predicted, then the decoder."""
    m = _model("bgap")
    out, dur = _run(m, gold, "bgap", monkeypatch)
    lens = dur.sum(1)
    assert np.array_equal(out["voiced_mask"].cpu().numpy(), gold["bgap_voiced_mask"])
    for k, atol in (("f0", 2e-3), ("energy_avg", 1e-4)):
        got, want = out[k].cpu(), torch.from_numpy(gold["bgap_" + k])
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=1e-3, atol=atol), (k, float((got - want).abs().max()))
    got, want = _valid(out["mel"], lens), _valid(torch.from_numpy(gold["bgap_mel"]), lens)
    assert torch.allclose(got, want, rtol=2e-3, atol=5e-3), float((got - want).abs().max())
    # duration predictor (DAP) on recorded noise
    with torch.no_grad():
        text = torch.from_numpy(gold["bgap_text"]).cuda()
        spk_vec = m.encode_speaker(torch.zeros(1, dtype=torch.long, device="cuda"))
        txt_enc, _ = m.encode_text(text, None)
        d = m.dur_pred_layer.infer(torch.from_numpy(gold["bgap_z_dur"]).cuda(), txt_enc, spk_vec)
    assert torch.allclose(d.cpu(), torch.from_numpy(gold["bgap_dur_raw"]), rtol=1e-3, atol=1e-4)


def test_batched_predictors_equal_per_utterance_runs(cuda_lib):
    """ConvLSTMLinear (DAP) on a padded batch -- masked batched convs + the persistent LSTM kernel -- must
    equal the reference's per-utterance crop-and-loop (common.py:246-297), i.e. B=1 runs of every cropped utterance."""
    m = _model("bgap")
    torch.manual_seed(11)
    B, T = 5, 37
    lens = torch.tensor([37, 30, 22, 9, 3], device="cuda")
    txt = torch.randn(B, 512, T, device="cuda") * 0.5
    spk = torch.randn(B, 16, device="cuda") * 0.3
    for b in range(B):
        txt[b, :, lens[b]:] = 0
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False      # TF32 GEMMs round differently for different batch sizes
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            for mod in (m.dur_pred_layer, m.v_pred_module):
                batched = mod.infer(None, txt, spk, lens=lens)
                assert batched.shape[0] == B and batched.shape[2] == T
                for b in range(B):
                    n = int(lens[b])
                    single = mod.infer(None, txt[b:b + 1, :, :n], spk[b:b + 1], lens=None)
                    assert torch.allclose(batched[b:b + 1, :, :n], single, rtol=1e-4, atol=1e-5), (type(mod).__name__, b)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)


def test_batched_infer_with_in_lens_equals_per_utterance_infer(cuda_lib, monkeypatch):
    """End-to-end batched synthesis (the `in_lens` extension of RADTTS.infer; the reference cannot batch its predictors,
    SURVEY Appendix A-5): config_ljs_bgap, voicing / F0 / energy predicted, 3 padded utterances at once == three B=1 calls
    fed the same noise."""
    m = _model("bgap")
    rng = np.random.default_rng(3)
    B, T2 = 3, 12
    in_lens = torch.tensor([12, 9, 5], device="cuda")
    text = torch.from_numpy(rng.integers(1, 185, (B, T2)).astype(np.int64)).cuda()
    dur = torch.from_numpy(rng.integers(2, 7, (B, T2)).astype(np.int64)).cuda()
    for b in range(B):
        text[b, in_lens[b]:] = 0
        dur[b, in_lens[b]:] = 0
        dur[b, 0] += (4 - int(dur[b].sum()) % 4) % 4
    bank = torch.randn(B, 160, int(dur.sum(1).max()), device="cuda")
    state = {"b": None}

    def noise(shape, device):
        n, c, t = shape
        src = bank if state["b"] is None else bank[state["b"]:state["b"] + 1]
        assert src.shape[0] == n
        return src[:, :c, :t].clone()

    monkeypatch.setattr(rmod, "_noise", noise)
    spk = torch.zeros(B, dtype=torch.long, device="cuda")
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            out = m.infer(spk, text, 0.8, dur=dur, in_lens=in_lens)
            for b in range(B):
                state["b"] = b
                n = int(in_lens[b])
                one = m.infer(spk[b:b + 1], text[b:b + 1, :n], 0.8, dur=dur[b:b + 1, :n])
                t = int(dur[b].sum())
                assert torch.equal(out["voiced_mask"][b, :t], one["voiced_mask"][0, :t]), b
                for k, tol in (("f0", 2e-3), ("energy_avg", 1e-4), ("mel", 5e-3)):
                    a, r = out[k][b][..., :t], one[k][0][..., :t]
                    assert torch.allclose(a, r, rtol=2e-3, atol=tol), (k, b, float((a - r).abs().max()))
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
