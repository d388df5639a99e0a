"""GPU parity tests for kernel 2 (decoder flow stack) through the C ABI, against reference-generated goldens
(tests/golden/radtts_forward.npz) and the CPU oracle.  fp32 bar: rtol 1e-3 on valid frames;
bf16 bar: z atol 0.1, log_s atol 2e-2 (stated looser bound)."""
import os

import numpy as np
import pytest
import torch

from oracle import flow as oflow
from radtts_b200 import configs, ops, synth
from radtts_b200.radtts import RADTTS

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


def _valid(x, lens):
    m = (torch.arange(x.shape[-1], device=x.device)[None, :] < lens.to(x.device)[:, None]).to(x.dtype)
    return x * m[:, None]


def _close(a, b, rtol, atol):
    a, b = a.float().cpu(), b.float().cpu()
    ok = torch.allclose(a, b, rtol=rtol, atol=atol)
    if not ok:
        d = (a - b).abs()
        print("max abs diff %.3e at %s; ref rms %.3e" % (d.max(), np.unravel_index(int(d.argmax()), d.shape),
                                                          b.pow(2).mean().sqrt()))
    return ok


@pytest.mark.parametrize("prec,rtol,atol_z,atol_ls", [("fp32", 1e-3, 1e-4, 1e-5), ("bf16", 0.0, 0.1, 2e-2)])
def test_decoder_forward_vs_reference_golden(model, gold, prec, rtol, atol_z, atol_ls):
    ops.set_precision(prec)
    try:
        batch = synth.synth_batch(3, 70, 24, seed=1234)
        ctx = torch.from_numpy(gold["context"]).cuda()
        with torch.no_grad():
            z, logdets, log_s = ops.decoder_forward(model, batch["mel"].cuda(), ctx, batch["out_lens"].cuda())
        lens = batch["out_lens"] // 2
        assert _close(_valid(z, lens), _valid(torch.from_numpy(gold["z_mel"]), lens), rtol, atol_z)
        assert np.allclose(np.array([float(x) for x in logdets]), gold["log_det_W"], rtol=1e-4, atol=1e-5)
        for i, ls in enumerate(log_s):
            assert _close(_valid(ls, lens), _valid(torch.from_numpy(gold["log_s_%d" % i]), lens), rtol, atol_ls), i
        loss, _ = oflow.flow_loss(z.cpu(), [x.cpu() for x in logdets], [x.cpu() for x in log_s], batch["out_lens"])
        tol = 1e-3 if prec == "fp32" else 2e-3
        assert abs(float(loss) - float(gold["loss_mel"])) < tol * abs(float(gold["loss_mel"]))
    finally:
        ops.set_precision(None)


@pytest.mark.parametrize("prec,rtol,atol", [("fp32", 1e-3, 2e-4), ("bf16", 0.0, 0.15)])
def test_decoder_inverse_vs_reference_golden(model, gold, prec, rtol, atol):
    ops.set_precision(prec)
    try:
        batch = synth.synth_batch(3, 70, 24, seed=1234)
        with torch.no_grad():
            mel = ops.decoder_inverse(model, torch.from_numpy(gold["residual"]).cuda(),
                                      torch.from_numpy(gold["context"]).cuda(), batch["out_lens"].cuda())
        lens = batch["out_lens"] // 2 * 2
        assert _close(_valid(mel, lens), _valid(torch.from_numpy(gold["mel_inferred"]), lens), rtol, atol)
    finally:
        ops.set_precision(None)


def test_flowstep_module_api_and_roundtrip(model):
    """FlowStep.forward(z, context, inverse, seq_lens) on reference-shaped tensors; forward -> inverse returns
    the input (reference self-consistency 2.5e-5)."""
    ops.set_precision("fp32")
    try:
        rng = np.random.default_rng(3)
        flow = model.flows[2]  # 158 channels, c_off = 2
        B, T = 2, 45
        z = torch.from_numpy(rng.standard_normal((B, 158, T), dtype=np.float32)).cuda()
        ctx = torch.from_numpy(rng.standard_normal((B, 1040, T), dtype=np.float32) * 0.5).cuda()
        lens = torch.tensor([45, 29]).cuda()
        with torch.no_grad():
            y, log_det, log_s = flow(z, ctx, seq_lens=lens)
            back = flow(y, ctx, inverse=True, seq_lens=lens)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        with torch.no_grad():
            y_ref, ld_ref, ls_ref = oflow.flow_step(sd, "flows.2.", z.cpu(), ctx.cpu(), lens.cpu(), False)
        assert _close(_valid(y, lens), _valid(y_ref, lens.cpu()), 1e-3, 1e-4)
        assert _close(_valid(log_s, lens), _valid(ls_ref, lens.cpu()), 1e-3, 1e-5)
        assert abs(float(log_det) - float(ld_ref)) < 1e-4
        assert _close(_valid(back, lens), _valid(z, lens), 0, 1e-4)
    finally:
        ops.set_precision(None)
