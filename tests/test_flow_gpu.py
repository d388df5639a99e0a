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


@pytest.mark.parametrize("prec,rtol_in,rtol_w", [("fp32", 2e-3, 3e-3), ("bf16", 0.08, 0.05)])
def test_decoder_gradients_vs_reference_golden(model, gold, prec, rtol_in, rtol_w):
    """Backward of the 8-flow stack (dgrad + wgrad + coupling + LUS) against the gradients PyTorch autograd
    produced for the unmodified reference (golden g_mel, g_context, per-parameter norms and samples)."""
    ops.set_precision(prec)
    try:
        model.zero_grad(set_to_none=True)
        batch = synth.synth_batch(3, 70, 24, seed=1234)
        mel = batch["mel"].cuda().requires_grad_(True)
        ctx = torch.from_numpy(gold["context"]).cuda().requires_grad_(True)
        out_lens = batch["out_lens"].cuda()
        z, logdets, log_s = ops.decoder_forward(model, mel, ctx, out_lens)
        loss, _ = oflow.flow_loss(z, logdets, log_s, out_lens)
        loss.backward()
        tol = 1e-3 if prec == "fp32" else 3e-3
        assert abs(float(loss) - float(gold["dec_loss"])) < tol * abs(float(gold["dec_loss"]))

        def rel(a, b):
            a, b = a.double().cpu().flatten(), torch.as_tensor(b).double().flatten()
            return float((a - b).norm() / (b.norm() + 1e-30))

        r_mel, r_ctx = rel(mel.grad, gold["g_mel"]), rel(ctx.grad, gold["g_context"])
        print("rel err g_mel %.2e g_context %.2e" % (r_mel, r_ctx))
        assert r_mel < rtol_in and r_ctx < rtol_in
        params = dict(model.named_parameters())
        worst = 0.0
        for name, sums, sample in zip(gold["grad_names"], gold["grad_sums"], gold["grad_samples"]):
            name = str(name)
            g = params[name].grad
            assert g is not None, name
            if name.endswith(("invtbl_conv.lower", "invtbl_conv.upper")):
                # the reference masks these with tril/triu inside forward; so do we
                pass
            err = abs(float(g.double().norm()) - sums[1]) / (sums[1] + 1e-12)
            got = g.detach().flatten()[::1009][:64].float().cpu().numpy()
            want = sample[:len(got)]
            serr = np.linalg.norm(got - want) / (np.linalg.norm(want) + 1e-12)
            worst = max(worst, err, serr if np.linalg.norm(want) > 1e-8 else 0.0)
            assert err < rtol_w, (name, err)
            if np.linalg.norm(want) > 1e-8:
                assert serr < 4 * rtol_w, (name, serr)
        print("worst relative parameter-gradient error %.2e" % worst)
    finally:
        ops.set_precision(None)


def test_lus_stack_compose_and_backward_match_autograd(cuda_lib):
    """csrc/lus.cu: W = P L U and log|det W| of a stack of Invertible1x1ConvLUS layers (reference common.py:407-428) and
    their gradients, against the module's own torch composition under autograd."""
    from radtts_b200.common import Invertible1x1ConvLUS
    torch.manual_seed(5)
    convs = [Invertible1x1ConvLUS(c).cuda() for c in (160, 158, 33, 7)]
    for c in convs:
        with torch.no_grad():
            c.upper_diag.mul_(torch.exp(0.2 * torch.randn_like(c.upper_diag)))
    gw = [torch.randn(c.lower.shape, device="cuda") for c in convs]
    gl = [float(i + 1) * 0.7 for i in range(len(convs))]
    # reference: torch ops
    want = []
    for c, g, s in zip(convs, gw, gl):
        c.zero_grad()
        ((c.weight() * g).sum() + c.log_det() * s).backward()
        want.append((c.weight().detach(), c.log_det().detach(), c.lower.grad.clone(), c.upper.grad.clone(),
                     c.upper_diag.grad.clone()))
        c.zero_grad()
    ws, lds = ops.lus_compose_stack(convs)
    # a strided gradient for W, as the flow stack hands it over (a block of the identity-embedded matrix)
    total = 0
    for w, ld, g, s in zip(ws, lds, gw, gl):
        total = total + (w * g).sum() + ld * s
    total.backward()
    for c, w, ld, (w_ref, ld_ref, g_lo, g_up, g_ud) in zip(convs, ws, lds, want):
        assert torch.allclose(w, w_ref, rtol=1e-5, atol=1e-6)
        assert torch.allclose(ld, ld_ref, rtol=1e-5, atol=1e-6)
        assert torch.allclose(c.lower.grad, g_lo, rtol=1e-4, atol=1e-5)
        assert torch.allclose(c.upper.grad, g_up, rtol=1e-4, atol=1e-5)
        assert torch.allclose(c.upper_diag.grad, g_ud, rtol=1e-4, atol=1e-5)
