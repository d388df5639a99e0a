"""BGAP attribute flows (config_ljs_bgap): the CPU oracle against reference goldens (CPU test) and
the CUDA path -- SimpleConvNet row-GEMMs, RQ-spline / affine coupling kernels, small 1x1 conv -- against the same
goldens (GPU test)."""
import os

import numpy as np
import pytest
import torch

from oracle import flow as oflow
from radtts_b200 import configs, ops, synth
from radtts_b200.radtts import RADTTS

GROUP = {"f0": 2, "energy": 4}


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "bgap.npz"))


@pytest.fixture(scope="module")
def model_and_sd():
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("bgap")).eval()
    sd = synth.load_synth(m, seed=1234)
    return m, sd


def _valid(x, lens):
    m = (torch.arange(x.shape[-1])[None, :] < lens[:, None]).to(x.dtype)
    return x * m[:, None]


@pytest.mark.parametrize("name", ["f0", "energy"])
def test_oracle_bgap_matches_reference(gold, model_and_sd, name):
    _, sd = model_and_sd
    g = GROUP[name]
    txt, spk, lens = torch.from_numpy(gold["txt"]), torch.from_numpy(gold["spk"]), torch.from_numpy(gold["lens"])
    prefix = "%s_pred_module." % name
    with torch.no_grad():
        x_hat = oflow.bgap_infer(sd, prefix, torch.from_numpy(gold[name + "_z_in"]), txt, spk, lens, n_group=g)
        z, logdets, log_s = oflow.bgap_forward(sd, prefix, torch.from_numpy(gold[name + "_x_in"]), txt, spk, lens,
                                               n_group=g)
    ref = torch.from_numpy(gold[name + "_x_hat"])
    assert torch.allclose(_valid(x_hat, lens), _valid(ref, lens), rtol=1e-3, atol=1e-4)
    assert torch.allclose(_valid(z, lens // g), _valid(torch.from_numpy(gold[name + "_z_out"]), lens // g), rtol=1e-3,
                          atol=1e-4)
    assert np.allclose(np.array([float(v) for v in logdets]), gold[name + "_log_det_W"], rtol=1e-4, atol=1e-5)
    for i, ls in enumerate(log_s):
        want = torch.from_numpy(gold[name + "_log_s_%d" % i])
        assert torch.allclose(_valid(ls, lens // g), _valid(want, lens // g), rtol=1e-3, atol=1e-4), i


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["f0", "energy"])
def test_cuda_bgap_infer_and_forward_match_reference(gold, model_and_sd, name, cuda_lib):
    from radtts_b200 import ops
    model, _ = model_and_sd
    mod = getattr(model, name + "_pred_module").cuda()
    g = GROUP[name]
    txt, spk = torch.from_numpy(gold["txt"]).cuda(), torch.from_numpy(gold["spk"]).cuda()
    lens = torch.from_numpy(gold["lens"]).cuda()
    ops.set_precision("fp32")
    try:
        with torch.no_grad():
            x_hat = mod.infer(torch.from_numpy(gold[name + "_z_in"]).cuda(), txt, spk, lens)
            out = mod(txt, spk, torch.from_numpy(gold[name + "_x_in"]).cuda(), lens)
    finally:
        ops.set_precision(None)
    lc = lens.cpu()
    assert torch.allclose(_valid(x_hat.cpu(), lc), _valid(torch.from_numpy(gold[name + "_x_hat"]), lc), rtol=1e-3,
                          atol=2e-4)
    assert torch.allclose(_valid(out["z"].cpu(), lc // g), _valid(torch.from_numpy(gold[name + "_z_out"]), lc // g),
                          rtol=1e-3, atol=2e-4)
    assert np.allclose(np.array([float(v) for v in out["log_det_W_list"]]), gold[name + "_log_det_W"], rtol=1e-4,
                       atol=1e-5)
    for i, ls in enumerate(out["log_s_list"]):
        want = torch.from_numpy(gold[name + "_log_s_%d" % i])
        assert torch.allclose(_valid(ls.cpu(), lc // g), _valid(want, lc // g), rtol=1e-3, atol=2e-4), i


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["f0", "energy"])
def test_cuda_bgap_training_direction_gradients_match_reference(gold, model_and_sd, name, cuda_lib):
    """Training direction of the attribute flows: loss value, input gradients and every parameter
    gradient against the reference's autograd (golden: oracle/make_golden.py::gen_bgap)."""
    from radtts_b200 import ops
    model, _ = model_and_sd
    mod = getattr(model, name + "_pred_module").cuda()
    mod.zero_grad()
    txt = torch.from_numpy(gold["txt"]).cuda().requires_grad_(True)
    spk = torch.from_numpy(gold["spk"]).cuda()
    lens = torch.from_numpy(gold["lens"]).cuda()
    x = torch.from_numpy(gold[name + "_x_in"]).cuda().requires_grad_(True)
    ops.set_precision("fp32")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        out = mod(txt, spk, x, lens)
        loss = 0.5 * (out["z"] ** 2).sum() - sum(ls.sum() for ls in out["log_s_list"]) - 7.0 * sum(out["log_det_W_list"])
        loss.backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
    want = float(gold[name + "_train_loss"])
    assert abs(float(loss) - want) < 1e-4 * abs(want), (float(loss), want)
    gx = torch.from_numpy(gold[name + "_g_x"])
    assert torch.allclose(x.grad.cpu(), gx, rtol=2e-3, atol=2e-4), float((x.grad.cpu() - gx).abs().max())

    def check(t, summary, sample, what):
        f = t.detach().double().flatten().cpu()
        got = np.array([f.sum().item(), f.norm().item()])
        assert abs(got[1] - summary[1]) <= 2e-3 * summary[1] + 1e-6, (what, got, summary)
        assert np.allclose(f[::1009][:64].float().numpy(), sample, rtol=5e-3, atol=1e-4 * max(summary[1], 1e-3)), what

    check(txt.grad, gold[name + "_g_txt_summary"], gold[name + "_g_txt_sample"], "txt")
    params = dict(mod.named_parameters())
    for i, pn in enumerate(gold[name + "_gp_names"]):
        p = params[str(pn)]
        assert p.grad is not None, pn
        check(p.grad, gold["%s_gp_%d_summary" % (name, i)], gold["%s_gp_%d_sample" % (name, i)], str(pn))


def _torch_simple_conv_net(net, x, seq_lens):
    """SimpleConvNet.forward (reference common.py:503-515) through the ConvNorm / PartialConv1d module mirrors (torch ops)."""
    mask = (torch.arange(x.shape[2], device=x.device)[None, :] < seq_lens.to(x.device)[:, None])[:, None].to(x.dtype)
    for layer in net.layers:
        x = torch.relu(layer(x, mask))
    return net.last_layer(x)


@pytest.mark.gpu
@pytest.mark.parametrize("partial", [True, False])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_conv_stack_forward_and_backward_match_torch(partial, prec, cuda_lib):
    """radtts_conv_rows / radtts_conv_rows_backward (packed-row GEMMs, both engines) against cuDNN/ATen autograd on the
    same SimpleConvNet: output, input gradient and every weight / bias gradient, ragged lengths, dilations 1..8."""
    from radtts_b200.common import SimpleConvNet
    torch.manual_seed(7)
    net = SimpleConvNet(2, 80, 66, n_layers=4, zero_init=False, use_partial_padding=partial).cuda()
    B, T = 5, 93
    lens = torch.tensor([93, 80, 61, 33, 7], device="cuda")
    x = torch.randn(B, 82, T, device="cuda")
    g = torch.randn(B, 66, T, device="cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        xr = x.clone().requires_grad_(True)
        yr = _torch_simple_conv_net(net, xr, lens)
        (yr * g).sum().backward()
        want = {n: p.grad.clone() for n, p in net.named_parameters()}
        want_x = xr.grad.clone()
        net.zero_grad()
        # what cuDNN's own bf16 (autocast) does to the same net: the yardstick for the bf16 engine's error
        xa = x.clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ya = _torch_simple_conv_net(net, xa, lens)
        (ya.float() * g).sum().backward()
        lib16 = {n: p.grad.clone() for n, p in net.named_parameters()}
        net.zero_grad()
        ops.set_precision(prec)
        xg = x.clone().requires_grad_(True)
        y = ops.simple_conv_net(net, xg, lens)
        (y * g).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
    tol = 2e-4 if prec == "fp32" else 8e-2   # the stated bf16 gradient bound
    rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-20))   # noqa: E731
    assert rel(y.detach(), yr.detach()) < tol
    # the input gradient has been through five bf16 layers and four ReLU masks: bounded by what cuDNN's bf16 path loses
    tol_x = tol if prec == "fp32" else max(tol, 1.5 * rel(xa.grad, want_x))
    assert rel(xg.grad, want_x) < tol_x, (rel(xg.grad, want_x), rel(xa.grad, want_x))
    for n, p in net.named_parameters():
        assert p.grad is not None, n
        # same yardstick for the weight gradients of the early layers (their dy went through the later bf16 layers)
        tol_n = tol if prec == "fp32" else max(tol, 1.5 * rel(lib16[n], want[n]))
        assert rel(p.grad, want[n]) < tol_n, (n, rel(p.grad, want[n]), rel(lib16[n], want[n]))


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_attention_projections_on_library_kernels_match_torch(prec, cuda_lib):
    """ConvAttention.key_proj / query_proj (reference common.py:843-858,903-905) through ops.conv_stack vs the nn modules."""
    from radtts_b200.common import ConvAttention
    torch.manual_seed(9)
    att = ConvAttention(80, 512, 80).cuda()
    keys = torch.randn(3, 512, 37, device="cuda")
    queries = torch.randn(3, 80, 150, device="cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for seq, x in ((att.key_proj, keys), (att.query_proj, queries)):
            xr = x.clone().requires_grad_(True)
            yr = seq(xr)
            g = torch.randn_like(yr)
            (yr * g).sum().backward()
            want = {n: p.grad.clone() for n, p in seq.named_parameters()}
            seq.zero_grad()
            ops.set_precision(prec)
            xg = x.clone().requires_grad_(True)
            y = ops._projection_stack(seq, xg)
            (y * g).sum().backward()
            ops.set_precision(None)
            tol = 2e-4 if prec == "fp32" else 8e-2   # the stated bf16 gradient bound
            rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-20))   # noqa: E731
            assert rel(y.detach(), yr.detach()) < tol
            assert rel(xg.grad, xr.grad) < tol
            for n, p in seq.named_parameters():
                assert rel(p.grad, want[n]) < tol, (n, rel(p.grad, want[n]))
            seq.zero_grad()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
        ops.set_precision(None)
