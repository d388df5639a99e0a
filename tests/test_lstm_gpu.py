"""Persistent BiLSTM kernel (csrc/lstm.cu) vs the cuDNN packed-sequence path the reference uses
(radtts.py:284-293): outputs and every gradient, with spectral norm on the recurrent weights."""
import os

import numpy as np
import pytest
import torch
from torch import nn

from radtts_b200 import lstm_ops
from radtts_b200.common import _apply_lstm_norm

pytestmark = pytest.mark.gpu


def _cudnn_path(lstm, x, lens):
    packed = nn.utils.rnn.pack_padded_sequence(x, lens.cpu(), batch_first=True, enforce_sorted=False)
    out, _ = nn.utils.rnn.pad_packed_sequence(lstm(packed)[0], batch_first=True, total_length=x.shape[1])
    return out


@pytest.mark.parametrize("In,H,B,T", [(64, 40, 5, 23), (1040, 520, 8, 61), (512, 256, 32, 37)])
def test_bilstm_matches_cudnn(cuda_lib, In, H, B, T):
    torch.manual_seed(0)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        # plain weights here: an old-style spectral-norm hook power-iterates (and mutates u, v) on every train-mode
        # call, so two calls never see the same effective weights; the spectral case is checked forward-only below
        lstm = nn.LSTM(In, H, 1, batch_first=True, bidirectional=True).cuda()
        x = (torch.randn(B, T, In, device="cuda") * 0.5).requires_grad_(True)
        lens = torch.randint(max(1, T // 3), T + 1, (B,), device="cuda")
        lens[0] = T
        w = torch.randn(B, T, 2 * H, device="cuda")
        assert lstm_ops.supported(lstm, x)
        y = lstm_ops.bilstm(lstm, x, lens)
        (y * w).sum().backward()
        got = {n: p.grad.clone() for n, p in lstm.named_parameters()}
        gx = x.grad.clone()
        lstm.zero_grad()
        x.grad = None
        y_ref = _cudnn_path(lstm, x, lens)
        (y_ref * w).sum().backward()
        assert torch.allclose(y, y_ref, rtol=1e-4, atol=2e-5), float((y - y_ref).abs().max())
        assert torch.allclose(gx, x.grad, rtol=1e-3, atol=1e-4), float((gx - x.grad).abs().max())
        for n, p in lstm.named_parameters():
            ref = p.grad
            err = float((got[n] - ref).norm() / (ref.norm() + 1e-12))
            assert err < 2e-3, (n, err)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


def test_bilstm_spectral_norm_eval_forward(cuda_lib):
    torch.manual_seed(1)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        lstm = _apply_lstm_norm(nn.LSTM(96, 72, 1, batch_first=True, bidirectional=True), "spectral").cuda()
        x = torch.randn(6, 29, 96, device="cuda") * 0.5
        lens = torch.tensor([29, 11, 20, 29, 5, 17], device="cuda")
        with torch.no_grad():
            lstm.train()
            for _ in range(5):            # let the power iteration settle, then freeze it
                _cudnn_path(lstm, x, lens)
            lstm.eval()
            y = lstm_ops.bilstm(lstm, x, lens)
            y_ref = _cudnn_path(lstm, x, lens)
        assert torch.allclose(y, y_ref, rtol=1e-4, atol=2e-5), float((y - y_ref).abs().max())
    finally:
        torch.backends.cudnn.allow_tf32 = tf32


@pytest.mark.parametrize("In,H,B,T", [(1040, 520, 32, 61), (512, 256, 19, 37), (64, 40, 5, 23)])
def test_bilstm_bf16_tensor_core_path(cuda_lib, In, H, B, T):
    """Under autocast the recurrent product runs on tensor cores with bf16 W_hh / exchanged state (fp32 accumulate,
    fp32 cell state and outputs).  Bar: outputs within 2e-2 abs of the fp32 kernel (|h| < 1), gradients within 5 %
    relative (norm) -- the same order as cuDNN's own bf16 LSTM vs fp32."""
    torch.manual_seed(0)
    lstm = nn.LSTM(In, H, 1, batch_first=True, bidirectional=True).cuda()
    x = (torch.randn(B, T, In, device="cuda") * 0.5).requires_grad_(True)
    lens = torch.randint(max(1, T // 3), T + 1, (B,), device="cuda")
    lens[0] = T
    w = torch.randn(B, T, 2 * H, device="cuda")
    y32 = lstm_ops.bilstm(lstm, x, lens)
    (y32 * w).sum().backward()
    ref = {n: p.grad.clone() for n, p in lstm.named_parameters()}
    gx32 = x.grad.clone()
    lstm.zero_grad()
    x.grad = None
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y16 = lstm_ops.bilstm(lstm, x, lens)
    assert y16.dtype == torch.float32
    (y16 * w).sum().backward()
    assert float((y16 - y32).abs().max()) < 2e-2, float((y16 - y32).abs().max())
    mask = (torch.arange(T, device="cuda")[None, :] >= lens[:, None])
    assert float(y16[mask].abs().max()) == 0.0          # packed-sequence semantics: zeros beyond each length
    assert float((x.grad - gx32).norm() / gx32.norm()) < 5e-2
    for n, p in lstm.named_parameters():
        err = float((p.grad - ref[n]).norm() / (ref[n].norm() + 1e-12))
        assert err < 5e-2, (n, err)


@pytest.mark.parametrize("In,H,B,T", [(512, 256, 32, 37), (64, 40, 5, 23)])
def test_bilstm_split_bf16_path_is_tf32_grade(cuda_lib, In, H, B, T):
    """fp32 LSTMs with cuDNN's TF32 switch on (the PyTorch default) run the recurrence with hi + lo bf16 operands
    (16 mantissa bits, fp32 accumulate) and the GEMMs around it in TF32, like cuDNN's own RNN under that switch.
    Bar: within 1e-3 abs of the exact fp32 path on outputs (|h| < 1; measured 2e-4, dominated by the TF32 input
    projection), 5e-3 relative (norm) on gradients."""
    torch.manual_seed(0)
    lstm = nn.LSTM(In, H, 1, batch_first=True, bidirectional=True).cuda()
    x = (torch.randn(B, T, In, device="cuda") * 0.5).requires_grad_(True)
    lens = torch.randint(max(1, T // 3), T + 1, (B,), device="cuda")
    lens[0] = T
    w = torch.randn(B, T, 2 * H, device="cuda")
    res = {}
    saved = torch.backends.cudnn.allow_tf32
    try:
        for flag in (False, True):
            torch.backends.cudnn.allow_tf32 = flag
            lstm.zero_grad()
            x.grad = None
            y = lstm_ops.bilstm(lstm, x, lens)
            (y * w).sum().backward()
            res[flag] = (y.detach().clone(), x.grad.clone(), {n: p.grad.clone() for n, p in lstm.named_parameters()})
    finally:
        torch.backends.cudnn.allow_tf32 = saved
    (y0, gx0, gp0), (y1, gx1, gp1) = res[False], res[True]
    assert float((y1 - y0).abs().max()) < 1e-3, float((y1 - y0).abs().max())
    # the input-projection / weight-gradient GEMMs around the recurrence are TF32 when the switch is on: 5e-3 there
    assert float((gx1 - gx0).norm() / gx0.norm()) < 5e-3
    for n in gp0:
        assert float((gp1[n] - gp0[n]).norm() / (gp0[n].norm() + 1e-12)) < 5e-3, n


@pytest.mark.parametrize("mode", ["0", "1"])
def test_all_recurrence_kernel_families_pass_the_same_checks(mode):
    """RADTTS_LSTM_CLUSTER selects the kernel family once per process (default 2: cluster forward + cooperative backward).
    Re-run this file's checks in a child process with the cooperative kernels only (0) and the cluster kernels for both
    passes (1), so that every recurrence kernel in the library stays under test."""
    import subprocess
    import sys
    env = dict(os.environ, RADTTS_LSTM_CLUSTER=mode)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-m", "gpu", "-q", "-x", "-k",
                        "not kernel_families"], capture_output=True, text=True, timeout=900, env=env,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
