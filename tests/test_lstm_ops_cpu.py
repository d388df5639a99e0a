"""Host-side pieces of radtts_b200/lstm_ops.py that need no GPU: the closed-form spectral-norm gradient node against
torch's own autograd through torch.nn.utils.spectral_norm (reference common.py:359-371 / radtts.py:284-293 put that norm
on weight_hh_l0 and weight_hh_l0_reverse), and the GEMM identities the LSTM wrapper's backward relies on."""
import torch

from radtts_b200 import lstm_ops


def _sn_lstm(seed=0):
    torch.manual_seed(seed)
    lstm = torch.nn.LSTM(6, 5, bidirectional=True, batch_first=True).double()
    lstm = torch.nn.utils.spectral_norm(lstm, "weight_hh_l0")
    return torch.nn.utils.spectral_norm(lstm, "weight_hh_l0_reverse")


def test_spectral_hooks_are_recognised():
    lstm = _sn_lstm()
    hooks = lstm_ops._spectral_hooks(lstm)
    assert sorted(hooks) == ["weight_hh_l0", "weight_hh_l0_reverse"]
    plain = torch.nn.LSTM(6, 5, bidirectional=True)
    assert lstm_ops._spectral_hooks(plain) is None
    wn = torch.nn.utils.weight_norm(torch.nn.LSTM(6, 5, bidirectional=True), "weight_hh_l0")
    assert lstm_ops._spectral_hooks(wn) is None          # any other hook: the generic path runs the hooks themselves


def test_closed_form_spectral_gradient_equals_autograd():
    lstm = _sn_lstm(1).eval()
    for name in ("weight_hh_l0", "weight_hh_l0_reverse"):
        G = torch.randn(20, 5, dtype=torch.float64)
        for h in lstm._forward_pre_hooks.values():
            h(lstm, ())
        W = getattr(lstm, name + "_orig")
        (getattr(lstm, name) * G).sum().backward()
        want, W.grad = W.grad.clone(), None
        u, v = getattr(lstm, name + "_u"), getattr(lstm, name + "_v")
        with torch.no_grad():
            sigma = torch.dot(u, torch.mv(W, v))
            w_eff = W / sigma
        got_w = lstm_ops._SpectralWeight.apply(W, u, v, sigma, w_eff)
        assert torch.equal(got_w, getattr(lstm, name))
        (got_w * G).sum().backward()
        assert torch.allclose(W.grad, want, rtol=1e-12, atol=1e-12)
        W.grad = None


def test_wrapper_backward_identities():
    """d_w_hh without zero-padded copies of h, the direction sum of d_x as an accumulating GEMM, bias in the GEMM."""
    torch.manual_seed(2)
    T, B, H, In = 5, 3, 4, 6
    dg = torch.randn(2, T, B, 4 * H, dtype=torch.float64)
    h_all = torch.randn(T, B, 2 * H, dtype=torch.float64)
    w_ih = torch.randn(2, 4 * H, In, dtype=torch.float64)
    x = torch.randn(T, B, In, dtype=torch.float64)
    bias = torch.randn(2, 4 * H, dtype=torch.float64)
    dg2 = dg.reshape(2, T * B, 4 * H)
    zeros = torch.zeros(1, B, H, dtype=torch.float64)
    hpf = torch.cat((zeros, h_all[:-1, :, :H]), 0).reshape(T * B, H)       # h_{t-1}, forward direction
    hpr = torch.cat((h_all[1:, :, H:], zeros), 0).reshape(T * B, H)       # h_{t+1}, reverse direction
    want = torch.stack((dg2[0].t() @ hpf, dg2[1].t() @ hpr))
    got = torch.stack((dg2[0][B:].t() @ h_all[:-1, :, :H].reshape((T - 1) * B, H),
                       dg2[1][:(T - 1) * B].t() @ h_all[1:, :, H:].reshape((T - 1) * B, H)))
    assert torch.allclose(got, want, rtol=1e-12, atol=1e-12)
    assert torch.allclose(torch.addmm(dg2[0] @ w_ih[0], dg2[1], w_ih[1]), torch.matmul(dg2, w_ih).sum(0), rtol=1e-12, atol=1e-12)
    gx = lstm_ops._input_projection(x, w_ih, bias)
    assert gx.dtype == torch.float32 or gx.dtype == torch.float64
    ref = torch.matmul(x.reshape(1, T * B, In), w_ih.transpose(1, 2)) + bias[:, None, :]
    assert torch.allclose(gx.double(), ref, rtol=1e-6, atol=1e-6)
