"""CPU checks of host-side pieces that are plain tensor programs (no CUDA library involved): the vectorised
LengthRegulator against the reference's per-token loop semantics (common.py:171-200), and the oracle's two spline
restatements (closed-form gradients vs the differentiable torch formulation of splines.py:221-319) against each other --
the closed form is the blueprint of the CUDA backward kernel."""
import math

import pytest
import torch

from oracle import flow as oflow
from oracle import spline_grad
from radtts_b200.common import LengthRegulator


def test_length_regulator_matches_token_loop():
    torch.manual_seed(0)
    x = torch.randn(4, 9, 6)
    dur = torch.tensor([[2, 0, 3, 1, 0, 0, 4, 1, 1], [1] * 9, [0, 0, 5, 0, 2, 0, 0, 0, 0], [0] * 9]).float()
    dur[0, 0] = 1.6                                    # rounds to 2 (int(dur + 0.5))
    out = LengthRegulator()(x, dur)
    reps = (dur + 0.5).floor().long()
    ref = torch.zeros(4, int(reps.sum(1).max()), 6)
    for b in range(4):
        rows = [x[b, j] for j in range(9) for _ in range(int(reps[b, j]))]   # the reference's loop over tokens
        if rows:
            ref[b, :len(rows)] = torch.stack(rows)
    assert out.shape == ref.shape and torch.equal(out, ref)


@pytest.mark.parametrize("inverse", [False, True])
def test_spline_autograd_formulation_matches_oracle(inverse):
    torch.manual_seed(1)
    n, h, k = 257, 3, 16
    x = torch.rand(n, h) * 1.6 - 0.3                    # some elements outside [0, 1): identity there
    w = torch.randn(n, h, k)
    v = torch.randn(n, h, k + 1)
    want, want_lj = oflow.rq_spline_unbounded(x.clone(), w, v, inverse)
    xg = x.clone().requires_grad_(True)
    wg, vg = w.clone().requires_grad_(True), v.clone().requires_grad_(True)
    got, got_lj = spline_grad.rq_spline_autograd(xg, wg, vg, inverse)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    if not inverse:
        assert torch.allclose(got_lj, want_lj, rtol=1e-5, atol=1e-6)
        (got.sum() + got_lj.sum()).backward()
    else:
        got.sum().backward()
    for t in (xg.grad, wg.grad, vg.grad):
        assert t is not None and bool(torch.isfinite(t).all())
    outside = (x < 0) | (x >= 1)
    assert torch.equal(xg.grad[outside], torch.ones_like(xg.grad[outside]))   # identity outside the unit interval
    assert float(wg.grad[outside].abs().max()) == 0.0


def test_analytic_spline_backward_matches_autograd():
    """oracle/spline_grad.py (closed-form gradients, the blueprint for a spline-backward kernel) vs autograd, float64."""
    torch.manual_seed(3)
    n, k = 400, 16
    x = (torch.rand(n, dtype=torch.float64) * 1.5 - 0.25)
    w = torch.randn(n, k, dtype=torch.float64)
    v = torch.randn(n, k + 1, dtype=torch.float64)
    gy, gl = torch.randn(n, dtype=torch.float64), torch.randn(n, dtype=torch.float64)
    xg, wg, vg = x.clone().requires_grad_(True), w.clone().requires_grad_(True), v.clone().requires_grad_(True)
    y, lj = spline_grad.rq_spline_autograd(xg[:, None], wg[:, None, :], vg[:, None, :], False)
    ((y[:, 0] * gy).sum() + (lj[:, 0] * gl).sum()).backward()
    y2, lj2, g_x, g_w, g_v = spline_grad.rq_spline_forward_backward(x, w, v, gy, gl)
    assert torch.allclose(y2, y[:, 0].detach(), rtol=1e-12, atol=1e-12)
    assert torch.allclose(lj2, lj[:, 0].detach(), rtol=1e-12, atol=1e-12)
    assert torch.allclose(g_x, xg.grad, rtol=1e-9, atol=1e-10), float((g_x - xg.grad).abs().max())
    assert torch.allclose(g_w, wg.grad, rtol=1e-8, atol=1e-10), float((g_w - wg.grad).abs().max())
    assert torch.allclose(g_v, vg.grad, rtol=1e-8, atol=1e-10), float((g_v - vg.grad).abs().max())


def test_model_configs_equal_the_reference_json_when_mounted():
    """radtts_b200.configs mirrors configs/*.json: model_config of the reference (drop-in constructor keys)."""
    import json
    import os
    from radtts_b200 import configs
    ref_dir = os.path.join(os.environ.get("RADTTS_REFERENCE", "/root/reference"), "configs")
    if not os.path.isdir(ref_dir):
        pytest.skip("reference tree not mounted")
    for name, fn in (("radtts", "config_ljs_radtts.json"), ("decoder", "config_ljs_decoder.json"),
                     ("bgap", "config_ljs_bgap.json"), ("dap", "config_ljs_dap.json")):
        with open(os.path.join(ref_dir, fn)) as f:
            ref = json.load(f)["model_config"]
        assert configs.model_config(name) == ref, name
