"""Fused RAdam kernel vs a plain-PyTorch transcription of the reference update rule (radam.py:76-116)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_radam_step(p, g, m, v, step, lr, b1, b2, eps, wd):
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    m.mul_(b1).add_(g, alpha=1 - b1)
    b2t = b2 ** step
    nmax = 2 / (1 - b2) - 1
    nsma = nmax - 2 * step * b2t / (1 - b2t)
    if nsma >= 5:
        ss = lr * math.sqrt((1 - b2t) * (nsma - 4) / (nmax - 4) * (nsma - 2) / nsma * nmax / (nmax - 2)) / (1 - b1 ** step)
    else:
        ss = lr / (1 - b1 ** step)
    if wd:
        p.add_(p, alpha=-wd * lr)
    if nsma >= 5:
        p.addcdiv_(m, v.sqrt().add_(eps), value=-ss)
    else:
        p.add_(m, alpha=-ss)


def test_fused_radam_matches_reference_rule(cuda_lib):
    from radtts_b200.optim import FusedRAdam
    torch.manual_seed(0)
    shapes = [(37, 5), (1024,), (3, 7, 11), (1,)]
    params = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in shapes]
    ref_p = [p.detach().clone() for p in params]
    ref_m = [torch.zeros_like(p) for p in ref_p]
    ref_v = [torch.zeros_like(p) for p in ref_p]
    opt = FusedRAdam(params, lr=1e-2, weight_decay=1e-3)
    for step in range(1, 9):   # crosses the N_sma >= 5 switch (step 6 for beta2 = 0.999)
        opt.zero_grad()
        grads = [torch.randn_like(p) for p in params]
        for p, g in zip(params, grads):
            p.grad.add_(g)
        scale = opt.clip_coefficient(1.0)
        opt.step(scale)
        total = torch.sqrt(sum((g ** 2).sum() for g in grads))
        coef = min(1.0, 1.0 / (float(total) + 1e-6))
        for rp, rm, rv, g in zip(ref_p, ref_m, ref_v, grads):
            _ref_radam_step(rp, g * coef, rm, rv, step, 1e-2, 0.9, 0.999, 1e-8, 1e-3)
        for p, rp in zip(params, ref_p):
            assert torch.allclose(p.detach(), rp, rtol=1e-5, atol=1e-6), step
    assert int(opt.step_dev) == 8
