"""TEST INFRASTRUCTURE ONLY -- analytic backward of the piecewise-quadratic spline (reference splines.py:254-319,
forward direction), written out term by term so that a CUDA kernel can be derived from it line by line.

The reference gets these gradients from autograd; this file restates them in closed form and is pinned against autograd
on the same formulas (tests/test_host_logic_cpu.py).  Nothing in the product package imports it.

Forward (x inside [0, 1), K bins):
    w = softmax(w~)                                   u_i = exp(v~_i - max v~) + 1e-8,  S = sum_i (u_i + u_{i+1})/2 w_i
    v = u / S                                         wc0_b = sum_{j<b} w_j,   cdf0_b = sum_{j<b} (v_j + v_{j+1})/2 w_j
    b = bin of x in wc,  alpha = (x - wc0_b) / w_b,   D = v_{b+1} - v_b,  L = v_b + alpha D
    y = alpha^2/2 D w_b + alpha v_b w_b + cdf0_b      log_j = log L
Outside [0, 1): y = x, log_j = 0.
"""
import torch


def rq_spline_forward_backward(x, w_tilde, v_tilde, g_y, g_lj):
    """x (N,), w_tilde (N, K), v_tilde (N, K+1), upstream gradients g_y, g_lj (N,).
    Returns (y, log_j, g_x, g_w_tilde, g_v_tilde) without using autograd."""
    eps = torch.finfo(x.dtype).eps
    K = w_tilde.shape[-1]
    inside = (x >= 0) & (x < 1)
    xc = torch.where(inside, x, torch.full_like(x, 0.5))
    w = torch.softmax(w_tilde, -1)
    m, am = v_tilde.max(-1, keepdim=True)
    e = torch.exp(v_tilde - m)
    u = e + 1e-8
    mid_u = (u[:, :-1] + u[:, 1:]) / 2
    S = (mid_u * w).sum(-1, keepdim=True)
    v = u / S
    mid_v = (v[:, :-1] + v[:, 1:]) / 2
    wc = torch.cumsum(w, -1)
    wc = torch.cat((wc[:, :-1], torch.ones_like(wc[:, -1:])), -1)
    wc0 = torch.nn.functional.pad(wc, (1, 0))
    cdf0 = torch.nn.functional.pad(torch.cumsum(mid_v * w, -1), (1, 0))
    b = torch.searchsorted(wc, xc[:, None]).clamp(max=K - 1)
    take = lambda t, i: torch.gather(t, -1, i).squeeze(-1)
    w_b, wc0_b, v_b, v_n, cdf0_b = take(w, b), take(wc0, b), take(v, b), take(v, b + 1), take(cdf0, b)
    alpha = (xc - wc0_b) / w_b.clamp(min=eps)
    D = v_n - v_b
    L = v_b + alpha * D
    y_in = alpha ** 2 / 2 * D * w_b + alpha * v_b * w_b + cdf0_b
    y = torch.where(inside, y_in.clamp(min=eps, max=1 - eps), x)
    log_j = torch.where(inside, L.clamp(min=eps).log(), torch.zeros_like(x))

    # ---- backward (elements outside the unit interval: identity for x, nothing for the parameters)
    gy = torch.where(inside & (y_in > eps) & (y_in < 1 - eps), g_y, torch.zeros_like(g_y))   # the clamp's gradient
    gl = torch.where(inside, g_lj, torch.zeros_like(g_lj))
    g_alpha = gy * w_b * L + gl * D / L
    g_x = torch.where(inside, g_alpha / w_b, g_y)
    # direct terms on the bin's own w, v_b, v_{b+1}
    g_w = torch.zeros_like(w)
    g_v = torch.zeros_like(v)
    g_w.scatter_add_(1, b, (gy * (alpha ** 2 / 2 * D + alpha * v_b) - g_alpha * alpha / w_b)[:, None])
    g_v.scatter_add_(1, b, (gy * (alpha - alpha ** 2 / 2) * w_b + gl * (1 - alpha) / L)[:, None])
    g_v.scatter_add_(1, b + 1, (gy * alpha ** 2 / 2 * w_b + gl * alpha / L)[:, None])
    # the two prefix sums: wc0_b (through alpha) and cdf0_b (through y), both over bins j < b
    j = torch.arange(K, device=x.device)[None, :]
    before = (j < b).to(x.dtype)                          # (N, K)
    g_w += before * (-(g_alpha / w_b))[:, None]
    g_w += before * gy[:, None] * mid_v
    half = before * gy[:, None] * w / 2
    g_v[:, :-1] += half
    g_v[:, 1:] += half
    # v = u / S,  S = sum mid_u w
    dot = (g_v * v).sum(-1, keepdim=True)
    wl = torch.nn.functional.pad(w, (1, 0))               # w_{i-1}
    wr = torch.nn.functional.pad(w, (0, 1))               # w_i
    g_u = (g_v - dot * (wl + wr) / 2) / S
    g_w += -dot * mid_v
    # u = exp(v~ - max) + 1e-8 (the max's sub-gradient goes to the arg-max entry, as autograd does)
    g_vt = g_u * e
    g_vt.scatter_add_(1, am, -(g_u * e).sum(-1, keepdim=True))
    # w = softmax(w~)
    g_wt = w * (g_w - (g_w * w).sum(-1, keepdim=True))
    return y, log_j, g_x, g_wt, g_vt


# Differentiable torch restatement of splines.py:221-319 without boolean-index gathers (torch.where instead): the autograd
# side of tests/test_host_logic_cpu.py, which pins the closed-form gradients above (and, through them, the CUDA backward
# kernel radtts_rqspline_backward).  TEST INFRASTRUCTURE ONLY.
def rq_spline_autograd(x, w_tilde, v_tilde, inverse):
    """splines.py:221-319 (unbounded piecewise-quadratic transform) without boolean-index gathers: evaluated for every
    element on a clamped copy of x and selected with torch.where, so there is no host synchronisation and the
    gradient of elements outside [0, 1) is exactly the identity's."""
    eps = torch.finfo(x.dtype).eps
    inside = (x >= 0) & (x < 1)
    xc = torch.where(inside, x, torch.full_like(x, 0.5))
    w = torch.softmax(w_tilde, dim=-1)
    v = torch.exp(v_tilde - v_tilde.max(dim=-1, keepdim=True)[0]) + 1e-8
    v = v / (((v[..., :-1] + v[..., 1:]) / 2) * w).sum(-1, keepdim=True)
    wc = torch.cumsum(w, -1)
    wc = torch.cat((wc[..., :-1], torch.ones_like(wc[..., -1:])), -1)
    wc0 = torch.nn.functional.pad(wc, (1, 0))
    cdf = torch.cumsum((v[..., 1:] + v[..., :-1]) / 2 * w, -1)
    cdf = torch.cat((cdf[..., :-1], torch.ones_like(cdf[..., -1:])), -1)
    cdf0 = torch.nn.functional.pad(cdf, (1, 0))
    knots = cdf if inverse else wc
    idx = torch.searchsorted(knots.detach(), xc.detach().unsqueeze(-1)).clamp(max=w.shape[-1] - 1)
    take = lambda t, i: torch.gather(t, -1, i).squeeze(-1)
    w_b, w_lo = take(w, idx), take(wc0, idx)
    v_b, v_n = take(v, idx), take(v, idx + 1)
    c_lo = take(cdf0, idx)
    if not inverse:
        alpha = (xc - w_lo) / w_b.clamp(min=eps)
        out = alpha ** 2 / 2 * (v_n - v_b) * w_b + alpha * v_b * w_b + c_lo
        log_j = torch.lerp(v_b, v_n, alpha).clamp(min=eps).log()
        out = out.clamp(min=eps, max=1.0 - eps)
        return torch.where(inside, out, x), torch.where(inside, log_j, torch.zeros_like(log_j))
    qa = (v_n - v_b) * w_b / 2
    qb = v_b * w_b
    qc = c_lo - xc
    alpha = (-qb + torch.sqrt(qb ** 2 - 4 * qa * qc)) / (2 * qa)
    out = (alpha * w_b + w_lo).clamp(min=eps, max=1.0 - eps)
    return torch.where(inside, out, x), None
