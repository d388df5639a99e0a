"""Elementwise gradient comparison against the committed goldens (fixture side: oracle/make_golden.py::grad_samples).

For every parameter the golden holds (sum, L2 norm) and 64 strided VALUES of the reference's gradient.  A check passes
when the norm agrees AND the strided values agree (relative L2 error of the sample vector, scaled by the gradient's RMS
so that a sample that happens to be all ~0 does not fail on noise).  A permuted or sign-flipped gradient with the right
norm fails the second part."""
import numpy as np
import torch


def strided(t):
    f = t.detach().flatten()
    sp = f[::max(1, f.numel() // 64)][:64].float().cpu().numpy()
    return np.pad(sp, (0, 64 - len(sp))), min(64, len(sp))


def check_param_grads(named_params, gold, rtol_norm=5e-3, rtol_elem=5e-3, key="grad_strided"):
    """named_params: dict name -> parameter (with .grad).  Returns the list of failures (empty = pass)."""
    bad = []
    for i, name in enumerate(gold["grad_names"]):
        name = str(name)
        p = named_params[name]
        if p.grad is None:
            bad.append((name, "no grad"))
            continue
        want_norm = float(gold["grad_sums"][i][1])
        norm = float(p.grad.double().norm())
        if abs(norm - want_norm) > rtol_norm * want_norm + 1e-7:
            bad.append((name, "norm", norm, want_norm))
            continue
        if key in gold.files:
            got, n = strided(p.grad)
            want = gold[key][i]
            rms = want_norm / max(1.0, float(p.grad.numel())) ** 0.5
            scale = max(float(np.linalg.norm(want[:n])), rms * n ** 0.5)
            err = float(np.linalg.norm(got[:n] - want[:n]))
            if err > rtol_elem * scale + 1e-7:
                bad.append((name, "values", err / max(scale, 1e-30)))
    return bad
