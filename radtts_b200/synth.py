"""Deterministic synthetic weights and LJS-shaped batches (no dataset / checkpoint access in this project).

Everything is drawn from numpy's PCG64 (platform independent), so the build container, the GPU box and the
golden-vector generator all see bit-identical inputs.  Zero-initialised layers (WN.end, SimpleConvNet.last_layer)
are perturbed so that parity tests are not vacuous (trap 1).
"""
import math

import numpy as np
import torch


def _normal(rng, shape, std):
    return torch.from_numpy((rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32))


def synth_tensor(name, shape, rng):
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    if name.endswith("invtbl_conv.p"):
        c = shape[0]
        return torch.eye(c)[torch.from_numpy(rng.permutation(c))].contiguous()
    if name.endswith("invtbl_conv.lower_diag"):
        return torch.ones(shape)
    if name.endswith("invtbl_conv.lower"):
        return torch.tril(_normal(rng, shape, 0.05), -1)
    if name.endswith("invtbl_conv.upper"):
        return torch.triu(_normal(rng, shape, 0.05), 1)
    if name.endswith("invtbl_conv.upper_diag"):
        sign = torch.from_numpy(rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=shape))
        return sign * torch.exp(_normal(rng, shape, 0.1))
    if "convinv" in name and leaf == "weight":
        c = shape[0]
        return (torch.eye(c) + _normal(rng, (c, c), 0.1)).reshape(shape)
    if "affine_param_predictor.end." in name:
        return _normal(rng, shape, 2e-3)
    if "last_layer." in name:
        return _normal(rng, shape, 1e-2)
    if leaf == "weight_g":
        return torch.exp(_normal(rng, shape, 0.05))  # multiplied by ||v|| in synth_state_dict
    if leaf.endswith("_u") or leaf.endswith("_v") and "weight_hh" in leaf:
        v = _normal(rng, shape, 1.0)
        return v / v.norm()
    if "lstm" in name and (leaf.startswith("weight_") or leaf.startswith("bias_")):
        hidden = shape[0] // 4
        k = 1.0 / math.sqrt(hidden)
        return torch.from_numpy(rng.uniform(-k, k, size=shape).astype(np.float32))
    if "embedding" in name and leaf == "weight":
        return _normal(rng, shape, 0.5)
    if len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        return _normal(rng, shape, 1.0 / math.sqrt(fan_in))
    if leaf == "weight":  # 1-D affine scale of a norm layer
        return 1.0 + _normal(rng, shape, 0.05)
    return _normal(rng, shape, 0.02)


def synth_state_dict(named_shapes, seed=1234):
    """named_shapes: iterable of (name, shape) -- e.g. [(k, v.shape) for k, v in model.state_dict().items()].
    Returns {name: float32 CPU tensor}.  Order-independent: each tensor gets its own sub-stream."""
    out = {}
    for name, shape in named_shapes:
        sub = np.random.default_rng([seed, *[ord(ch) for ch in name]])
        out[name] = synth_tensor(name, shape, sub)
    for name in list(out):
        if name.endswith("weight_g"):
            v = out[name[:-1] + "v"]
            out[name] = out[name] * v.flatten(1).norm(dim=1).view(out[name].shape)
        if name.endswith("_orig") and name[:-5] + "_u" in out:
            # spectral norm divides by sigma = u^T W v: give it (approximate) leading singular vectors, as a trained
            # checkpoint would hold; random u, v make sigma ~ 0 and drive the LSTMs into chaotic saturation
            w = out[name].double().flatten(1)
            u = out[name[:-5] + "_u"].double()
            for _ in range(30):
                v = torch.nn.functional.normalize(w.t() @ u, dim=0)
                u = torch.nn.functional.normalize(w @ v, dim=0)
            out[name[:-5] + "_u"] = u.float()
            out[name[:-5] + "_v"] = v.float()
    return out


def load_synth(model, seed=1234):
    sd = synth_state_dict([(k, v.shape) for k, v in model.state_dict().items()], seed)
    ref = model.state_dict()
    sd = {k: v.to(ref[k].dtype) for k, v in sd.items()}
    model.load_state_dict(sd, strict=True)
    return sd


def beta_binomial_prior(n_text, n_mel, scaling=1.0):
    """Beta-binomial alignment prior (reference data.py:58-69), evaluated with lgamma instead of scipy."""
    from math import lgamma
    out = np.zeros((n_mel, n_text), dtype=np.float64)
    n = n_text - 1
    ks = np.arange(n_text)
    for i in range(1, n_mel + 1):
        a, b = scaling * i, scaling * (n_mel + 1 - i)
        logp = np.array([lgamma(n + 1) - lgamma(k + 1) - lgamma(n - k + 1) + lgamma(k + a) + lgamma(n - k + b)
                         - lgamma(n + a + b) - (lgamma(a) + lgamma(b) - lgamma(a + b)) for k in ks])
        out[i - 1] = np.exp(logp)
    return out.astype(np.float32)


def synth_batch(B, T1, T2, seed=1234, n_text=185, n_mel=80, with_attributes=False, frac_min=0.6):
    """LJS-shaped synthetic batch in the reference DataCollate layout (data.py:483-494): in_lens sorted
    descending with in_lens[0] == T2, max(out_lens) == T1, out_lens >= in_lens."""
    rng = np.random.default_rng([seed, B, T1, T2])
    in_lens = np.sort(rng.integers(max(1, int(frac_min * T2)), T2 + 1, B))[::-1].copy()
    in_lens[0] = T2
    out_lens = rng.integers(max(1, int(frac_min * T1)), T1 + 1, B)
    out_lens = np.maximum(out_lens, in_lens)
    out_lens[int(rng.integers(0, B))] = T1
    out_lens = np.minimum(out_lens, T1)
    mel = rng.standard_normal((B, n_mel, T1), dtype=np.float32)
    text = rng.integers(1, n_text, (B, T2)).astype(np.int64)
    prior = np.zeros((B, T1, T2), dtype=np.float32)
    for b in range(B):
        mel[b, :, out_lens[b]:] = 0
        text[b, in_lens[b]:] = 0
        prior[b, :out_lens[b], :in_lens[b]] = beta_binomial_prior(int(in_lens[b]), int(out_lens[b]))
    batch = {"mel": torch.from_numpy(mel), "text": torch.from_numpy(text),
             "in_lens": torch.from_numpy(in_lens.astype(np.int64)),
             "out_lens": torch.from_numpy(out_lens.astype(np.int64)),
             "attn_prior": torch.from_numpy(prior), "speaker_ids": torch.zeros(B, dtype=torch.int64)}
    if with_attributes:
        voiced = (rng.random((B, T1)) < 0.7).astype(np.float32)
        f0 = rng.uniform(80, 640, (B, T1)).astype(np.float32) * voiced
        energy = rng.random((B, T1), dtype=np.float32)
        for b in range(B):
            voiced[b, out_lens[b]:] = 0
            f0[b, out_lens[b]:] = 0
            energy[b, out_lens[b]:] = 0
        batch.update(voiced_mask=torch.from_numpy(voiced), f0=torch.from_numpy(f0),
                     energy_avg=torch.from_numpy(energy), p_voiced=torch.from_numpy(voiced.copy()))
    return batch
