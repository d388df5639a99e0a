"""Generates the committed golden vectors under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE
(/root/reference, via oracle/ref_shim.py) on CPU in this build container.  TEST INFRASTRUCTURE ONLY.

    python oracle/make_golden.py [mas] [...]

The GPU box has no /root/reference; tests there compare against these files.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

from oracle import ref_shim  # noqa: E402


def synth_attention(rng, t1, t2, sharp=4.0):
    """Soft-attention-like rows: positive, each row sums to 1, diagonal-ish ridge plus noise."""
    i = np.arange(t1, dtype=np.float32)[:, None] / max(t1 - 1, 1)
    j = np.arange(t2, dtype=np.float32)[None, :] / max(t2 - 1, 1)
    logits = -sharp * t2 * (i - j) ** 2 + rng.normal(0, 1.0, (t1, t2)).astype(np.float32)
    logits -= logits.max(1, keepdims=True)
    p = np.exp(logits).astype(np.float32)
    return (p / p.sum(1, keepdims=True)).astype(np.float32)


def gen_mas(ns):
    from oracle import mas as omas
    rng = np.random.default_rng(1234)
    cases = {}
    shapes = [(1, 1), (1, 3), (2, 4), (6, 3), (5, 5), (17, 9), (40, 33), (64, 32), (33, 64), (97, 35), (120, 41),
              (200, 50), (160, 97)]
    for n, (t1, t2) in enumerate(shapes):
        p = synth_attention(rng, t1, t2)
        if n == 3:
            p[:] = 1.0 / t2  # uniform: every comparison is a tie
        if n == 6:
            p[7] = 0.0  # a zero-probability row (log -> -inf)
        if n == 8:
            p[:, 5] = 0.0  # a zero-probability column
        ref = ns.alignment.mas_width1(p.copy())
        cases["p%d" % n] = p
        cases["logp%d" % n] = omas.libm_logf(p)
        cases["hard%d" % n] = ref.astype(np.float32)
    # a padded batch through the reference's own binarize_attention loop (radtts.py:320-334)
    import torch
    B, T1, T2 = 5, 48, 21
    out_lens = np.array([48, 40, 31, 20, 9], dtype=np.int64)
    in_lens = np.array([21, 17, 12, 20, 3], dtype=np.int64)
    attn = np.zeros((B, 1, T1, T2), dtype=np.float32)
    for b in range(B):
        attn[b, 0, :out_lens[b], :in_lens[b]] = synth_attention(rng, int(out_lens[b]), int(in_lens[b]))
    attn[attn == 0] = 1e-30  # the reference never sees exact zeros in the padded area either way
    dummy = type("D", (), {})()
    hard = ns.radtts.RADTTS.binarize_attention(dummy, torch.from_numpy(attn), torch.from_numpy(in_lens),
                                               torch.from_numpy(out_lens)).numpy()
    cases["batch_attn"] = attn
    cases["batch_logp"] = omas.libm_logf(attn)
    cases["batch_in_lens"] = in_lens
    cases["batch_out_lens"] = out_lens
    cases["batch_hard"] = hard
    cases["n_single"] = np.array(len(shapes))
    np.savez_compressed(os.path.join(GOLD, "mas_cases.npz"), **cases)
    print("wrote mas_cases.npz", os.path.getsize(os.path.join(GOLD, "mas_cases.npz")), "bytes")


GENERATORS = {"mas": gen_mas}

if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ns = ref_shim.load()
    which = sys.argv[1:] or list(GENERATORS)
    for w in which:
        GENERATORS[w](ns)
