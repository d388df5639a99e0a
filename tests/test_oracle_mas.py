"""CPU tests: the MAS oracle (oracle/mas_oracle.c, oracle/mas.py) against reference-generated golden
vectors and -- when /root/reference is mounted -- against the reference itself."""
import os

import numpy as np
import pytest

from oracle import mas as omas
from oracle import ref_shim


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "mas_cases.npz"))


def test_c_oracle_matches_golden_prob_and_logp(gold):
    for n in range(int(gold["n_single"])):
        p, lp, hard = gold["p%d" % n], gold["logp%d" % n], gold["hard%d" % n]
        assert np.array_equal(omas.mas_width1(p, is_prob=True), hard), n
        assert np.array_equal(omas.mas_width1(lp, is_prob=False), hard), n


def test_numpy_restatement_matches_golden(gold):
    for n in range(int(gold["n_single"])):
        lp, hard = gold["logp%d" % n], gold["hard%d" % n]
        if lp.size <= 4000:
            assert np.array_equal(omas.mas_width1_numpy(lp), hard), n


def test_batch_binarize_matches_golden(gold):
    out = omas.binarize(gold["batch_attn"], gold["batch_in_lens"], gold["batch_out_lens"], is_prob=True)
    assert np.array_equal(out, gold["batch_hard"])
    out = omas.binarize(gold["batch_logp"], gold["batch_in_lens"], gold["batch_out_lens"], is_prob=False)
    assert np.array_equal(out, gold["batch_hard"])


def test_pinned_corner_cases():
    # uniform 6x3, T1<T2 (2x4) and 1x3
    u = omas.mas_width1(np.full((6, 3), 1 / 3, np.float32))
    assert u.argmax(1).tolist() == [0, 0, 0, 0, 1, 2]
    assert omas.mas_width1(np.full((2, 4), 0.25, np.float32)).tolist() == [[1, 0, 1, 0], [0, 0, 0, 1]]
    assert omas.mas_width1(np.full((1, 3), 1 / 3, np.float32)).tolist() == [[1, 0, 1]]


def test_structure_properties():
    rng = np.random.default_rng(7)
    for _ in range(20):
        t2 = int(rng.integers(2, 60))
        t1 = int(rng.integers(t2, 200))
        p = rng.random((t1, t2), dtype=np.float32) + 1e-3
        h = omas.mas_width1(p)
        assert set(np.unique(h)) <= {0.0, 1.0}
        assert (h.sum(1) == 1).all()          # one token per frame when T1 >= T2
        assert (h.sum(0) >= 1).all()          # every token gets a frame
        assert (np.diff(h.argmax(1)) >= 0).all() and (np.diff(h.argmax(1)) <= 1).all()


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")
def test_against_live_reference():
    ns = ref_shim.load()
    rng = np.random.default_rng(99)
    for trial in range(40):
        t2 = int(rng.integers(1, 50))
        t1 = int(rng.integers(1, 120))
        p = rng.random((t1, t2), dtype=np.float32) ** 3 + 1e-6
        if trial % 5 == 0:
            p[:] = 0.5
        ref = ns.alignment.mas_width1(p.copy())
        assert np.array_equal(omas.mas_width1(p), ref)
        assert np.array_equal(omas.mas_width1(omas.libm_logf(p), is_prob=False), ref)
