"""TrainStep (radtts_b200/trainer.py): the flat-gradient-buffer / direct-accumulation / CTC-prefetch machinery must
compute the same gradients as the plain autograd path it replaces.  (The captured-graph step itself is exercised by
bench.py; capturing AFTER an eager step on the legacy default stream is a known limitation)"""
import pytest
import torch

from radtts_b200 import configs, loss as rloss, ops, synth
from radtts_b200.radtts import RADTTS
from radtts_b200.trainer import TrainStep

pytestmark = pytest.mark.gpu


def _model():
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("radtts")).eval()   # eval: no dropout, no spectral-norm power iteration
    synth.load_synth(m, seed=1234)
    return m.cuda()


def test_trainstep_gradients_equal_plain_autograd(cuda_lib):
    batch = {k: v.cuda() for k, v in synth.synth_batch(4, 96, 24, seed=4242).items()}
    # ---- plain path: autocast forward, RADTTSLoss + binarization loss, backward; gradients land in fresh .grad tensors
    ref = _model()
    crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = ref(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                  binarize_attention=True, attn_prior=batch["attn_prior"])
        ld = crit(out, batch["in_lens"], batch["out_lens"])
        total = sum(v * w for v, w in ld.values() if w > 0)
        total = total + rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"]) * \
            configs.LOSS_WEIGHTS["binarization_loss_weight"]
    total.backward()
    want = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None}
    want_loss = float(total)
    del ref, out, ld, total
    try:
        # ---- TrainStep: flat gradient buffer, direct accumulation from the flow stack, CTC on the side stream
        m = _model()
        ts = TrainStep(m, configs.LOSS_WEIGHTS, bf16=True, capturable=True)
        loss = ts._fwd_bwd(batch)
        assert abs(float(loss) - want_loss) < 2e-3 * abs(want_loss), (float(loss), want_loss)
        bad = []
        for n, p in m.named_parameters():
            if n not in want:
                continue
            ref_norm = float(want[n].double().norm())
            err = float((p.grad.double() - want[n].double()).norm())
            if err > 2e-2 * ref_norm + 1e-6:
                bad.append((n, err, ref_norm))
        assert not bad, bad[:6]
    finally:
        ops.set_direct_grad_accumulation(False)


def _lens_variant(batch, seed):
    """Same padded shape, different (valid) lengths: what changes from replay to replay in a captured step."""
    g = torch.Generator().manual_seed(seed)
    b = {k: v.clone() for k, v in batch.items()}
    B, _, T1 = b["mel"].shape
    T2 = b["text"].shape[1]
    il = torch.sort(torch.randint(T2 // 2, T2 + 1, (B,), generator=g), descending=True)[0]
    il[0] = T2
    ol = torch.maximum(torch.randint(T1 // 2, T1 + 1, (B,), generator=g), il)
    ol[int(torch.randint(0, B, (1,), generator=g))] = T1
    for i in range(B):
        b["mel"][i, :, ol[i]:] = 0
        b["text"][i, il[i]:] = 0
        pr = synth.beta_binomial_prior(int(il[i]), int(ol[i]))
        b["attn_prior"][i] = 0
        b["attn_prior"][i, :ol[i], :il[i]] = torch.from_numpy(pr)
    b["in_lens"], b["out_lens"] = il, ol
    return b


def test_graph_replayed_step_equals_eager_step(cuda_lib):
    """The benchmarked executable: TrainStep captured as ONE CUDA graph (forward, losses, backward, clip, RAdam) against
    the same TrainStep run eagerly, over 3 steps whose lengths change from step to step: loss, the flat gradient buffer
    and the parameters after the fused RAdam update."""
    base = synth.synth_batch(8, 320, 60, seed=777)
    batches = [{k: v.cuda() for k, v in _lens_variant(base, s).items()} for s in (1, 2, 3)]
    m_e, m_g = _model(), _model()         # eval mode: no dropout, no spectral-norm power iteration -> deterministic
    eager = TrainStep(m_e, configs.LOSS_WEIGHTS, bf16=True, capturable=True)
    graph = TrainStep(m_g, configs.LOSS_WEIGHTS, bf16=True, capturable=True)
    # capture() warms up with 3 eager steps on its example batch: give the eager twin the same history
    for _ in range(3):
        eager._eager_step(batches[0])
    graph.capture(batches[0])
    assert graph.graph is not None and graph.launches_per_replay > 100
    for i, b in enumerate(batches):
        le = float(eager.step(b))
        lg = float(graph.step(b))
        assert abs(le - lg) < 1e-4 * abs(le), (i, le, lg)
        # the update zeroes the gradient buffer in the pass that consumes it: compare the first moments (clipped gradient
        # history) instead
        ge, gg = eager.optimizer.exp_avg, graph.optimizer.exp_avg
        # same kernels in both; what differs is the order of fp32 atomics (weight-gradient split-K, column sums) and the
        # bf16 roundings that order flips downstream
        assert float((ge - gg).norm() / ge.norm()) < 3e-4, i
        pe, pg = eager.optimizer.flat, graph.optimizer.flat
        assert float((pe - pg).norm() / pe.norm()) < 1e-6, i
        assert int(eager.optimizer.step_dev) == int(graph.optimizer.step_dev) == 4 + i


def test_deferred_update_equals_plain_schedule(cuda_lib):
    """TrainStep(deferred_update=True) applies the flow-parameter region of every RAdam update at the top of the NEXT step
    (underneath its front end) instead of at the end of its own: same kernels on the same data, so after flush() the
    parameters, both moment buffers and the step counter equal the plain schedule's -- eagerly and as a captured graph."""
    base = synth.synth_batch(8, 320, 60, seed=778)
    batches = [{k: v.cuda() for k, v in _lens_variant(base, s).items()} for s in (4, 5, 6)]
    m_p, m_d, m_g = _model(), _model(), _model()
    plain = TrainStep(m_p, configs.LOSS_WEIGHTS, bf16=True, capturable=True)
    defer = TrainStep(m_d, configs.LOSS_WEIGHTS, bf16=True, capturable=True, deferred_update=True)
    graph = TrainStep(m_g, configs.LOSS_WEIGHTS, bf16=True, capturable=True, deferred_update=True)
    assert defer.deferred_update and defer.n_bulk > 0.9 * defer.optimizer.flat.numel()
    for _ in range(3):
        plain._eager_step(batches[0])
        defer._eager_step(batches[0])
    defer.flush()
    graph.capture(batches[0])          # warms up with 3 eager (deferred) steps on batches[0], leaves one update pending
    for i, b in enumerate(batches):
        lp, ld, lg = float(plain.step(b)), float(defer.step(b)), float(graph.step(b))
        assert abs(lp - ld) < 1e-4 * abs(lp), (i, lp, ld)
        assert abs(lp - lg) < 1e-4 * abs(lp), (i, lp, lg)
        if i == 1:
            defer.flush()              # a flush in the middle (checkpoint) must not change anything
            defer.flush()
    # before the flush the flow region lags one update behind, the remainder does not
    nb = defer.n_bulk
    assert float((plain.optimizer.flat[nb:] - defer.optimizer.flat[nb:]).norm() / plain.optimizer.flat[nb:].norm()) < 1e-6
    assert int(defer.optimizer.step_dev) == int(plain.optimizer.step_dev) - 1
    defer.flush()
    graph.flush()
    for other in (defer, graph):
        assert int(other.optimizer.step_dev) == int(plain.optimizer.step_dev) == 6
        # the moments carry the run-to-run noise of the gradients themselves (fp32 atomics order, see the graph test)
        for name, tol in (("flat", 2e-6), ("exp_avg", 5e-4), ("exp_avg_sq", 1e-3)):
            a, c = getattr(plain.optimizer, name), getattr(other.optimizer, name)
            assert float((a - c).norm() / a.norm()) < tol, (name, float((a - c).norm() / a.norm()))
        assert float(other.optimizer.grad.abs().max()) == 0.0      # every update zeroed the gradients it consumed
