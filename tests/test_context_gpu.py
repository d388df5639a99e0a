"""Hard-attention context: the gather / segment-sum kernels against the reference formulation
context = bmm(text_enc, attn_hard^T) (radtts.py:399) on the hard map of the MAS kernel, values and gradient."""
import pytest
import torch

from radtts_b200 import alignment, ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T1,T2,C", [(3, 40, 12, 16), (4, 97, 35, 24), (2, 5, 9, 8), (5, 300, 150, 512)])
def test_context_gather_matches_bmm(cuda_lib, B, T1, T2, C):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + T1)
    attn = torch.rand((B, 1, T1, T2), device="cuda", generator=g) + 1e-3
    attn = attn / attn.sum(3, keepdim=True)
    out_lens = torch.randint(max(1, T1 // 2), T1 + 1, (B,), device="cuda", generator=g)
    in_lens = torch.randint(max(1, T2 // 2), T2 + 1, (B,), device="cuda", generator=g)
    out_lens[0], in_lens[0] = T1, T2
    if T1 < T2:
        in_lens[:] = T2          # T1 < T2: the path cannot start on token 0 -> two ones in row 0 (alignment.py:59)
        out_lens[:] = T1
    hard, f2t, dur = alignment.mas_forward(attn, in_lens, out_lens, is_prob=True, return_indices=True)
    text = torch.randn((B, C, T2), device="cuda", generator=g, requires_grad=True)
    w = torch.randn((B, C, T1), device="cuda", generator=g)
    ref = torch.bmm(text, hard.squeeze(1).transpose(1, 2))
    (ref * w).sum().backward()
    g_ref = text.grad.clone()
    text.grad = None
    got = ops.hard_attention_context(text, f2t)
    (got * w).sum().backward()
    assert torch.equal(got, ref) or torch.allclose(got, ref, rtol=0, atol=1e-6)
    assert torch.allclose(text.grad, g_ref, rtol=1e-5, atol=1e-5), float((text.grad - g_ref).abs().max())
    if T1 < T2:
        assert bool((hard[:, 0, 0].sum(1) == 2).all())
