"""Fused text-encoder conv-block epilogue (csrc/encnorm.cu: partial-conv renormalisation + masked InstanceNorm1d + ReLU +
dropout + length mask, reference common.py:348-356) against the same math written with torch ops, values and gradients."""
import pytest
import torch
from torch.nn import functional as F

from radtts_b200 import ops
from radtts_b200.common import Encoder, get_mask_from_lengths

pytestmark = pytest.mark.gpu


def _torch_block(conv, norm, x, mask, n, drop, scale):
    y = conv(x, mask) * mask                                   # PartialConv1d mirror + ConvNorm's mask
    mean = (y * mask).sum(2, keepdim=True) / n
    var = (((y - mean) * mask) ** 2).sum(2, keepdim=True) / n
    y = (y - mean) * torch.rsqrt(var + norm.eps)
    y = y * norm.weight[None, :, None] + norm.bias[None, :, None]
    y = F.relu(y)
    if drop is not None:
        y = y * drop * scale
    return y * mask


@pytest.mark.parametrize("with_dropout", [False, True])
def test_fused_encoder_block_matches_torch(with_dropout, cuda_lib):
    torch.manual_seed(3)
    enc = Encoder(encoder_embedding_dim=512, norm_fn=torch.nn.InstanceNorm1d).cuda()
    with torch.no_grad():
        for blk in enc.convolutions:
            blk[1].weight.add_(0.1 * torch.randn_like(blk[1].weight))
            blk[1].bias.add_(0.1 * torch.randn_like(blk[1].bias))
    B, T = 5, 47
    lens = torch.tensor([47, 40, 33, 9, 2], device="cuda")
    x = torch.randn(B, 512, T, device="cuda")
    mask = get_mask_from_lengths(lens, T)[:, None].float()
    n = lens.float().clamp(min=1)[:, None, None]
    g = torch.randn(B, 512, T, device="cuda")
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for blk in enc.convolutions[:2]:
            conv, norm = blk[0].conv, blk[1]
            drop = torch.empty(B, 512, T, device="cuda").bernoulli_(0.5) if with_dropout else None
            scale = 2.0 if with_dropout else 1.0
            xr = (x * mask).clone().requires_grad_(True)
            yr = _torch_block(conv, norm, xr, mask, n, drop, scale)
            (yr * g).sum().backward()
            want = {k: p.grad.clone() for k, p in list(conv.named_parameters()) + [("gamma", norm.weight), ("beta", norm.bias)]}
            want_x = xr.grad.clone()
            conv.zero_grad(); norm.zero_grad()
            xf = (x * mask).clone().requires_grad_(True)
            raw = F.conv1d(xf, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation)
            yf = ops._EncNormFn.apply(raw, conv.bias, lens, norm.weight, norm.bias, drop, scale, conv.kernel_size[0], norm.eps)
            (yf * g).sum().backward()
            got = {k: p.grad.clone() for k, p in list(conv.named_parameters()) + [("gamma", norm.weight), ("beta", norm.bias)]}
            conv.zero_grad(); norm.zero_grad()
            assert torch.allclose(yf, yr, rtol=1e-4, atol=1e-5), float((yf - yr).abs().max())
            rel = lambda a, b: float((a - b).norm() / b.norm().clamp_min(1e-20))   # noqa: E731
            # the reference conv multiplies its input by the mask (no gradient beyond the length); the fused block is fed the
            # already-masked batch, so compare the input gradient on the valid frames
            assert rel(xf.grad * mask, want_x) < 1e-4, rel(xf.grad * mask, want_x)
            for k in want:
                if k == "bias":
                    # analytically zero (the instance norm removes a per-channel constant): both sides are rounding noise
                    assert float((got[k] - want[k]).norm()) < 1e-4 * float(want["beta"].norm()), k
                    continue
                assert rel(got[k], want[k]) < 1e-4, (k, rel(got[k], want[k]))
            x = yr.detach()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
