"""One launch each of the MAS kernel and the ConvAttention forward / backward core at the cfg5 corner (64 x 2000 x 300), for
`ncu --set full` captures (profiles/ recipe).  Prints CUDA-event timings when run without ncu."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radtts_b200 import alignment, ops

B, T1, T2 = 64, 2000, 300
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
attn = torch.rand((B, 1, T1, T2), device=dev, generator=g).add_(1e-6)
attn = attn / attn.sum(3, keepdim=True)
logp = torch.log(attn)
il = torch.full((B,), T2, dtype=torch.int64, device=dev)
ol = torch.full((B,), T1, dtype=torch.int64, device=dev)
q = torch.randn((B, 80, T1), device=dev, generator=g).requires_grad_(True)
k = torch.randn((B, 80, T2), device=dev, generator=g).requires_grad_(True)
prior = torch.rand((B, T1, T2), device=dev, generator=g)


def once():
    alignment.mas_forward(logp, il, ol, is_prob=False)
    a, lp = ops._ConvAttnFn.apply(q, k, prior, il, 0.0005)
    (a.sum() + lp.sum() * 0.01).backward()


for _ in range(3):
    once()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
once()
e1.record()
torch.cuda.synchronize()
print("mas + attention fwd + bwd at %dx%dx%d: %.3f ms" % (B, T1, T2, e0.elapsed_time(e1)))
