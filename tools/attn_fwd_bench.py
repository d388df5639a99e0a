"""CUDA-event timing of the ConvAttention forward core (csrc/attention.cu) on the cfg5 grid corners, B = 64."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radtts_b200 import ops

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
for (B, T1, T2) in ((64, 2000, 300), (64, 2000, 150), (64, 800, 150), (64, 2000, 50), (64, 200, 50), (32, 800, 150)):
    il = torch.full((B,), T2, dtype=torch.int64, device=dev)
    q = torch.randn((B, 80, T1), device=dev, generator=g)
    k = torch.randn((B, 80, T2), device=dev, generator=g)
    prior = torch.rand((B, T1, T2), device=dev, generator=g)
    with torch.no_grad():
        for _ in range(3):
            ops._ConvAttnFn.apply(q, k, prior, il, 0.0005)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for _ in range(10):
            flush.zero_()
            e0.record()
            ops._ConvAttnFn.apply(q, k, prior, il, 0.0005)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gb = 12.0 * B * T1 * T2 / 1e9
    print("attn fwd %3d x %4d x %3d: %.4f ms  %.0f GB/s (12 B/cell)  frac of 6536 GB/s %.3f" % (B, T1, T2, ms, gb / ms * 1e3, gb / ms * 1e3 / 6536))
