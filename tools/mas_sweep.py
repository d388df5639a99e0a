"""cfg5 sweep: MAS ms/batch on the GPU (CUDA events) for batch 64, text 50-300 x mel 200-2000."""
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radtts_b200 import alignment, _lib

PEAK = 6536.4


def run(B, T1, T2, iters=20, is_prob=False):
    g = torch.Generator(device="cuda").manual_seed(0)
    attn = torch.rand((B, 1, T1, T2), device="cuda", generator=g).add_(1e-6)
    x = torch.log(attn / attn.sum(3, keepdim=True)) if not is_prob else attn / attn.sum(3, keepdim=True)
    out_lens = torch.full((B,), T1, dtype=torch.int64, device="cuda")
    in_lens = torch.full((B,), T2, dtype=torch.int64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        alignment.mas_forward(x, in_lens, out_lens, is_prob=is_prob)
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        alignment.mas_forward(x, in_lens, out_lens, is_prob=is_prob)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    gbs = B * T1 * T2 * 8 / ms / 1e6
    return {"B": B, "T1": T1, "T2": T2, "is_prob": is_prob, "ms": round(ms, 4), "GBps": round(gbs, 1),
            "frac_hbm": round(gbs / PEAK, 4)}


if __name__ == "__main__":
    for (t2, t1) in [(50, 200), (100, 400), (150, 800), (200, 1200), (300, 2000)]:
        print(json.dumps(run(64, t1, t2)))
    print(json.dumps(run(32, 800, 150)))
    print(json.dumps(run(64, 2000, 300, is_prob=True)))
