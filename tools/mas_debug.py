import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from radtts_b200 import alignment, _lib
from oracle import mas as omas

def run(B, T1, T2, full=False):
    rng = np.random.default_rng(B * 1000003 + T1 * 1009 + T2)
    attn = rng.random((B, 1, T1, T2), dtype=np.float32) ** 2 + 1e-7
    attn /= attn.sum(3, keepdims=True)
    out_lens = rng.integers(max(1, T1 // 2), T1 + 1, B)
    in_lens = rng.integers(max(1, T2 // 2), T2 + 1, B)
    out_lens[0], in_lens[0] = T1, T2
    if full:
        out_lens[:] = T1; in_lens[:] = T2
    logp = np.log(attn)
    print("case", B, T1, T2, "lens", out_lens, in_lens, flush=True)
    try:
        hard = alignment.mas_forward(torch.from_numpy(logp).cuda(), torch.from_numpy(in_lens).cuda(), torch.from_numpy(out_lens).cuda(), is_prob=False)
        torch.cuda.synchronize()
        want = omas.binarize(logp, in_lens, out_lens, is_prob=False)
        print("  launch ok, equal:", bool(np.array_equal(hard.cpu().numpy(), want)), flush=True)
    except Exception as e:
        print("  ERR", str(e)[:120], flush=True)
        return False
    buf = (ctypes.c_ulonglong * 16)()
    _lib.check(_lib.lib().radtts_mas_debug_timeline(buf))
    t = list(buf)
    print("  err code", t[13], "warp", t[14] >> 32, "row", t[14] & 0xffffffff, "cta", t[15], flush=True)
    return True

for case in [(2, 801, 161), (8, 2000, 300), (64, 2000, 300), (64, 2000, 300, True), (64, 800, 150, True)]:
    if not run(*case):
        break
