import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radtts_b200 import alignment, _lib
for (B, T1, T2) in [(64, 2000, 300), (64, 800, 150), (64, 2000, 30), (64, 2000, 576)]:
    attn = torch.rand((B, 1, T1, T2), device="cuda").add_(1e-6)
    logp = torch.log(attn / attn.sum(3, keepdim=True))
    il = torch.full((B,), T2, dtype=torch.int64, device="cuda"); ol = torch.full((B,), T1, dtype=torch.int64, device="cuda")
    for _ in range(3):
        alignment.mas_forward(logp, il, ol, is_prob=False)
    buf = (ctypes.c_ulonglong * 16)()
    _lib.check(_lib.lib().radtts_mas_debug_timeline(buf))
    t = list(buf)
    print("   DP: %d cycles in %d ns -> %.0f MHz, %.1f cycles/row" % (t[7]-t[6], t[1]-t[0], (t[7]-t[6])/max(t[1]-t[0],1)*1e3, (t[7]-t[6])/T1))
    n=max(t[11],1); print("   last DP warp, cycles per chunk of %d rows: wait %.0f rows %.0f (%d chunks)" % (t[12], t[8]/n, t[9]/n, n))
    print("   warp 0 rows-phase cycles per chunk %.0f; last-warp group-wait spins per chunk %.2f" % (t[13]/n, t[14]/n))
    print((B, T1, T2), "dp0 %.1f us | all dp+fill %.1f | fill %.1f | backtrack %.1f | scatter %.1f" % (
        (t[1]-t[0])/1e3, (t[2]-t[0])/1e3, (t[5]-t[0])/1e3, (t[3]-t[2])/1e3, (t[4]-t[3])/1e3))
