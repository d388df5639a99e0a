import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from radtts_b200 import configs, loss as rloss, ops, synth
from radtts_b200.radtts import RADTTS
torch.manual_seed(0)
m = RADTTS(**configs.model_config("radtts")); synth.load_synth(m, 1234); m = m.cuda().train()
ops.set_precision("bf16")
b = {k: v.cuda() for k, v in synth.synth_batch(3, 70, 24, seed=1234).items()}
out = m(b["mel"], b["speaker_ids"], b["text"], b["in_lens"], b["out_lens"], binarize_attention=True, attn_prior=b["attn_prior"])
out["attn_soft"].register_hook(lambda g: print("g_attn_soft finite:", bool(torch.isfinite(g).all()), float(g[torch.isfinite(g)].abs().max())))
out["attn_logprob"].register_hook(lambda g: print("g_attn_logprob finite:", bool(torch.isfinite(g).all()), float(g[torch.isfinite(g)].abs().max()), int((~torch.isfinite(g)).sum())))
crit = rloss.RADTTSLoss(sigma=1.0, n_group_size=2, loss_weights=configs.LOSS_WEIGHTS)
ld = crit(out, b["in_lens"], b["out_lens"])
print({k: float(v[0]) for k, v in ld.items()})
loss = sum(v * w for v, w in ld.values()) + rloss.AttentionBinarizationLoss()(out["attn"], out["attn_soft"])
loss.backward()
for n, p in m.named_parameters():
    if n.startswith("attention."):
        print(n, None if p.grad is None else (bool(torch.isfinite(p.grad).all()), float(p.grad.abs().max())))
