"""Kernel-level breakdown of one train step (torch.profiler / CUPTI) + NaN hunt for the attention grads."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from radtts_b200 import configs, loss as rloss
from radtts_b200.trainer import TrainStep

dev = torch.device("cuda", 0)
model = bench.make_model(dev).train()
ts = TrainStep(model, configs.LOSS_WEIGHTS, bf16=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hb = bench.pinned_batch(B, 800, 150, seed=1000)
b = bench.to_device(hb, dev)

# ---- NaN hunt
total, out = ts.forward_loss(b)
out["attn_soft"].register_hook(lambda g: print("g_attn_soft finite:", bool(torch.isfinite(g).all()), float(g.abs().max())))
out["attn_logprob"].register_hook(lambda g: print("g_attn_logprob finite:", bool(torch.isfinite(g).all()), float(g.abs().max())))
out["text_embeddings"].register_hook(lambda g: print("g_text_emb finite:", bool(torch.isfinite(g).all())))
total.backward()
for n, p in model.named_parameters():
    if n.startswith("attention."):
        print(n, None if p.grad is None else (bool(torch.isfinite(p.grad).all()), float(p.grad.abs().max())))
ts.optimizer.zero_grad(set_to_none=True)

for _ in range(3):
    ts.step(b)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(3):
    ts.step(b)
torch.cuda.synchronize()
print("wall ms/step", (time.perf_counter() - t0) / 3 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ts.step(b)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
print("==== kernels only")
from torch.autograd import DeviceType
ks = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
ks.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ks)
print("total kernel time %.2f ms, %d launches" % (tot / 1e3, sum(e.count for e in ks)))
for e in ks[:90]:
    print("%9.1f us %5d x %8.2f  %s" % (e.device_time_total, e.count, e.device_time_total / e.count, e.key[:120]))
print("==== CPU")
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=40, max_name_column_width=60))
