"""Kernel-level breakdown of RADTTS.infer (text -> mel, durations given) on the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from torch.autograd import DeviceType
import bench

dev = torch.device("cuda", 0)
model = bench.make_model(dev).eval()
b = bench.to_device(bench.pinned_batch(32, 800, 150, seed=1000), dev)
B, T2 = b["text"].shape
in_l, out_l = b["in_lens"].clamp(min=1), b["out_lens"]
base = (out_l // in_l)[:, None].expand(-1, T2)
tok = torch.arange(T2, device=dev)[None, :]
dur = torch.where(tok < in_l[:, None], base + (tok < (out_l - (out_l // in_l) * in_l)[:, None]).long(), torch.zeros_like(base))
spk = torch.zeros(B, dtype=torch.long, device=dev)
with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
    for _ in range(3):
        model.infer(spk, b["text"], 0.8, dur=dur)
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(5):
        model.infer(spk, b["text"], 0.8, dur=dur)
    torch.cuda.synchronize()
    print("wall ms/infer", (time.perf_counter() - t0) / 5 * 1e3)
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        model.infer(spk, b["text"], 0.8, dur=dur)
        torch.cuda.synchronize()
ks = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
ks.sort(key=lambda e: -e.device_time_total)
print("total kernel time %.2f ms, %d launches" % (sum(e.device_time_total for e in ks) / 1e3, sum(e.count for e in ks)))
for e in ks[:30]:
    print("%9.1f us %5d x %8.2f  %s" % (e.device_time_total, e.count, e.device_time_total / e.count, e.key[:110]))
cs = [e for e in prof.key_averages() if e.device_type == DeviceType.CPU]
cs.sort(key=lambda e: -e.self_cpu_time_total)
print("---- CPU self time")
for e in cs[:14]:
    print("%9.1f us %5d  %s" % (e.self_cpu_time_total, e.count, e.key[:90]))
