"""Kernel-level and region-level breakdown of one eager train step (torch.profiler / CUPTI).

    python tools/profile_step.py [B] > gpurun_out/profile_step.log

Regions are record_function ranges wrapped around the forward's stages; backward time shows up under the autograd
node names (…Backward).  Library kernels (rb::*) vs everything else (at::native, cuDNN, cuBLAS) are totalled at the end."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity, record_function
import bench
from radtts_b200 import alignment, configs, ops
from radtts_b200.trainer import TrainStep

dev = torch.device("cuda", 0)
model = bench.make_model(dev).train()
ts = TrainStep(model, configs.LOSS_WEIGHTS, bf16=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
hb = bench.pinned_batch(B, 800, 150, seed=1000)
b = bench.to_device(hb, dev)


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def wrapped(*a, **k):
        with record_function("REGION:" + label):
            return fn(*a, **k)
    setattr(obj, name, wrapped)


wrap(model.encoder, "forward", "text_encoder")
wrap(model.attention, "forward", "conv_attention")
wrap(alignment, "mas_forward", "mas")
wrap(model, "preprocess_context", "preprocess_context(ctx lstm)")
wrap(ops, "decoder_forward", "decoder_forward")
wrap(ts.criterion, "forward", "criterion")
wrap(ts.bin_loss, "forward", "bin_loss")
wrap(ts, "_update", "update(clip+radam)")
wrap(ops, "hard_attention_context", "context_gather")

for _ in range(3):
    ts.step(b)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for _ in range(3):
    ts.step(b)
torch.cuda.synchronize()
print("wall ms/step (eager)", (time.perf_counter() - t0) / 3 * 1e3)
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    ts.step(b)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=140, max_name_column_width=80))
print("==== kernels only")
from torch.autograd import DeviceType
ks = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
ks.sort(key=lambda e: -e.device_time_total)
tot = sum(e.device_time_total for e in ks)
rb = sum(e.device_time_total for e in ks if "rb::" in e.key)
print("total kernel time %.2f ms, %d launches; rb:: %.2f ms (%d launches); other %.2f ms (%d launches)" % (
    tot / 1e3, sum(e.count for e in ks), rb / 1e3, sum(e.count for e in ks if "rb::" in e.key), (tot - rb) / 1e3,
    sum(e.count for e in ks if "rb::" not in e.key)))
for e in ks[:130]:
    print("%9.1f us %5d x %8.2f  %s" % (e.device_time_total, e.count, e.device_time_total / e.count, e.key[:140]))
