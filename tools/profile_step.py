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

# ---- attribution of the non-library kernels: which torch op / autograd node / region launched them (from the trace events)
if os.environ.get("PROFILE_ATTRIB", "1") == "1":
    import collections
    import json
    import tempfile
    path = os.path.join(tempfile.gettempdir(), "step_trace.json")
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    kern = [e for e in ev if e.get("cat") == "kernel"]
    launches = {e["args"]["correlation"]: e for e in ev if e.get("cat") == "cuda_runtime" and "correlation" in e.get("args", {})}
    cpu = [e for e in ev if e.get("cat") in ("cpu_op", "user_annotation") and "dur" in e]
    by_tid = collections.defaultdict(list)
    for e in cpu:
        by_tid[e["tid"]].append(e)
    for v in by_tid.values():
        v.sort(key=lambda e: e["ts"])
    agg = collections.defaultdict(lambda: [0.0, 0])
    for k in kern:
        if "rb::" in k["name"]:
            continue
        la = launches.get(k["args"].get("correlation"))
        chain = []
        if la is not None:
            for e in by_tid.get(la["tid"], ()):
                if e["ts"] > la["ts"]:
                    break
                if e["ts"] + e["dur"] >= la["ts"]:
                    chain.append(e["name"])
        outer = [c for c in chain if c.startswith("REGION:") or c.endswith("Backward") or "Backward" in c or c.startswith("autograd::engine")]
        key = (" > ".join(outer[-2:]) if outer else "(top)", chain[-1] if chain else "?", k["name"][:60])
        agg[key][0] += k["dur"]
        agg[key][1] += 1
    print("==== non-library kernels by launching op (us, launches)")
    tot_o = sum(v[0] for v in agg.values())
    print("total %.1f us" % tot_o)
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:120]:
        print("%9.1f us %4d x  %-60s | %-40s | %s" % (v[0], v[1], key[0][:60], key[1][:40], key[2]))
    reg = collections.defaultdict(lambda: [0.0, 0])
    for key, v in agg.items():
        reg[key[0]][0] += v[0]
        reg[key[0]][1] += v[1]
    print("==== non-library kernels by region")
    for key, v in sorted(reg.items(), key=lambda kv: -kv[1][0])[:40]:
        print("%9.1f us %4d x  %s" % (v[0], v[1], key))
