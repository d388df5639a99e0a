"""Kernel timeline of ONE replay of the captured train step (the benchmarked executable): per-kernel start / duration /
stream from CUPTI, written as a compact JSON list for offline analysis (critical path, per-stream busy time, gaps).

    python tools/graph_timeline.py gpurun_out/rNN_timeline.json
"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from radtts_b200 import configs
from radtts_b200.trainer import TrainStep

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/graph_timeline.json"
dev = torch.device("cuda", 0)
model = bench.make_model(dev).train()
ts = TrainStep(model, configs.LOSS_WEIGHTS, bf16=True, capturable=True, deferred_update=os.environ.get("DEFERRED", "0") == "1")
hb = bench.pinned_batch(32, 800, 150, seed=1000)
b = bench.to_device(hb, dev)
ts.capture(b)
for _ in range(3):
    ts.step(b)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    ts.step(b)
    torch.cuda.synchronize()
path = "/tmp/graph_trace.json"
prof.export_chrome_trace(path)
ev = json.load(open(path))["traceEvents"]
ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
rows = [[round(e["ts"] - t0, 2), round(e["dur"], 2), e["args"].get("stream", -1), e["name"][:90]] for e in ks]
json.dump(rows, open(out, "w"))
end = max(r[0] + r[1] for r in rows)
print("kernels %d, span %.3f ms, sum of durations %.3f ms" % (len(rows), end / 1e3, sum(r[1] for r in rows) / 1e3))
