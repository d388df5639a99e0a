"""Lists every implicit host<->device synchronisation in one train step (torch sync debug mode)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from radtts_b200 import configs
from radtts_b200.trainer import TrainStep
dev = torch.device("cuda", 0)
model = bench.make_model(dev).train()
ts = TrainStep(model, configs.LOSS_WEIGHTS, bf16=True)
b = bench.to_device(bench.pinned_batch(32, 800, 150, seed=1000), dev)
for _ in range(2):
    ts.step(b)
torch.cuda.synchronize()
import traceback
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    ts.step(b)
torch.cuda.set_sync_debug_mode("default")
print("sync warnings:", len(w))
seen = {}
for x in w:
    key = (x.filename, x.lineno)
    seen[key] = seen.get(key, 0) + 1
for (f, l), n in sorted(seen.items(), key=lambda kv: -kv[1]):
    print(n, f.replace(os.getcwd(), "."), l)
