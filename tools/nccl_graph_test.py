"""Can an NCCL all-reduce be captured into a CUDA graph here (side stream, thread_local capture mode)?"""
import os, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
x = torch.full((64 << 20,), float(rank + 1), device="cuda")
y = torch.zeros_like(x)
dist.all_reduce(x)                      # eager warm-up: communicator setup outside the capture
torch.cuda.synchronize()
x.fill_(float(rank + 1))
side = torch.cuda.Stream()
g = torch.cuda.CUDAGraph()
mode = sys.argv[1] if len(sys.argv) > 1 else "thread_local"
print(rank, "capturing, mode", mode, flush=True)
with torch.cuda.graph(g, capture_error_mode=mode):
    y.copy_(x)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        dist.all_reduce(y)
    z = x * 2
    torch.cuda.current_stream().wait_stream(side)
    out = y + z
print(rank, "captured", flush=True)
for it in range(3):
    x.fill_(float(rank + 1 + it))
    g.replay()
    torch.cuda.synchronize()
    want = sum(r + 1 + it for r in range(world)) + 2 * (rank + 1 + it)
    print(rank, it, float(out[0]), want, flush=True)
dist.destroy_process_group()
