"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink on the B200 box, gloo in CPU tests).

The path shards by utterance: training is data parallel (batch split, model replicated, one gradient all-reduce per
step -- the reference's only strategy, reference distributed.py:101-153), inference and MAS/attention are replicas
only (no exchange)."""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and joins the process group."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return world, rank, local


def shard_range(n_items, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_items independent utterances for `rank` (replicas-only paths)."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def reduce_scalar(x, op="max", device=None):
    """max / sum of a python float over ranks (device-timed numbers are max-over-ranks, throughputs are sums)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_mean_grads(params, bucket_bytes=128 << 20, async_op=False):
    """Bucketed gradient averaging: gradients are flattened into <= bucket_bytes buffers (per dtype, in parameter
    order), all-reduced and copied back.  The reference does the same with ONE flat 904 MB buffer after backward
    (distributed.py:133-140); buckets let NCCL start on the first flows' gradients while later buckets are packed."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return []
    world = dist.get_world_size()
    buckets, cur, cur_bytes, cur_dtype = [], [], 0, None
    for p in params:
        if p.grad is None:
            continue
        nb = p.grad.numel() * p.grad.element_size()
        if cur and (cur_bytes + nb > bucket_bytes or p.grad.dtype != cur_dtype):
            buckets.append(cur)
            cur, cur_bytes = [], 0
        cur.append(p)
        cur_bytes += nb
        cur_dtype = p.grad.dtype
    if cur:
        buckets.append(cur)
    pending = []
    for bucket in buckets:
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        work = dist.all_reduce(flat, async_op=True)
        pending.append((work, flat, bucket))
    for work, flat, bucket in pending:
        work.wait()
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
    return [len(b) for b in buckets]
