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


def flow_first_order(model):
    """Flat-buffer order of the trainable parameters for data parallelism: the decoder flows' parameter networks first,
    flow by flow (94 % of the bytes; flow i's gradients are final as soon as ITS backward has run, and the flows finish in
    the order n-1 .. 0), everything else after.  Returns (parameters in that order, [(lo, hi)] flat-buffer ELEMENT range
    of every flow's region in flow order); every tensor is padded to a multiple of 4 elements exactly as FusedRAdam lays
    the buffer out, so regions[-1][1] is where the remainder starts."""
    params = [p for p in model.parameters() if p.requires_grad]
    seen, ordered, regions, off = set(), [], [], 0
    for f in getattr(model, "flows", []):
        tfn = getattr(f, "affine_tfn", None)
        net = getattr(tfn, "affine_param_predictor", None)
        lo = off
        for p in ([] if net is None else net.parameters()):
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                ordered.append(p)
                off += (p.numel() + 3) // 4 * 4
        regions.append((lo, off))
    ordered += [p for p in params if id(p) not in seen]
    return ordered, regions


def allreduce_flat(flat, regions=(), ready=None, comm_stream=None, chunk_elems=32 << 20):
    """SUM all-reduce of the flat gradient buffer (the reference does ONE flat call after backward,
    distributed.py:133-140; the mean is folded into the optimizer's scale).  `regions` = [(lo, hi)] element ranges whose
    gradients become final early, listed in the order they do; ready[i] = a CUDA event recorded when region i is final,
    True (final, nothing to wait for -- CPU / gloo tests) or None (not tracked: reduced with the remainder).  Ready
    regions are reduced first -- on `comm_stream` as their events fire, i.e. underneath the rest of the backward pass --
    and everything else follows on the current stream.  Chunks of `chunk_elems` let NCCL pipeline over NVLink / NVSwitch.
    Returns the number of collective calls."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    n_calls = 0
    ready = list(ready) if ready is not None else [None] * len(regions)
    early = [(r, ev) for r, ev in zip(regions, ready) if ev is not None and r[1] > r[0]]

    def reduce_range(lo, hi):
        n = 0
        for chunk in flat[lo:hi].split(chunk_elems):
            dist.all_reduce(chunk)
            n += 1
        return n

    if early and comm_stream is not None:
        cur = torch.cuda.current_stream(flat.device)
        with torch.cuda.stream(comm_stream):
            for (lo, hi), ev in early:
                if ev is not True:
                    comm_stream.wait_event(ev)
                n_calls += reduce_range(lo, hi)
    else:
        for (lo, hi), _ in early:
            n_calls += reduce_range(lo, hi)
    pos = 0          # the remainder: the gaps between the early ranges
    for lo, hi in sorted(r for r, _ in early) + [(flat.numel(), flat.numel())]:
        if lo > pos:
            n_calls += reduce_range(pos, lo)
        pos = max(pos, hi)
    if early and comm_stream is not None:
        cur.wait_stream(comm_stream)
    return n_calls
