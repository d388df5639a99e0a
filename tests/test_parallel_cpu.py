"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: utterance sharding, max/sum reductions used by
bench.py, and the bucketed gradient averaging (equivalent to the reference's flat all-reduce, distributed.py:116-140)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from radtts_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    w, r, _ = parallel.init_from_env("gloo")
    assert (w, r) == (world, rank)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    full_x = torch.randn(8, 16)
    full_y = torch.randn(8, 4)
    lo, hi = parallel.shard_range(8, rank, world)
    loss = torch.nn.functional.mse_loss(model(full_x[lo:hi]), full_y[lo:hi])
    loss.backward()
    sizes = parallel.allreduce_mean_grads(list(model.parameters()), bucket_bytes=1500)
    # single-process ground truth: mean of the per-shard losses
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    ref.load_state_dict(model.state_dict())
    total = 0
    for rr in range(world):
        a, b = parallel.shard_range(8, rr, world)
        total = total + torch.nn.functional.mse_loss(ref(full_x[a:b]), full_y[a:b]) / world
    total.backward()
    err = max(float((p.grad - q.grad).abs().max()) for p, q in zip(model.parameters(), ref.parameters()))
    mx = parallel.reduce_scalar(10.0 + rank, "max")
    sm = parallel.reduce_scalar(100.0 * (rank + 1), "sum")
    if rank == 0:
        out.put((err, mx, sm, sizes))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


@pytest.mark.timeout(120)
def test_two_rank_gloo_gradient_averaging_and_reductions():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err, mx, sm, sizes = out.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6
    assert mx == 11.0 and sm == 300.0
    assert len(sizes) >= 2   # more than one bucket was exercised
