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
    # the trainer's layout: every gradient is a view of ONE flat buffer; parallel.allreduce_flat sums it over ranks with
    # the early regions first (here: the two Linear layers as two "flows", listed in finishing order) and the mean folded
    # into a scale afterwards
    plist = list(model.parameters())
    flat = torch.cat([p.grad.reshape(-1) for p in plist])
    n0 = plist[0].numel() + plist[1].numel()
    sizes = parallel.allreduce_flat(flat, regions=[(n0, flat.numel() - 4), (0, n0)], ready=[True, None], chunk_elems=100)
    flat /= world
    off = 0
    for p in plist:
        p.grad.copy_(flat[off:off + p.numel()].view_as(p.grad))
        off += p.numel()
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
    assert sizes >= 4        # early region + remainder, several chunks each


def test_bulk_first_order_is_a_permutation_with_the_flow_networks_in_front():
    """trainer.bulk_first_order: the flat gradient buffer of the data-parallel step is [flow WN parameters | rest]; the
    first region is what the trainer all-reduces while the rest of the backward pass still runs."""
    from radtts_b200 import configs
    from radtts_b200.radtts import RADTTS
    from radtts_b200.trainer import bulk_first_order
    torch.manual_seed(0)
    model = RADTTS(**configs.model_config("radtts"))
    ordered, n_bulk = bulk_first_order(model)
    trainable = [p for p in model.parameters() if p.requires_grad]
    assert len(ordered) == len(trainable) and {id(p) for p in ordered} == {id(p) for p in trainable}
    wn_ids = {id(p) for f in model.flows for p in f.affine_tfn.affine_param_predictor.parameters()}
    k = len(wn_ids)
    assert {id(p) for p in ordered[:k]} == wn_ids                      # the front region is exactly the flow networks
    assert n_bulk == sum((p.numel() + 3) // 4 * 4 for p in ordered[:k])   # FusedRAdam's 16-byte padded layout
    total = sum((p.numel() + 3) // 4 * 4 for p in ordered)
    assert 0.9 < n_bulk / total < 0.97                                  # "94 % of the bytes"
    rest = [p for p in trainable if id(p) not in wn_ids]
    assert [id(p) for p in ordered[k:]] == [id(p) for p in rest]       # the remainder keeps module order


def _two_region_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    parallel.init_from_env("gloo")
    # the trainer's exchange on a flat buffer (trainer._allreduce / _update): per-flow regions in finishing order, then
    # the remainder, SUM over ranks; the mean and the clip coefficient folded into one scale
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(10_000, generator=g)
    mine = flat.clone()
    regions = [(4_000, 6_400), (1_000, 4_000), (0, 1_000)]
    parallel.allreduce_flat(flat, regions, ready=[True, None, True], chunk_elems=3_000)
    clip = 0.5
    norm = torch.linalg.vector_norm(flat) / world
    scale = (clip / (norm + 1e-6)).clamp(max=1.0) / world
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    mean = sum(gathered) / world
    want = mean * (clip / (torch.linalg.vector_norm(mean) + 1e-6)).clamp(max=1.0)   # clip_grad_norm_ on the mean
    if rank == 0:
        out.put(float((flat * scale - want).abs().max()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_two_region_sum_with_folded_mean_and_clip():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_two_region_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err = out.get(timeout=100)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-6
