"""Worker of tests/test_dp_gpu.py (launched by torch.distributed.run, one process per GPU): the data-parallel TrainStep
(two CUDA graphs + per-flow NCCL all-reduces on the comm stream) against a single-process computation of the same
update -- the mean over ranks of the per-rank gradients (reference distributed.py:133-140), clipped, one RAdam step."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from radtts_b200 import configs, parallel, synth  # noqa: E402
from radtts_b200.radtts import RADTTS  # noqa: E402
from radtts_b200.trainer import TrainStep  # noqa: E402


def model_on(dev):
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("radtts")).eval()   # no dropout, no spectral-norm power iteration: deterministic
    synth.load_synth(m, seed=1234)
    return m.to(dev)


def main():
    world, rank, local = parallel.init_from_env("nccl")
    dev = torch.device("cuda", local)
    B, T1, T2 = 4, 200, 40
    batches = [{k: v.to(dev) for k, v in synth.synth_batch(B, T1, T2, seed=500 + r).items()} for r in range(world)]
    use_graph = os.environ.get("DP_TEST_GRAPH", "1") == "1"
    m = model_on(dev)
    ts = TrainStep(m, configs.LOSS_WEIGHTS, bf16=True, ddp=True, capturable=True)
    if use_graph:
        # capture() warms up with 3 real steps; undo them so that the compared step starts from the initial weights
        ts.capture(batches[rank])
        m2 = model_on(dev)
        with torch.no_grad():
            for p, q in zip(list(m.parameters()) + list(m.buffers()), list(m2.parameters()) + list(m2.buffers())):
                p.copy_(q)
        for t in (ts.optimizer.exp_avg, ts.optimizer.exp_avg_sq):
            t.zero_()
        ts.optimizer.step_dev.zero_()
        del m2
    loss = ts.step(batches[rank])
    torch.cuda.synchronize()
    got_flat = ts.optimizer.flat.clone()
    # the update zeroes the gradients in the pass that consumes them; after ONE step from zero moments the first moment is
    # (1 - beta1) * clipped mean gradient: the gradient check is made on it
    got_grad = ts.optimizer.exp_avg.clone()
    order = [id(p) for p in ts.optimizer.params]

    # single-process ground truth on THIS rank: per-rank gradients of every rank's batch, averaged
    ref = model_on(dev)
    rs = TrainStep(ref, configs.LOSS_WEIGHTS, bf16=True, ddp=False, capturable=True)
    ref_by_name = dict(ref.named_parameters())
    names = {id(p): n for n, p in m.named_parameters()}
    acc = None
    for r in range(world):
        rs._fwd_bwd(batches[r])
        acc = rs.optimizer.grad.clone() if acc is None else acc + rs.optimizer.grad
    rs.optimizer.grad.copy_(acc / world)
    rs._update()
    torch.cuda.synchronize()
    # compare parameter by parameter (the two flat buffers are ordered differently: flow networks first under DP)
    off = 0
    ref_off = {}
    for p in rs.optimizer.params:
        ref_off[id(p)] = off
        off += (p.numel() + 3) // 4 * 4
    ref_name_off = {n: ref_off[id(p)] for n, p in ref_by_name.items() if id(p) in ref_off}
    worst_g, worst_p, off = 0.0, 0.0, 0
    num_g = den_g = num_p = den_p = 0.0
    for p in ts.optimizer.params:
        n = names[id(p)]
        ro = ref_name_off[n]
        k = p.numel()
        g_dp = got_grad[off:off + k]
        g_rf = rs.optimizer.exp_avg[ro:ro + k]
        num_g += float((g_dp - g_rf).double().pow(2).sum()); den_g += float(g_rf.double().pow(2).sum())
        w_dp, w_rf = got_flat[off:off + k], rs.optimizer.flat[ro:ro + k]
        num_p += float((w_dp - w_rf).double().pow(2).sum()); den_p += float(w_rf.double().pow(2).sum())
        off += (k + 3) // 4 * 4
    res = {"rank": rank, "world": world, "graph": use_graph, "loss": float(loss),
           "grad_rel_err": (num_g / den_g) ** 0.5, "param_rel_err": (num_p / den_p) ** 0.5,
           "flow_final": list(map(bool, ts._flow_final)), "n_regions": len(ts.flow_regions)}
    # replicas must stay identical after the step
    chk = got_flat.double().sum().reshape(1)
    lst = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(lst, chk)
    res["replica_checksums_equal"] = bool(all(float(x) == float(lst[0]) for x in lst))
    if rank == 0:
        print("DPRESULT " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
