#!/usr/bin/env python
"""Headline benchmark: mel frames/s of the RADTTS train step (fwd + bwd + MAS + optimizer) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode train|infer|mas]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Prints ONE JSON line on rank 0 (see the contract in the task statement): whole-job throughput with inputs resident
in HBM (`value`), the same metric end to end from pinned host buffers (`e2e`), the roofline of the dominant kernel
(measured live with CUDA events), and the CPU baseline (the oracle port of the reference's hot path) on a bounded
sample.  `--impl reference` times that CPU oracle alone with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FLOP_PER_FRAME_TRAIN = 635.9e6      # 3 x 211.97 MFLOP per mel frame (fwd + dgrad + wgrad), flow stack only
FLOP_PER_FRAME_FWD = 211.97e6
IN_LAYER_FLOP_PER_GROUP = 2 * 5 * 1024 * 1024   # one dilated k5 1024->1024 conv, per frame group (2 mel frames)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def dist_setup(n):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def barrier_sync(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, world):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def make_model(device):
    from radtts_b200 import configs, synth
    from radtts_b200.radtts import RADTTS
    torch.manual_seed(1234)
    model = RADTTS(**configs.model_config("radtts"))
    synth.load_synth(model, seed=1234)
    return model.to(device)


def pinned_batch(B, T1, T2, seed):
    from radtts_b200 import synth
    b = synth.synth_batch(B, T1, T2, seed=seed)
    return {k: v.pin_memory() for k, v in b.items()}


def to_device(hb, device):
    return {k: v.to(device, non_blocking=True) for k, v in hb.items()}


def time_region(fn, steps, world):
    """CUDA-event timing of `steps` calls bracketed by barrier + synchronize; max over ranks (ms)."""
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier_sync(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


# DRAM traffic of one in_layer launch on this workload from the committed ncu capture (33.45 MB read + 0.24 MB written;
# algorithmic = 22.3 MB activations in + 10.5 MB weights + 22.3 MB out, the output is still in L2 when the launch ends)
IN_LAYER_DRAM_BYTES_NCU = 33.69e6


def in_layer_kernel_probe(model, batch, iters=10):
    """Times the dominant kernel (bf16 tcgen05 row-GEMM of one dilated in_layer conv, K = 5 x 1024) in isolation
    with CUDA events on the launching stream; returns (ms per launch, frame groups per launch)."""
    import ctypes
    from radtts_b200 import _lib, ops
    dev = batch["mel"].device
    g = model.n_group_size
    plan = ops.FramePlan(batch["out_lens"], g, batch["mel"].shape[2] // g)
    flow = model.flows[0]
    dims = ops._flow_dims(flow, 160, 160)
    with torch.no_grad():
        ws = ops._flow_weight_list(flow, False)
        blob = ops.prepare_flow(dims, ws, ops.PREC_BF16, False, dev)
    rows = plan.rows
    x = torch.randn((dims.n_layers + 1, rows, dims.n_ch), device=dev).to(torch.bfloat16)
    L = _lib.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def launch():
        _lib.check(L.radtts_wn_layer_forward(ctypes.byref(dims), _lib.ptr(blob), plan.ptr, plan.B, plan.Tmax,
                                             _lib.ptr(x), 3, ops.PREC_BF16, stream), "radtts_wn_layer_forward")
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        launch()
    e1.record()
    torch.cuda.synchronize()
    groups = int((batch["out_lens"] // g).sum())
    return e0.elapsed_time(e1) / iters, groups


def _time_ms(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def extras(model, batch, peaks):
    """The other two numbers BASELINE.json's metric names: batched decoder inference (frames/s) and MAS ms/batch."""
    from radtts_b200 import alignment, ops
    out = {}
    dev = batch["mel"].device
    B, _, T1 = batch["mel"].shape
    g = model.n_group_size
    frames = int(batch["out_lens"].sum())
    was_training = model.training
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ctx = torch.randn(B, 1040, T1 // g, device=dev) * 0.5
        residual = torch.randn(B, 80 * g, T1 // g, device=dev) * 0.8
        ms = _time_ms(lambda: ops.decoder_inverse(model, residual, ctx, batch["out_lens"]), 5)
        out["infer_decoder"] = {"value": round(frames / ms * 1e3, 1), "unit": "mel frames/s", "ms": round(ms, 3),
                                "what": "8-flow decoder sampling direction (config_ljs_radtts), bf16, batch %d x <=%d frames, "
                                        "conditioning given" % (B, T1),
                                "frac_of_bf16_burst_peak": round(frames * FLOP_PER_FRAME_FWD / ms / 1e9 / peaks["tf_burst"], 4)}
        # full RADTTS.infer (reference radtts.py:541-684): text encoder, length regulation by the given durations,
        # context BiLSTM, 8-flow decoder; durations chosen so that every utterance has exactly out_lens frames
        try:
            T2 = batch["text"].shape[1]
            in_l, out_l = batch["in_lens"].clamp(min=1), batch["out_lens"]
            base = (out_l // in_l)[:, None].expand(-1, T2)
            tok = torch.arange(T2, device=dev)[None, :]
            dur = torch.where(tok < in_l[:, None], base + (tok < (out_l - (out_l // in_l) * in_l)[:, None]).long(),
                              torch.zeros_like(base))
            spk = torch.zeros(B, dtype=torch.long, device=dev)
            ms = _time_ms(lambda: model.infer(spk, batch["text"], 0.8, dur=dur), 3)
            out["infer_e2e"] = {"value": round(frames / ms * 1e3, 1), "unit": "mel frames/s", "ms": round(ms, 3),
                                "what": "RADTTS.infer end to end (text -> mel, durations given), bf16, batch %d x <=%d "
                                        "frames x <=%d tokens; includes a host sync for the output length" % (B, T1, T2)}
        except Exception as e:   # reported, never required
            out["infer_e2e"] = {"error": repr(e)[:200]}
    model.train(was_training)
    for (b, t1, t2) in ((B, T1, batch["text"].shape[1]), (64, 2000, 300)):
        gen = torch.Generator(device=dev).manual_seed(0)
        attn = torch.rand((b, 1, t1, t2), device=dev, generator=gen).add_(1e-6)
        logp = torch.log(attn / attn.sum(3, keepdim=True))
        il = torch.full((b,), t2, dtype=torch.int64, device=dev)
        ol = torch.full((b,), t1, dtype=torch.int64, device=dev)
        ms = _time_ms(lambda: alignment.mas_forward(logp, il, ol, is_prob=False), 10)
        gbs = b * t1 * t2 * 8 / ms / 1e6
        out["mas_%dx%dx%d" % (b, t1, t2)] = {"ms_per_batch": round(ms, 4), "GBps": round(gbs, 1),
                                              "frac_hbm": round(gbs / peaks["hbm_gbs"], 4)}
    return out


def cpu_baseline(B, T1, T2, steps=1, warmup=0):
    """The oracle port of the reference hot path on the host cores (bounded sample)."""
    from oracle import train_step as ots
    from radtts_b200 import configs, synth
    from radtts_b200.radtts import RADTTS
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    m = RADTTS(**configs.model_config("radtts"))
    sd = synth.synth_state_dict([(k, v.shape) for k, v in m.state_dict().items()], seed=1234)
    del m
    batch = synth.synth_batch(B, T1, T2, seed=99)
    fps, sec_per_step, frames = ots.time_steps(sd, batch, steps=steps, warmup=warmup, backward=True, threads=threads)
    return {"value": round(fps, 2), "unit": "mel frames/s", "cores": threads, "kind": "port",
            "ms_per_step_sample": round(sec_per_step * 1e3, 1),
            "sample": "oracle hot path (ConvAttention + serial MAS + 8 decoder flows fwd+bwd, fp32) on B=%d x %d frames x "
                      "%d tokens, %d step(s), %.1f s/step" % (B, T1, T2, steps, sec_per_step)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" of this arm = the oracle's train step on a bounded sample (4 of the 32 utterances of a full batch):
    # ~2 s of CPU work per step on 16 cores, so a --steps 8 --warmup 3 run still ends within a minute
    cb = cpu_baseline(4, args.t1, args.t2, steps=max(1, args.steps), warmup=max(0, min(1, args.warmup)))
    line = {"impl": "reference", "metric": "mel frames/s, RADTTS decoder train step (fwd+bwd+MAS)", "value": cb["value"],
            "unit": "mel frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step_sample"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": workload_config(args), "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": "mel frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line["config"]["global_batch"] = args.batch * max(1, args.gpus)   # the arm's config is the GPU arm's, key for key
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "config_ljs_radtts full decoder train step (fwd+bwd+MAS+RAdam), batch %d/GPU x <=%d mel frames "
                        "x <=%d tokens, 80 mel bins" % (args.batch, args.t1, args.t2),
            "global_batch": None, "per_gpu_batch": args.batch, "max_frames": args.t1, "max_tokens": args.t2,
            "parallelism": "dp", "l2": "working set (>1.5 GB of activations per step) exceeds the 126 MB L2"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--t1", type=int, default=800)
    ap.add_argument("--t2", type=int, default=150)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="run the step eagerly instead of as one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: radtts_b200 has no CPU fallback")
    world, rank, local = dist_setup(args.gpus)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    from radtts_b200 import _lib, configs
    from radtts_b200.trainer import TrainStep
    peaks = load_peaks()

    model = make_model(device).train()
    use_graph = not args.no_graph   # N > 1: two graphs with the NCCL all-reduce launched eagerly in between
    ts = TrainStep(model, configs.LOSS_WEIGHTS, bf16=True, ddp=world > 1, device_ids=[local] if world > 1 else None,
                   capturable=use_graph)
    host_batches = [pinned_batch(args.batch, args.t1, args.t2, seed=1000 + 17 * rank + i) for i in range(2)]
    dev_batches = [to_device(b, device) for b in host_batches]
    graph_note = "eager"
    if use_graph:
        try:
            ts.capture(dev_batches[0])
            graph_note = "cuda_graph"
        except Exception as e:  # fall back to the eager step, say so in the JSON line
            graph_note = "eager (graph capture failed: %s)" % str(e).splitlines()[0][:120]
            torch.cuda.synchronize()
            ts.graph = None
    frames_per_step_local = float(sum(int(b["out_lens"].sum()) for b in host_batches)) / len(host_batches)
    h2d = sum(v.numel() * v.element_size() for v in host_batches[0].values())

    it = {"i": 0}

    def step_resident():
        ts.step(dev_batches[it["i"] % len(dev_batches)])
        it["i"] += 1

    losses = []

    def step_e2e():
        hb = host_batches[it["i"] % len(host_batches)]
        b = hb if ts.graph is not None else to_device(hb, device)   # graph mode copies host -> static device buffers
        loss = ts.step(b)
        losses.append(float(loss.item()))   # device -> host read of the step's result
        it["i"] += 1

    for _ in range(max(3, args.warmup)):
        step_resident()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count()
    torch.cuda.profiler.start()   # no-op unless run under `ncu --profile-from-start off` (profiles/ recipe)
    ms = time_region(step_resident, args.steps, world)
    torch.cuda.profiler.stop()
    launches = _lib.launch_count() - launches0
    if ts.graph is not None:
        launches = ts.launches_per_replay * args.steps   # replays re-issue the launches recorded at capture time
    clocks = sampler.stop() if sampler else None
    frames_total = sum_over_ranks(frames_per_step_local, world)
    value = frames_total * args.steps / (ms / 1e3)

    step_e2e()
    ms_e2e = time_region(step_e2e, args.steps, world)
    e2e_value = frames_total * args.steps / (ms_e2e / 1e3)

    roofline = None
    cb = None
    if rank == 0:
        k_ms, groups = in_layer_kernel_probe(model, dev_batches[0])
        achieved = groups * IN_LAYER_FLOP_PER_GROUP / (k_ms / 1e3) / 1e12
        roofline = {"bound": "tensor", "kernel": "rowgemm_tc_kernel<EpiBiasAct<bf16>> (WN in_layer, dilated k5 1024->1024)",
                    "achieved": round(achieved, 1), "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                    "frac": round(achieved / peaks["tf_burst"], 4), "traffic": IN_LAYER_DRAM_BYTES_NCU,
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, "
                                      "profiles/r01b_rowgemm_tc_ncu_raw.csv (same workload)",
                    "peak_source": peaks["source"] + " burst",
                    "ms_per_launch": round(k_ms, 4),
                    "step_frac_of_sustained_peak": round(value / world * FLOP_PER_FRAME_TRAIN / 1e12 / peaks["tf_sustained"], 4)}
    extra = None
    if rank == 0 and world == 1:
        try:
            extra = extras(model, dev_batches[0], peaks)
        except Exception as e:
            extra = {"error": repr(e)}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_baseline(4, args.t1, args.t2, steps=3, warmup=1)
        except Exception as e:  # the baseline is reported, never required
            cb = {"value": None, "error": repr(e)}
    if rank == 0:
        line = {"metric": "mel frames/s, RADTTS decoder train step (fwd+bwd+MAS)", "value": round(value, 1),
                "unit": "mel frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": round(e2e_value, 1), "unit": "mel frames/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3)},
                "gpu_launches": int(launches * world), "clocks": clocks, "roofline": roofline, "cpu_baseline": cb,
                "loss_last": losses[-1] if losses else None, "extra": extra}
        line["config"]["global_batch"] = args.batch * world
        line["config"]["execution"] = graph_note
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
