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


def make_model(device, name="radtts"):
    from radtts_b200 import configs, synth
    from radtts_b200.radtts import RADTTS
    # CPU generator only: every value is overwritten by load_synth anyway, and re-seeding the CUDA generator after a
    # graph capture that registered its state raises "Offset increment outside graph capture"
    torch.default_generator.manual_seed(1234)
    model = RADTTS(**configs.model_config(name))
    synth.load_synth(model, seed=1234)
    return model.to(device)


def pinned_batch(B, T1, T2, seed, with_attributes=False):
    from radtts_b200 import synth
    b = synth.synth_batch(B, T1, T2, seed=seed, with_attributes=with_attributes)
    return {k: v.pin_memory() for k, v in b.items()}


def to_device(hb, device):
    return {k: v.to(device, non_blocking=True) for k, v in hb.items()}


def time_region(fn, steps, world, tail=None):
    """CUDA-event timing of `steps` calls bracketed by barrier + synchronize; max over ranks (ms).  tail: called once
    after the last step, INSIDE the timed region (flushes a pipelined optimizer update: all the work of all `steps`
    steps is done when the clock stops)."""
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    if tail is not None:
        tail()
    e1.record()
    barrier_sync(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


def in_layer_traffic_from_profiles():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the in_layer GEMM, read from the newest committed
    `ncu --set full --page raw --csv` capture under profiles/ (tools/ncu_traffic.py writes the small JSON next to it).
    (bytes or None, description of where the number comes from)."""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_in_layer_traffic.json")))
    if not cands:
        return None, "no profiles/r*_in_layer_traffic.json in this tree"
    d = json.load(open(cands[-1]))
    return d["dram_bytes_per_launch"], "%s (%s)" % (os.path.relpath(cands[-1], ROOT), d.get("source", "ncu --set full"))


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


def durations_for(batch):
    """Token durations that give every utterance exactly out_lens frames (decoder-only synthesis, durations given)."""
    dev = batch["mel"].device
    T2 = batch["text"].shape[1]
    in_l, out_l = batch["in_lens"].clamp(min=1), batch["out_lens"]
    base = (out_l // in_l)[:, None].expand(-1, T2)
    tok = torch.arange(T2, device=dev)[None, :]
    return torch.where(tok < in_l[:, None], base + (tok < (out_l - (out_l // in_l) * in_l)[:, None]).long(),
                       torch.zeros_like(base))


def infer_legs(model, batch, peaks, world):
    """Batched synthesis of config_ljs_radtts (BASELINE metric "infer"): the 8-flow decoder alone, and RADTTS.infer end to
    end.  Utterances are independent: every rank runs its own batch (replicas only), the job's frames/s is the sum of the
    ranks' frames over the slowest rank's time."""
    from radtts_b200 import ops
    out = {}
    dev = batch["mel"].device
    B, _, T1 = batch["mel"].shape
    g = model.n_group_size
    frames = sum_over_ranks(float(batch["out_lens"].sum()), world)
    was_training = model.training
    model.eval()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ctx = torch.randn(B, 1040, T1 // g, device=dev) * 0.5
        residual = torch.randn(B, 80 * g, T1 // g, device=dev) * 0.8
        ms = max_over_ranks(_time_ms(lambda: ops.decoder_inverse(model, residual, ctx, batch["out_lens"]), 5), world)
        out["infer_decoder"] = {"value": round(frames / ms * 1e3, 1), "unit": "mel frames/s", "ms": round(ms, 3),
                                "what": "8-flow decoder sampling direction (config_ljs_radtts), bf16, batch %d/GPU x <=%d "
                                        "frames, conditioning given" % (B, T1),
                                "frac_of_bf16_burst_peak": round(frames / world * FLOP_PER_FRAME_FWD / ms / 1e9
                                                                 / peaks["tf_burst"], 4)}
        try:
            dur = durations_for(batch)
            spk = torch.zeros(B, dtype=torch.long, device=dev)
            ms = max_over_ranks(_time_ms(lambda: model.infer(spk, batch["text"], 0.8, dur=dur), 3), world)
            out["infer_e2e"] = {"value": round(frames / ms * 1e3, 1), "unit": "mel frames/s", "ms": round(ms, 3),
                                "what": "RADTTS.infer end to end (text -> mel, durations given), bf16, batch %d/GPU x <=%d "
                                        "frames x <=%d tokens; includes a host sync for the output length"
                                        % (B, T1, batch["text"].shape[1])}
        except Exception as e:   # reported, never required
            out["infer_e2e"] = {"error": repr(e)[:200]}
    model.train(was_training)
    return out


def cfg4_leg(device, world, rank, peaks):
    """BASELINE configs[3]: config_ljs_bgap batched sampling, sigma = 0.8, 32 utterances/GPU x 100 tokens x ~500 frames:
    durations given, voicing predicted (batched DAP), F0 and energy SAMPLED by the bi-partite RQ-spline flows, then the
    8-flow decoder."""
    from radtts_b200 import ops
    model = make_model(device, "bgap").eval()
    B, T2 = 32, 100
    gen = torch.Generator().manual_seed(4000 + rank)
    text = torch.randint(1, 185, (B, T2), generator=gen).to(device)
    dur = torch.randint(2, 9, (B, T2), generator=gen)
    dur[:, 0] += (4 - dur.sum(1) % 4) % 4                  # total frames a multiple of 4
    dur = dur.to(device)
    spk = torch.zeros(B, dtype=torch.long, device=device)
    in_lens = torch.full((B,), T2, dtype=torch.int64, device=device)
    frames = sum_over_ranks(float(dur.sum()), world)
    res = {}
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        fn = lambda: model.infer(spk, text, 0.8, sigma_f0=0.8, sigma_energy=0.8, dur=dur, in_lens=in_lens)  # noqa: E731
        ms = max_over_ranks(_time_ms(fn, 3), world)
        res["infer_cfg4_bgap"] = {
            "value": round(frames / ms * 1e3, 1), "unit": "mel frames/s", "ms": round(ms, 3),
            "what": "config_ljs_bgap RADTTS.infer, sigma=0.8, bf16, batch %d/GPU x %d tokens x ~%d frames: voicing (DAP), "
                    "F0 + energy (BGAP RQ-spline flows) predicted, 8-flow decoder" % (B, T2, int(dur.sum(1).float().mean()))}
        # the attribute flows alone
        T = int(dur.sum(1).max())
        lens = dur.sum(1)
        txt = torch.randn(B, 512, T, device=device) * 0.5
        spk_vec = model.encode_speaker(spk)
        z = torch.randn(B, 2, T, device=device) * 0.8
        ms_f0 = max_over_ranks(_time_ms(lambda: model.f0_pred_module.infer(z, txt, spk_vec, lens), 3), world)
        ms_en = max_over_ranks(_time_ms(lambda: model.energy_pred_module.infer(z, txt, spk_vec, lens), 3), world)
        res["bgap_flows"] = {"f0_ms": round(ms_f0, 3), "energy_ms": round(ms_en, 3),
                             "what": "BGAP.infer alone (6 flows: 2 affine + 4 RQ-spline), same batch"}
    del model
    ops.POOL.clear()
    torch.cuda.empty_cache()
    return res


MAS_GRID_T2 = (50, 100, 150, 200, 300)
MAS_GRID_T1 = (200, 400, 800, 1200, 2000)


def mas_attention_sweep(device, world, rank, peaks, cpu_time=True):
    """BASELINE configs[4]: standalone MAS / ConvAttention sweep, batch 64, text 50-300 x mel 200-2000 (T1 >= T2), the 64
    utterances sharded over the ranks (replicas only).  Per point: MAS ms/batch and GB/s at 8 B per lattice cell,
    ConvAttention core forward ms and GB/s at 12 B per cell + inputs, both against the measured HBM peak -- and, on rank 0
    at N=1, the reference's serial per-utterance loop (C restatement of alignment.py:31-59, one host thread)."""
    import numpy as np
    from radtts_b200 import alignment, ops, parallel
    rows = []
    lo, hi = parallel.shard_range(64, rank, world)
    b = hi - lo
    for t2 in MAS_GRID_T2:
        for t1 in MAS_GRID_T1:
            if t1 < t2:
                continue
            gen = torch.Generator(device=device).manual_seed(17 * t1 + t2)
            attn = torch.rand((64, 1, t1, t2), device=device, generator=gen).add_(1e-6)
            attn = (attn / attn.sum(3, keepdim=True))[lo:hi].contiguous()
            logp = torch.log(attn)
            il = torch.full((b,), t2, dtype=torch.int64, device=device)
            ol = torch.full((b,), t1, dtype=torch.int64, device=device)
            ms_lp = max_over_ranks(_time_ms(lambda: alignment.mas_forward(logp, il, ol, is_prob=False), 10), world)
            ms_p = max_over_ranks(_time_ms(lambda: alignment.mas_forward(attn, il, ol, is_prob=True), 10), world)
            cells = 64 * t1 * t2
            row = {"T1": t1, "T2": t2, "mas_ms": round(ms_lp, 4), "mas_prob_in_ms": round(ms_p, 4),
                   "mas_GBps": round(cells * 8 / ms_lp / 1e6, 1),
                   "mas_frac_hbm": round(cells * 8 / ms_lp / 1e6 / peaks["hbm_gbs"] / world, 4)}
            q = torch.randn((b, 80, t1), device=device, generator=gen)
            k = torch.randn((b, 80, t2), device=device, generator=gen)
            prior = torch.rand((b, t1, t2), device=device, generator=gen)
            with torch.no_grad():
                ms_a = max_over_ranks(_time_ms(lambda: ops._ConvAttnFn.apply(q, k, prior, il, 0.0005), 10), world)
            abytes = cells * 12 + 64 * 80 * (t1 + t2) * 4
            row.update(attn_fwd_ms=round(ms_a, 4), attn_fwd_GBps=round(abytes / ms_a / 1e6, 1),
                       attn_fwd_frac_hbm=round(abytes / ms_a / 1e6 / peaks["hbm_gbs"] / world, 4))
            if cpu_time and world == 1 and rank == 0:
                from oracle import mas as omas
                a_np = attn.cpu().numpy()
                iln, oln = np.full(64, t2, dtype=np.int64), np.full(64, t1, dtype=np.int64)
                t0 = time.perf_counter()
                omas.binarize(a_np, iln, oln, is_prob=True)
                row["cpu_serial_ms"] = round((time.perf_counter() - t0) * 1e3, 2)
            rows.append(row)
            del attn, logp, q, k, prior
    return {"n_gpus": world, "batch": 64, "utterances_per_gpu": b, "cpu": "C restatement of the reference's serial loop "
            "(radtts.py:326-334 + alignment.py:31-59), 1 host thread, probabilities in (logf inside)", "points": rows}


def gpu_torch_baseline(B, T1, T2, steps=3):
    """SURVEY 2.3's bar: the reference's hot path as STOCK PyTorch on this B200 (cuDNN / cuBLAS under bf16 autocast, the
    MAS round trip through the host exactly as radtts.py:326-334 does it) -- oracle/train_step.py with its tensors on
    cuda.  Same batch shape as the headline step; like the CPU baseline it leaves out the text encoder, the LSTMs, CTC
    and the optimizer, so it UNDERSTATES the reference's step time."""
    from oracle import train_step as ots
    from radtts_b200 import configs, synth
    from radtts_b200.radtts import RADTTS
    dev = torch.device("cuda", torch.cuda.current_device())
    m = RADTTS(**configs.model_config("radtts"))
    sd = synth.synth_state_dict([(k, v.shape) for k, v in m.state_dict().items()], seed=1234)
    del m
    sd = {k: v.to(dev) for k, v in sd.items()}
    batch = {k: v.to(dev) for k, v in synth.synth_batch(B, T1, T2, seed=99).items()}
    sd = {k: (v.clone().requires_grad_(True) if (k.startswith(("flows.", "attention.")) and v.dtype.is_floating_point
                                                 and not k.endswith((".p", "lower_diag"))) else v) for k, v in sd.items()}
    keys, text_enc, spk = ots.make_inputs(sd, batch)
    frames = int(batch["out_lens"].sum())

    def step():
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ots.hot_path_step(sd, batch, keys, text_enc, spk, True)
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    return {"value": round(frames / sec, 1), "unit": "mel frames/s", "ms_per_step": round(sec * 1e3, 2), "kind": "port",
            "what": "oracle hot path (ConvAttention + host-side serial MAS + 8 decoder flows fwd+bwd) as stock PyTorch ops on "
                    "cuda:0 under bf16 autocast, batch %d x <=%d frames x <=%d tokens; no text encoder / LSTM / CTC / "
                    "optimizer (understates the reference)" % (B, T1, T2)}


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
            "sample": "oracle hot path (ConvAttention + serial MAS + 8 decoder flows fwd+bwd, fp32; no text encoder / LSTM / "
                      "CTC / optimizer) on B=%d of the %d utterances x <=%d frames x <=%d tokens, %d step(s), %.1f s/step"
                      % (B, 32, T1, T2, steps, sec_per_step)}


REF_SAMPLE_B = 4
_KEEPALIVE = []


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" of this arm = the oracle's train step on a bounded sample (4 of the 32 utterances of a full batch):
    # ~2 s of CPU work per step on 16 cores, so a --steps 8 --warmup 3 run still ends within a minute
    cb = cpu_baseline(REF_SAMPLE_B, args.t1, args.t2, steps=max(1, args.steps), warmup=max(0, min(1, args.warmup)))
    cfg = workload_config(args)
    cfg["global_batch"] = args.batch * max(1, args.gpus)   # the arm's config is the GPU arm's, key for key
    cfg["execution"] = "cpu"
    cfg["reference_sample"] = ("each step ran B=%d of the %d utterances of that batch (frames/s is per frame, so the sample "
                               "size does not bias it), hot path only: no text encoder, LSTMs, CTC or optimizer -- less "
                               "work per frame than the GPU arm does" % (REF_SAMPLE_B, args.batch))
    line = {"impl": "reference", "metric": "mel frames/s, RADTTS decoder train step (fwd+bwd+MAS)", "value": cb["value"],
            "unit": "mel frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": cb["ms_per_step_sample"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": cfg, "cpu_baseline": cb, "gpu_launches": 0,
            "same_config": False,
            "e2e": {"value": cb["value"], "unit": "mel frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "config_ljs_radtts full decoder train step (fwd+bwd+MAS+RAdam), batch %d/GPU x <=%d mel frames "
                        "x <=%d tokens, 80 mel bins" % (args.batch, args.t1, args.t2),
            "global_batch": None, "per_gpu_batch": args.batch, "max_frames": args.t1, "max_tokens": args.t2,
            "parallelism": "dp", "l2": "working set (>1.5 GB of activations per step) exceeds the 126 MB L2"}


def train_leg(name, args, world, rank, local, device, steps, with_attributes=False, use_graph=True, sampler=None):
    """Times `steps` optimisation steps of config_ljs_<name> (forward, losses, backward, gradient exchange, clip, RAdam):
    resident inputs (value) and end to end from pinned host buffers with a device->host read of the loss (e2e)."""
    from radtts_b200 import _lib, configs, ops
    from radtts_b200.trainer import TrainStep
    cfg = configs.model_config(name)
    model = make_model(device, name).train()
    host_batches = [pinned_batch(args.batch, args.t1, args.t2, seed=1000 + 17 * rank + i, with_attributes=with_attributes)
                    for i in range(2)]
    dev_batches = [to_device(b, device) for b in host_batches]
    lk = dict(dur_model_config=cfg["dur_model_config"], f0_model_config=cfg["f0_model_config"],
              energy_model_config=cfg["energy_model_config"], vpred_model_config=cfg["v_model_config"]) \
        if name != "radtts" else None
    ts = TrainStep(model, configs.LOSS_WEIGHTS, bf16=True, ddp=world > 1, device_ids=[local] if world > 1 else None,
                   capturable=use_graph, loss_kwargs=lk, probe_batch=dev_batches[0] if name != "radtts" else None,
                   deferred_update=args.deferred_update)
    graph_note = "eager"
    if use_graph:
        try:
            ts.capture(dev_batches[0])
            graph_note = "cuda_graph"
        except Exception as e:  # fall back to the eager step, say so in the JSON line
            graph_note = "eager (graph capture failed: %s)" % str(e).splitlines()[0][:120]
            print("graph capture failed for %s: %r" % (name, e), file=sys.stderr)
            torch.cuda.synchronize()
            ts.graph = None
    frames_local = float(sum(int(b["out_lens"].sum()) for b in host_batches)) / len(host_batches)
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
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count()
    torch.cuda.profiler.start()   # no-op unless run under `ncu --profile-from-start off` (profiles/ recipe)
    ms = time_region(step_resident, steps, world, tail=ts.flush)
    torch.cuda.profiler.stop()
    launches = _lib.launch_count() - launches0
    if ts.graph is not None:
        launches = ts.launches_per_replay * steps   # replays re-issue the launches recorded at capture time
    clocks = sampler.stop() if sampler else None
    frames_total = sum_over_ranks(frames_local, world)
    step_e2e()
    ms_e2e = time_region(step_e2e, steps, world, tail=ts.flush)
    res = {"value": frames_total * steps / (ms / 1e3), "ms_per_step": ms / steps,
           "e2e_value": frames_total * steps / (ms_e2e / 1e3), "e2e_ms_per_step": ms_e2e / steps, "h2d": h2d,
           "launches": int(launches * world), "clocks": clocks, "loss_last": losses[-1] if losses else None,
           "execution": graph_note, "frames_per_step": frames_total}
    return res, model, ts, dev_batches


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
    ap.add_argument("--deferred-update", action="store_true",
                    help="TrainStep(deferred_update=True): the flow-parameter region of each RAdam update is applied at the top "
                         "of the next step, underneath its text encoder / attention / context LSTM, and flushed inside the "
                         "timed region (default: the whole update at the end of each step; measured equal within noise)")
    ap.add_argument("--no-extras", action="store_true", help="headline train step only (no cfg3 / cfg4 / infer / sweep legs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: radtts_b200 has no CPU fallback")
    world, rank, local = dist_setup(args.gpus)
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    from radtts_b200 import ops
    peaks = load_peaks()
    use_graph = not args.no_graph   # N > 1: two graphs with the NCCL all-reduces launched eagerly in between

    sampler = ClockSampler(local) if rank == 0 else None
    head, model, ts, dev_batches = train_leg("radtts", args, world, rank, local, device, args.steps, use_graph=use_graph,
                                             sampler=sampler)
    value = head["value"]

    roofline = None
    if rank == 0:
        k_ms, groups = in_layer_kernel_probe(model, dev_batches[0])
        achieved = groups * IN_LAYER_FLOP_PER_GROUP / (k_ms / 1e3) / 1e12
        traffic, traffic_src = in_layer_traffic_from_profiles()
        roofline = {"bound": "tensor", "kernel": "rowgemm_tc_kernel<EpiBiasAct<bf16>> (WN in_layer, dilated k5 1024->1024)",
                    "achieved": round(achieved, 1), "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                    "frac": round(achieved / peaks["tf_burst"], 4), "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peaks["source"] + " burst (kernel timed alone)",
                    "ms_per_launch": round(k_ms, 4),
                    "step_frac_of_sustained_peak": round(value / world * FLOP_PER_FRAME_TRAIN / 1e12 / peaks["tf_sustained"], 4)}
    extra = {}
    if not args.no_extras:
        def leg(name, fn):
            try:
                extra.update(fn())
            except Exception as e:   # the extra legs are reported, never required
                extra[name] = {"error": repr(e)[:300]}
            if world > 1:
                import torch.distributed as dist
                dist.barrier()

        leg("infer", lambda: infer_legs(model, dev_batches[0], peaks, world))
        # the captured step stays alive until the process ends: destroying a CUDA graph that registered the default
        # generator's state left torch 2.11's generator refusing every later eager random op ("Offset increment outside
        # graph capture"); 180 GB of HBM make keeping ~10 GB around a non-issue
        _KEEPALIVE.append((ts, model))

        def cfg3():
            a3 = argparse.Namespace(**vars(args))
            r, m3, t3, _ = train_leg("decoder", a3, world, rank, local, device, max(2, min(4, args.steps)),
                                     with_attributes=True, use_graph=use_graph)
            _KEEPALIVE.append((t3, m3))
            return {"train_cfg3_decoder": {
                "value": round(r["value"], 1), "unit": "mel frames/s", "ms_per_step": round(r["ms_per_step"], 3),
                "e2e_value": round(r["e2e_value"], 1), "execution": r["execution"], "loss_last": r["loss_last"],
                "what": "config_ljs_decoder (F0 / energy / voicing conditioned decoder, RADTTS++) full train step, bf16, "
                        "batch %d/GPU x <=%d frames x <=%d tokens, DP over %d GPU(s)" % (args.batch, args.t1, args.t2, world)}}
        leg("train_cfg3_decoder", cfg3)
        leg("infer_cfg4_bgap", lambda: cfg4_leg(device, world, rank, peaks))
        leg("mas_attention_sweep", lambda: {"mas_attention_sweep": mas_attention_sweep(device, world, rank, peaks)})
        if rank == 0 and world == 1:
            leg("gpu_torch_baseline", lambda: {"gpu_torch_baseline": gpu_torch_baseline(args.batch, args.t1, args.t2)})
    cb = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_baseline(REF_SAMPLE_B, args.t1, args.t2, steps=3, warmup=1)
        except Exception as e:  # the baseline is reported, never required
            cb = {"value": None, "error": repr(e)}
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    if rank == 0:
        line = {"metric": "mel frames/s, RADTTS decoder train step (fwd+bwd+MAS)", "value": round(value, 1),
                "unit": "mel frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
                "ms_per_step": round(head["ms_per_step"], 3), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
                "e2e": {"value": round(head["e2e_value"], 1), "unit": "mel frames/s", "h2d_bytes_per_step": head["h2d"],
                        "d2h_bytes_per_step": 4, "ms_per_step": round(head["e2e_ms_per_step"], 3)},
                "gpu_launches": head["launches"], "clocks": head["clocks"], "roofline": roofline, "cpu_baseline": cb,
                "loss_last": head["loss_last"], "extra": extra or None}
        line["config"]["global_batch"] = args.batch * world
        line["config"]["execution"] = head["execution"]
        line["config"]["optimizer_update"] = ("whole update at the end of the step" if not args.deferred_update else
                                              "flow-parameter region applied at the top of the next step (same kernels, "
                                              "pipelined); the last one is flushed inside the timed region")
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
