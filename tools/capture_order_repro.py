"""Bisects the one known TrainStep.capture failure: an eager step on the legacy default stream followed
by capture() raised cudaErrorStreamCaptureImplicit from the autograd engine.  Each scenario runs in its own process:

    python tools/capture_order_repro.py            # runs all scenarios, one line each
    python tools/capture_order_repro.py <name>     # runs one scenario in this process

Scenarios: what happens BEFORE capture()
    none            nothing (the benchmarked order)                          -- expected ok
    legacy          one eager _fwd_bwd on the legacy stream                  -- failed in round 1
    legacy_gc       the same, then gc.collect()
    side            one eager _fwd_bwd on a side stream
    legacy_nohook   legacy, with the CTC prefetch hook removed
    legacy_nodirect legacy, with direct gradient accumulation switched off
    legacy_fwd_only a grad-mode forward without backward on the legacy stream
"""
import gc
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SCENARIOS = ["none", "legacy", "legacy_gc", "side", "legacy_nohook", "legacy_nodirect", "legacy_fwd_only"]


def run(name):
    import torch
    from radtts_b200 import configs, ops, synth
    from radtts_b200.radtts import RADTTS
    from radtts_b200.trainer import TrainStep
    torch.manual_seed(0)
    m = RADTTS(**configs.model_config("radtts"))
    synth.load_synth(m, seed=1234)
    m = m.cuda().train()
    batch = {k: v.cuda() for k, v in synth.synth_batch(4, 96, 24, seed=4242).items()}
    ts = TrainStep(m, configs.LOSS_WEIGHTS, bf16=True, capturable=True)
    if name == "legacy_nohook":
        ts._ctc_hook.remove()
    if name == "legacy_nodirect":
        ops.set_direct_grad_accumulation(False)
    if name.startswith("legacy"):
        if name == "legacy_fwd_only":
            ts.forward_loss(batch)
        else:
            ts._fwd_bwd(batch)
        if name == "legacy_gc":
            gc.collect()
    elif name == "side":
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            ts._fwd_bwd(batch)
        torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    ts.capture(batch)
    l1 = float(ts.step(batch))
    l2 = float(ts.step(batch))
    print("ok  losses %.4f %.4f" % (l1, l2))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for name in SCENARIOS:
            r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=300)
            last = (r.stdout.strip().splitlines() or [""])[-1]
            err = [ln for ln in r.stderr.splitlines() if "Error" in ln]
            print("%-16s rc=%d %s %s" % (name, r.returncode, last, err[-1][:160] if err and r.returncode else ""), flush=True)
