"""Times the decoder flow stack (8 flows) forward / inverse at cfg2 shapes with CUDA events."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radtts_b200 import configs, ops, synth
from radtts_b200.radtts import RADTTS

FLOP_PER_FRAME_FWD = 211.97e6


def main(prec="bf16", B=32, T1=800, iters=5):
    torch.manual_seed(0)
    model = RADTTS(**configs.model_config("radtts")).eval()
    synth.load_synth(model, seed=1234)
    model = model.cuda()
    batch = synth.synth_batch(B, T1, 150, seed=7)
    mel = batch["mel"].cuda()
    out_lens = batch["out_lens"].cuda()
    ctx = torch.randn(B, 1040, T1 // 2, device="cuda") * 0.5
    frames = int(batch["out_lens"].sum())
    ops.set_precision(prec)
    res = {}
    with torch.no_grad():
        for name, fn in (("forward", lambda: ops.decoder_forward(model, mel, ctx, out_lens)),
                         ("inverse", lambda: ops.decoder_inverse(model, ops.squeeze_time(mel, 2), ctx, out_lens))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(iters):
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ts.sort()
            ms = ts[len(ts) // 2]
            res[name] = {"ms": round(ms, 3), "frames_per_s": round(frames / ms * 1e3),
                         "tflops": round(frames * FLOP_PER_FRAME_FWD / ms / 1e9, 1)}
    print(json.dumps({"prec": prec, "B": B, "T1": T1, "valid_frames": frames, **res}))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "bf16")
