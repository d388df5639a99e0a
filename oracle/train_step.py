"""CPU restatement of the hot path of one RADTTS training step -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

What is timed here is the path the B200 kernels replace, run the way the reference runs it on a CPU:
ConvAttention (common.py:886-924) -> MAS binarisation, serial over the batch on one thread (radtts.py:320-334,
alignment.py:31-59) -> hard-attention context (radtts.py:399) -> 8 decoder flows (radtts.py:431-444) -> flow loss
(loss.py:27-52) -> backward through all of it with PyTorch autograd (what train.py:416 does).
The text encoder / context LSTM / CTC loss are outside the hot-path scope and are not part of this baseline;
the decoder conditioning is the squeezed hard-attention context plus the speaker vector (the reference's
n_flowstep_cond_dims = 16 + 512 * 2 = 1040 channels, radtts.py:121-123) without the LSTM in between.
"""
import time

import numpy as np
import torch

from . import flow as oflow
from . import mas as omas


def make_inputs(sd, batch, seed=0):
    rng = np.random.default_rng(seed)
    B, _, T1 = batch["mel"].shape
    T2 = batch["text"].shape[1]
    keys = sd["embedding.weight"][batch["text"]].transpose(1, 2).contiguous()            # (B, 512, T2)
    text_enc = torch.from_numpy(rng.standard_normal((B, 512, T2), dtype=np.float32) * 0.3).to(batch["mel"].device)
    spk = sd["speaker_embedding.weight"][batch["speaker_ids"]]                          # (B, 16)
    return keys, text_enc, spk


def hot_path_step(sd, batch, keys, text_enc, spk, backward=True):
    """Returns (loss, valid_frames).  `sd` must hold leaf tensors with requires_grad for flows.* / attention.*."""
    dev = batch["mel"].device
    key_mask = ~(torch.arange(batch["text"].shape[1], device=dev)[None, :] < batch["in_lens"][:, None])
    attn_soft, attn_logprob = oflow.conv_attention(sd, "attention.", batch["mel"], keys, key_mask, batch["attn_prior"])
    # the reference's round trip (radtts.py:326-334): D2H of attn_soft, serial host-side MAS, H2D of the hard map
    hard = torch.from_numpy(omas.binarize(attn_soft.detach().float().cpu().numpy(), batch["in_lens"].cpu().numpy(),
                                          batch["out_lens"].cpu().numpy(), is_prob=True)).to(dev)
    context = oflow.attention_context(text_enc, hard)                                  # consumer of the hard map
    ctx = oflow.squeeze_time(context, 2)
    ctx = torch.cat((ctx, spk[:, :, None].expand(-1, -1, ctx.shape[2])), 1)            # (B, 1040, T')
    z, logdets, log_s = oflow.decoder_forward(sd, batch["mel"], ctx, batch["out_lens"])
    loss, _ = oflow.flow_loss(z, logdets, log_s, batch["out_lens"])
    bin_loss = -(torch.log(attn_soft.clamp_min(1e-45)) * hard).sum() / hard.sum()
    total = loss + bin_loss
    if backward:
        total.backward()
    return float(total.detach()), int(batch["out_lens"].sum())


def time_steps(sd, batch, steps=1, warmup=1, backward=True, threads=None):
    if threads:
        torch.set_num_threads(threads)
    sd = {k: (v.clone().requires_grad_(True) if (k.startswith(("flows.", "attention.")) and v.dtype.is_floating_point
                                                 and not k.endswith((".p", "lower_diag"))) else v)
          for k, v in sd.items()}
    keys, text_enc, spk = make_inputs(sd, batch)
    for _ in range(warmup):
        hot_path_step(sd, batch, keys, text_enc, spk, backward)
    t0 = time.perf_counter()
    frames = 0
    for _ in range(steps):
        for v in sd.values():
            if v.requires_grad:
                v.grad = None
        _, f = hot_path_step(sd, batch, keys, text_enc, spk, backward)
        frames += f
    dt = time.perf_counter() - t0
    return frames / dt, dt / steps, frames // max(steps, 1)
