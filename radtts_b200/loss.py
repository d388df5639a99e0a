"""Loss functions with the reference's interface (reference loss.py), free of host synchronisation points.

compute_flow_loss / RADTTSLoss follow loss.py:27-52,147-203; AttentionCTCLoss is the batched equivalent of the
per-utterance loop of loss.py:111-135; AttentionBinarizationLoss replaces the boolean-mask gather of
loss.py:138-144 by a masked sum.  Attribute-predictor losses (DAP/BGAP regression and flow NLL) are included for
config parity; they are ordinary PyTorch (outside the hot path).
"""
import torch
import torch.nn as nn
from torch.nn import functional as F

from .common import get_mask_from_lengths


def compute_flow_loss(z, log_det_W_list, log_s_list, n_elements, n_dims, mask, sigma=1.0, packed=None):
    """(prior NLL - sum log_s - n_elements * sum log|det W|) / (n_elements * n_dims); does not mutate its inputs.
    packed = (z_packed, [log_s_packed, ...]): the same tensors in the packed-row layout of the decoder kernels, whose
    padding rows are exactly zero -- plain sums, no mask (only valid when `mask` is the decoder's own frame mask)."""
    log_s_total = 0.0
    if packed is not None:
        for log_s in packed[1]:
            log_s_total = log_s_total + torch.sum(log_s)
    else:
        for log_s in log_s_list:
            log_s_total = log_s_total + torch.sum(log_s * mask)
    log_det_total = 0.0
    if len(log_det_W_list):
        log_det_total = torch.stack([ld.reshape(()) for ld in log_det_W_list]).sum() * n_elements
    if packed is not None:
        z = packed[0]
    else:
        z = z * mask
    prior_nll = torch.sum(z * z) / (2 * sigma * sigma)
    denom = n_elements * n_dims
    return (prior_nll - log_s_total - log_det_total) / denom, prior_nll / denom


def compute_regression_loss(x_hat, x, mask, name=False):
    x = x[:, None] if x.dim() == 2 else x
    mask = mask[:, None] if mask.dim() == 2 else mask
    x = x * mask
    x_hat = x_hat * mask
    if name == "vpred":
        loss = F.binary_cross_entropy_with_logits(x_hat, x, reduction="sum")
    else:
        loss = F.mse_loss(x_hat, x, reduction="sum")
    return {"loss_{}".format(name): loss / mask.sum()}


class AttributePredictionLoss(nn.Module):
    def __init__(self, name, model_config, loss_weight, sigma=1.0):
        super().__init__()
        self.name = name
        self.sigma = sigma
        self.model_name = model_config["name"]
        self.loss_weight = loss_weight
        self.n_group_size = model_config["hparams"].get("n_group_size", 1)

    def forward(self, model_output, lens):
        mask = get_mask_from_lengths(lens // self.n_group_size)[:, None].float()
        out = {}
        if "z" in model_output:
            n_elements = lens.sum() // self.n_group_size
            n_dims = model_output["z"].size(1)
            loss, loss_prior = compute_flow_loss(model_output["z"], model_output["log_det_W_list"],
                                                 model_output["log_s_list"], n_elements, n_dims, mask, self.sigma)
            out = {"loss_{}".format(self.name): (loss, self.loss_weight),
                   "loss_prior_{}".format(self.name): (loss_prior, 0.0)}
        elif "x_hat" in model_output:
            for k, v in compute_regression_loss(model_output["x_hat"], model_output["x"], mask, self.name).items():
                out[k] = (v, self.loss_weight)
        if not out:
            raise Exception("loss not supported")
        return out


class AttentionCTCLoss(nn.Module):
    """CTC over the (frames x tokens+blank) alignment lattice, one batched call.

    Equivalent to the reference's loop (loss.py:121-133): class 0 is a blank with constant log-prob
    `blank_logprob`, classes beyond an utterance's key_len are excluded from its log-softmax, each utterance's
    loss is divided by its target length (nn.CTCLoss reduction='mean' on a batch of one) and the batch is averaged."""

    def __init__(self, blank_logprob=-1):
        super().__init__()
        self.blank_logprob = blank_logprob
        self._prefetched = None     # (attn_logprob tensor, loss, side stream)
        self._side = None

    def forward(self, attn_logprob, in_lens, out_lens):
        B, _, T1, T2 = attn_logprob.shape
        pre, self._prefetched = self._prefetched, None
        if pre is not None and pre[0] is attn_logprob:
            torch.cuda.current_stream(attn_logprob.device).wait_stream(pre[2])
            return pre[1]
        if attn_logprob.is_cuda and 2 * T2 + 1 <= 1024:
            from . import ops
            return ops.attention_ctc_loss(attn_logprob, in_lens, out_lens, self.blank_logprob)
        return self.forward_torch(attn_logprob, in_lens, out_lens)

    def prefetch_from(self, attention_module):
        """Overlap: the fused CTC kernel occupies one SM per utterance for ~1 ms (a serial alpha/beta recursion over the
        frames) and needs nothing but attn_logprob.  This hooks the model's ConvAttention so that the kernel is
        launched on a side stream the moment the attention is computed, and runs underneath MAS / the context LSTM /
        the decoder flows; forward() later just joins the stream.  Returns the hook handle."""
        def hook(module, args, kwargs, output):
            attn_logprob = output[1]
            if not (torch.is_grad_enabled() and attn_logprob.is_cuda and 2 * attn_logprob.shape[3] + 1 <= 1024):
                return None
            out_lens = args[2]
            in_lens = kwargs.get("key_lens", args[4] if len(args) > 4 else None)
            if in_lens is None or out_lens is None:
                return None
            from . import ops
            dev = attn_logprob.device
            if self._side is None:
                self._side = torch.cuda.Stream(device=dev)
            cur = torch.cuda.current_stream(dev)
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                loss = ops.attention_ctc_loss(attn_logprob, in_lens, out_lens, self.blank_logprob)
            self._prefetched = (attn_logprob, loss, self._side)
            return None
        return attention_module.register_forward_hook(hook, with_kwargs=True)

    def forward_torch(self, attn_logprob, in_lens, out_lens):
        """Batched PyTorch formulation (CPU tensors, or text longer than the fused kernel supports)."""
        B, _, T1, T2 = attn_logprob.shape
        lp = F.pad(attn_logprob[:, 0], (1, 0), value=self.blank_logprob)                 # (B, T1, T2+1)
        cls = torch.arange(T2 + 1, device=lp.device)[None, None, :]
        # classes beyond key_len are excluded from the softmax; a large finite negative (exp underflows to exactly 0)
        # instead of -inf keeps the CTC backward free of (-inf) - (-inf)
        lp = lp.float().masked_fill(cls > in_lens[:, None, None], -1e4)
        lp = F.log_softmax(lp, dim=2).permute(1, 0, 2)                                   # (T1, B, T2+1)
        targets = torch.arange(1, T2 + 1, device=lp.device)[None, :].expand(B, -1)
        losses = F.ctc_loss(lp, targets, out_lens, in_lens, blank=0, reduction="none", zero_infinity=True)
        return (losses / in_lens.clamp(min=1).to(losses.dtype)).mean()


class AttentionBinarizationLoss(nn.Module):
    def forward(self, hard_attention, soft_attention):
        on_path = hard_attention == 1
        log_sum = torch.where(on_path, torch.log(soft_attention.clamp_min(1e-45)), torch.zeros_like(soft_attention)).sum()
        return -log_sum / hard_attention.sum()


class RADTTSLoss(nn.Module):
    def __init__(self, sigma=1.0, n_group_size=1, dur_model_config=None, f0_model_config=None,
                 energy_model_config=None, vpred_model_config=None, loss_weights=None):
        super().__init__()
        self.sigma = sigma
        self.n_group_size = n_group_size
        self.loss_weights = loss_weights
        self.attn_ctc_loss = AttentionCTCLoss(blank_logprob=loss_weights.get("blank_logprob", -1))
        self.loss_fns = {}
        for key, name, cfg, wkey in (("duration_model_outputs", "duration", dur_model_config, "dur_loss_weight"),
                                     ("f0_model_outputs", "f0", f0_model_config, "f0_loss_weight"),
                                     ("energy_model_outputs", "energy", energy_model_config, "energy_loss_weight"),
                                     ("vpred_model_outputs", "vpred", vpred_model_config, "vpred_loss_weight")):
            if cfg is not None:
                self.loss_fns[key] = AttributePredictionLoss(name, cfg, loss_weights[wkey])

    def forward(self, model_output, in_lens, out_lens):
        loss_dict = {}
        if len(model_output["z_mel"]):
            n_elements = out_lens.sum() // self.n_group_size
            z = model_output["z_mel"]
            mask = get_mask_from_lengths(out_lens // self.n_group_size, z.shape[2])[:, None].float()
            packed = None
            rb = getattr(z, "_rb_packed", None)
            ls_list = model_output["log_s_list"]
            if rb is not None and rb[3] is out_lens and len(rb[2]) == len(ls_list) and \
                    all(a is b for a, b in zip(rb[2], ls_list)):
                packed = (rb[0], rb[1])    # produced by ops.decoder_forward for exactly these tensors and lengths
            loss_mel, loss_prior_mel = compute_flow_loss(z, model_output["log_det_W_list"], ls_list,
                                                         n_elements, z.size(1), mask, self.sigma, packed=packed)
            loss_dict["loss_mel"] = (loss_mel, 1.0)
            loss_dict["loss_prior_mel"] = (loss_prior_mel, 0.0)
        ctc = self.attn_ctc_loss(model_output["attn_logprob"], in_lens, out_lens)
        loss_dict["loss_ctc"] = (ctc, self.loss_weights["ctc_loss_weight"])
        for k, fn in self.loss_fns.items():
            mout = model_output.get(k)
            if mout is not None and len(mout) > 0:
                for name, v in fn(mout, in_lens if "dur" in k else out_lens).items():
                    loss_dict[name] = v
        return loss_dict
