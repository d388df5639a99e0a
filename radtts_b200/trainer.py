"""One optimisation step of RADTTS on the B200 hot path, mirroring the reference training loop body
(reference train.py:385-422): forward under autocast -> RADTTSLoss (+ binarization loss) -> backward ->
gradient all-reduce (DDP/NCCL) -> clip -> optimizer step."""
import os

import torch

from . import loss as rloss


def bulk_first_order(model):
    """Flat-buffer order of the trainable parameters for data parallelism: the decoder flows' parameter networks first
    (94 % of the bytes; their gradients are final when the flow stack's backward returns), everything else after --
    two contiguous all-reduce regions.  Returns (parameters in that order, number of flat-buffer ELEMENTS of the first
    region, every tensor padded to a multiple of 4 elements exactly as FusedRAdam lays the buffer out)."""
    params = [p for p in model.parameters() if p.requires_grad]
    bulk_ids, bulk = set(), []
    for f in model.flows:
        tfn = getattr(f, "affine_tfn", None)
        net = getattr(tfn, "affine_param_predictor", None)
        for p in ([] if net is None else net.parameters()):
            if p.requires_grad and id(p) not in bulk_ids:
                bulk_ids.add(id(p))
                bulk.append(p)
    ordered = bulk + [p for p in params if id(p) not in bulk_ids]
    return ordered, sum((p.numel() + 3) // 4 * 4 for p in bulk)


class TrainStep:
    def __init__(self, model, loss_weights, lr=1e-4, weight_decay=1e-6, grad_clip_val=1.0, bf16=True,
                 binarize_attention=True, use_binarization_loss=True, ddp=False, device_ids=None, capturable=False,
                 fused_optimizer=True):
        self.raw_model = model
        self.model = model
        self.world = 1
        if ddp:
            import torch.distributed as dist
            self.world = dist.get_world_size()
            if not fused_optimizer:
                from torch.nn.parallel import DistributedDataParallel as DDP
                self.model = DDP(model, device_ids=device_ids, bucket_cap_mb=128, gradient_as_bucket_view=True)
            else:
                # replicas start identical (reference distributed.py:111-114 broadcasts every tensor from rank 0)
                for t in list(model.parameters()) + list(model.buffers()):
                    dist.broadcast(t.data, 0)
        self.criterion = rloss.RADTTSLoss(sigma=1.0, n_group_size=model.n_group_size, loss_weights=loss_weights)
        self.bin_loss = rloss.AttentionBinarizationLoss()
        if next(model.parameters()).is_cuda and hasattr(model, "attention"):
            # launch the attention CTC kernel as soon as ConvAttention is done, on a side stream (captured as a parallel
            # branch of the step's CUDA graph): it then overlaps the rest of the forward pass
            self._ctc_hook = self.criterion.attn_ctc_loss.prefetch_from(model.attention)
        self.loss_weights = loss_weights
        self.bf16 = bf16
        self.binarize = binarize_attention
        self.use_bin_loss = use_binarization_loss
        self.grad_clip_val = grad_clip_val
        params = [p for p in model.parameters() if p.requires_grad]
        self.n_bulk = 0
        if ddp and fused_optimizer and hasattr(model, "flows"):
            params, self.n_bulk = bulk_first_order(model)
        self.capturable = bool(capturable)
        self.fused_optimizer = bool(fused_optimizer)
        if self.fused_optimizer:
            from . import ops
            from .optim import FusedRAdam
            self.optimizer = FusedRAdam(params, lr=lr, weight_decay=weight_decay)
            # gradients live in the optimizer's flat buffer and this class reduces them itself (no DDP hooks): let
            # the flow stack's backward add weight_v / weight_g gradients straight into it
            ops.set_direct_grad_accumulation(True)
            if self.world > 1 and self.n_bulk > 0 and not os.environ.get("RADTTS_NO_COMM_OVERLAP"):
                dev = next(model.parameters()).device
                self.comm = torch.cuda.Stream(device=dev)
                # external: inside the captured step this becomes an event-record NODE that the eagerly launched NCCL
                # all-reduce (comm stream) can wait on -- the collective then overlaps the rest of the backward graph
                self.ev_bulk = torch.cuda.Event(external=True)
                ops.flow_backward_done = lambda: self.ev_bulk.record(torch.cuda.current_stream(dev))
        else:
            self.optimizer = torch.optim.RAdam(params, lr=lr, weight_decay=weight_decay, foreach=True,
                                               capturable=self.capturable)
        self.graph = None
        self.graph_update = None
        self.launches_per_replay = 0
        self.static_batch = None
        self.static_loss = None

    def forward_loss(self, batch):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.bf16):
            out = self.model(batch["mel"], batch["speaker_ids"], batch["text"], batch["in_lens"], batch["out_lens"],
                             binarize_attention=self.binarize, attn_prior=batch["attn_prior"],
                             f0=batch.get("f0"), energy_avg=batch.get("energy_avg"),
                             voiced_mask=batch.get("voiced_mask"), p_voiced=batch.get("p_voiced"))
            losses = self.criterion(out, batch["in_lens"], batch["out_lens"])
            total = None
            for v, w in losses.values():
                if w > 0:
                    total = v * w if total is None else total + v * w
            if self.binarize and self.use_bin_loss:
                total = total + self.bin_loss(out["attn"], out["attn_soft"]) * self.loss_weights["binarization_loss_weight"]
        return total, out

    def capture(self, example_batch, warmup=3):
        """Captures the step into CUDA graphs.  Possible because the step has no host synchronisation left
        (device-side frame plans, fused CTC, persistent LSTM); requires a fixed batch shape -- lengths stay device
        data and may change from replay to replay.
        One process: ONE graph (forward, losses, backward, clip, RAdam).  Data parallel: TWO graphs with the NCCL
        gradient all-reduce launched eagerly between them (collectives inside a capture hung in testing)."""
        assert self.capturable and self.fused_optimizer, "CUDA graphs need TrainStep(capturable=True, fused_optimizer=True)"
        from . import _lib
        import gc
        gc.collect()   # drop autograd graphs of earlier eager steps: their nodes remember the stream they were built on
        self.static_batch = {k: v.clone() for k, v in example_batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step(self.static_batch)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            self.static_loss = self._fwd_bwd(self.static_batch)
            if self.world == 1:
                self._update()
        self.graph = graph
        if self.world > 1:
            self.graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_update, pool=graph.pool()):
                self._update()
        self.launches_per_replay = _lib.launch_count() - n0   # kernels of libradtts_b200.so inside one replay

    def step(self, batch):
        if self.graph is not None:
            for k, v in batch.items():
                self.static_batch[k].copy_(v, non_blocking=True)
            self.graph.replay()
            if self.world > 1:
                self._allreduce()
                self.graph_update.replay()
            return self.static_loss
        return self._eager_step(batch)

    def _fwd_bwd(self, batch):
        self.optimizer.zero_grad(set_to_none=not self.fused_optimizer)
        total, _ = self.forward_loss(batch)
        total.backward()
        return total.detach()

    def _allreduce(self):
        # data-parallel gradient exchange on the flat buffer: the reference does one flat all-reduce after backward
        # (distributed.py:133-140); 128 MB chunks let NCCL pipeline over NVLink / NVSwitch.  The flow stack's region
        # (first n_bulk elements) is reduced on the comm stream as soon as ev_bulk fires, i.e. while the LSTM /
        # attention / encoder part of the backward pass is still running; the remainder follows at the end.
        import torch.distributed as dist
        g = self.optimizer.grad
        n_bulk = self.n_bulk if getattr(self, "comm", None) is not None else 0
        if n_bulk > 0:
            cur = torch.cuda.current_stream(g.device)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self.ev_bulk)
                for chunk in g[:n_bulk].split(32 << 20):
                    dist.all_reduce(chunk)
            for chunk in g[n_bulk:].split(32 << 20):
                dist.all_reduce(chunk)
            cur.wait_stream(self.comm)
        else:
            for chunk in g.split(32 << 20):
                dist.all_reduce(chunk)

    def _update(self):
        if self.grad_clip_val > 0:
            # with data parallelism the buffer holds the SUM over ranks: ||mean|| = ||sum|| / world, and the mean
            # itself is folded into the scale the optimizer kernel applies
            norm = torch.linalg.vector_norm(self.optimizer.grad) / self.world
            scale = ((self.grad_clip_val / (norm + 1e-6)).clamp(max=1.0) / self.world).reshape(1)
        else:
            scale = torch.full((1,), 1.0 / self.world, device=self.optimizer.grad.device) if self.world > 1 else None
        self.optimizer.step(scale)

    def _eager_step(self, batch):
        if self.fused_optimizer:
            loss = self._fwd_bwd(batch)
            if self.world > 1:
                self._allreduce()
            self._update()
            return loss
        self.optimizer.zero_grad(set_to_none=True)
        total, _ = self.forward_loss(batch)
        total.backward()
        if self.grad_clip_val > 0:
            torch.nn.utils.clip_grad_norm_(self.raw_model.parameters(), self.grad_clip_val, foreach=True)
        self.optimizer.step()
        return total.detach()
