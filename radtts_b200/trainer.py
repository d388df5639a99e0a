"""One optimisation step of RADTTS on the B200 hot path, mirroring the reference training loop body
(reference train.py:385-422): forward under autocast -> RADTTSLoss (+ binarization loss) -> backward ->
gradient all-reduce (DDP/NCCL) -> clip -> optimizer step."""
import os

import torch

from . import loss as rloss
from . import parallel


def bulk_first_order(model):
    """(parameters in data-parallel flat-buffer order, ELEMENTS of the flow-network region in front) -- see
    parallel.flow_first_order, which also gives the per-flow sub-ranges."""
    ordered, regions = parallel.flow_first_order(model)
    return ordered, (regions[-1][1] if regions else 0)


class TrainStep:
    def __init__(self, model, loss_weights, lr=1e-4, weight_decay=1e-6, grad_clip_val=1.0, bf16=True,
                 binarize_attention=True, use_binarization_loss=True, ddp=False, device_ids=None, capturable=False,
                 fused_optimizer=True, probe_batch=None, loss_kwargs=None, deferred_update=False):
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
        self.criterion = rloss.RADTTSLoss(sigma=1.0, n_group_size=getattr(model, "n_group_size", 1),
                                          loss_weights=loss_weights, **(loss_kwargs or {}))
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
        self.capturable = bool(capturable)
        self.fused_optimizer = bool(fused_optimizer)
        self.comm = None
        self.flow_regions = []
        self.ev_flow = []
        self._flow_final = []
        params = [p for p in model.parameters() if p.requires_grad]
        if fused_optimizer and hasattr(model, "flows"):
            # flow parameter networks first (94 % of the bytes): the region data parallelism reduces flow by flow, and the
            # one a deferred update applies underneath the next step's text encoder / attention / context LSTM
            params, self.flow_regions = parallel.flow_first_order(model)
        if probe_batch is not None:
            # the reference optimizer skips parameters whose .grad is None (radam.py:53-55) -- e.g. v_pred_module /
            # v_embeddings in decoder-only training.  One plain forward/backward finds them; they stay out of the flat
            # buffer, so neither weight decay nor the moment updates touch them.
            used = self._probe_used(probe_batch)
            params = [p for p in params if id(p) in used]
            if self.flow_regions:
                assert all(id(p) in used for f in model.flows for p in f.affine_tfn.affine_param_predictor.parameters())
        if self.fused_optimizer:
            from .optim import FusedRAdam
            self.optimizer = FusedRAdam(params, lr=lr, weight_decay=weight_decay)
            # gradients live in the optimizer's flat buffer and this class reduces them itself (no DDP hooks): the flow
            # stack's backward adds weight_v / weight_g gradients straight into it (ops.trainer_scope in _fwd_bwd)
            if self.world > 1 and self.flow_regions and not os.environ.get("RADTTS_NO_COMM_OVERLAP"):
                dev = next(model.parameters()).device
                self.comm = torch.cuda.Stream(device=dev)
                # external: inside the captured step these become event-record NODES that the eagerly launched NCCL
                # all-reduces (comm stream) wait on -- flow i's region is reduced while the graph is still running the
                # backward of flows i-1 .. 0, the LSTMs, the attention and the encoder
                self.ev_flow = [torch.cuda.Event(external=True) for _ in self.flow_regions]
                self._flow_final = [False] * len(self.flow_regions)
        else:
            self.optimizer = torch.optim.RAdam(params, lr=lr, weight_decay=weight_decay, foreach=True,
                                               capturable=self.capturable)
        # Deferred update (fused optimizer only).  The RAdam pass is HBM-bound (28 B per parameter, ~1 ms for 226 M) and
        # nothing can overlap it at the end of a step: the clip coefficient needs every gradient, the next forward needs
        # the parameters.  But the next forward needs the FLOW parameters only after the text encoder, the attention, MAS
        # and the context LSTM -- ~2 ms that leave most SMs and nearly all of HBM idle.  So a step updates the small
        # remainder at its end and leaves the flow region PENDING; the next step applies it first thing on the side stream
        # the flow weight preparation runs on.  Same kernels on the same data in a different order: parameters are
        # identical to the plain schedule once flush() (or the next step) has run.  step_dev counts APPLIED bulk updates.
        self.deferred_update = bool(deferred_update) and self.fused_optimizer and bool(self.flow_regions)
        self.n_bulk = self.flow_regions[-1][1] if self.flow_regions else 0
        self._grads_clean = False
        if self.fused_optimizer:
            dev = self.optimizer.flat.device
            self.pending_dev = torch.zeros(1, dtype=torch.int32, device=dev)
            self.scale_dev = torch.ones(1, dtype=torch.float32, device=dev)
        self.graph = None
        self.graph_update = None
        self.launches_per_replay = 0
        self.static_batch = None
        self.static_loss = None

    def _probe_used(self, batch):
        self.raw_model.zero_grad(set_to_none=True)
        # on a side stream, like capture()'s warm-up: autograd nodes remember the stream they were created on, and an eager
        # backward on the legacy default stream before a capture makes the captured backward fail (implicit-sync error)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            total, _ = self.forward_loss(batch)
            total.backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        used = {id(p) for p in self.raw_model.parameters() if p.grad is not None}
        self.raw_model.zero_grad(set_to_none=True)
        self.criterion.attn_ctc_loss._prefetched = None
        return used

    def _on_flow_grads(self, i, final):
        if i < len(self.ev_flow):
            self._flow_final[i] = bool(final)
            if final:
                self.ev_flow[i].record(torch.cuda.current_stream(self.optimizer.grad.device))

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
        from . import ops
        ops.POOL.end_capture()

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

    def flush(self):
        """Applies a pending deferred update (no-op otherwise): call before reading parameters outside step() -- checkpoint,
        evaluation, the end of training."""
        if self.deferred_update:
            self.optimizer.step_range(0, self.n_bulk, self.scale_dev, enable=self.pending_dev, zero_grad=True, bump=True)
            self.pending_dev.zero_()

    def _fwd_bwd(self, batch):
        from . import ops
        if self.fused_optimizer:
            if self.deferred_update:
                dev = self.optimizer.flat.device
                side = ops.prep_stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):      # begin_decoder_prep queues the flow weight preparation behind it
                    self.optimizer.step_range(0, self.n_bulk, self.scale_dev, enable=self.pending_dev, zero_grad=True,
                                              bump=True)
            elif not self._grads_clean:
                self.optimizer.zero_grad(set_to_none=False)
            self._grads_clean = False
        else:
            self.optimizer.zero_grad(set_to_none=True)
        # (fused optimizer: the flat gradient buffer is zero here -- memset above or zeroed by the previous update -- and
        # this is the step's only backward, so weight-norm gradients are stored, not accumulated)
        with ops.trainer_scope(self.fused_optimizer, self._on_flow_grads if self.ev_flow else None,
                               grads_zeroed=self.fused_optimizer):
            total, _ = self.forward_loss(batch)
            total.backward()
        return total.detach()

    def _allreduce(self):
        # data-parallel gradient exchange on the flat buffer (parallel.allreduce_flat): flow i's region is reduced on
        # the comm stream as soon as ITS event fires (flows finish 7 .. 0), underneath the rest of the backward pass;
        # the non-flow remainder (6 % of the bytes) follows at the end.  Regions whose gradients did not all take the
        # direct route (a frozen parameter, a non-fp32 .grad) are not final at their event and go with the remainder.
        order = list(reversed(range(len(self.flow_regions))))
        ready = [self.ev_flow[i] if (self.ev_flow and self._flow_final[i]) else None for i in order]
        parallel.allreduce_flat(self.optimizer.grad, [self.flow_regions[i] for i in order], ready, self.comm)

    def _update(self):
        if self.grad_clip_val > 0:
            # with data parallelism the buffer holds the SUM over ranks: ||mean|| = ||sum|| / world, and the mean
            # itself is folded into the scale the optimizer kernel applies
            norm = torch.linalg.vector_norm(self.optimizer.grad) / self.world
            scale = ((self.grad_clip_val / (norm + 1e-6)).clamp(max=1.0) / self.world).reshape(1)
        else:
            scale = torch.full((1,), 1.0 / self.world, device=self.optimizer.grad.device) if self.world > 1 else None
        if not self.fused_optimizer:
            self.optimizer.step(scale)
            return
        # the coefficient lives in a persistent buffer (a deferred update reads it at the top of the NEXT replay, when
        # a tensor from the graph's private pool may already have been reused); the update writes zeros over the
        # gradients it consumed, so the next step starts without a memset of the 0.9 GB buffer
        if scale is None:
            self.scale_dev.fill_(1.0)
        else:
            self.scale_dev.copy_(scale)
        total = self.optimizer.flat.numel()
        if self.deferred_update:
            self.optimizer.step_range(self.n_bulk, total, self.scale_dev, zero_grad=True, bump=False)
            self.pending_dev.fill_(1)
        else:
            self.optimizer.step_range(0, total, self.scale_dev, zero_grad=True, bump=True)
        self._grads_clean = True

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
