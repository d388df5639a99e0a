"""Bridge between the reference-shaped nn.Modules and the C-ABI CUDA library.

Everything here is plumbing: tensor allocation, ctypes structs, autograd.Function wrappers and the packed
frame-plan bookkeeping.  All math of the hot path happens inside libradtts_b200.so; if the library is
missing or an input is not on a CUDA device these functions raise (no CPU fallback).
"""
import ctypes
import os

import torch

from . import _lib

PREC_FP32, PREC_BF16 = 0, 1
MAX_LAYERS = 8
_SCALING = {"tanh": 0, "exp": 1, "sigmoid": 2, "translate": 3}

_forced_precision = None


_direct_grad = False
_grads_zeroed = False       # set by a trainer that KNOWS every .grad it sinks into is zero when backward starts (see below)
flow_grads_ready = None     # optional callable(flow_index, final): invoked by the flow stack's backward right after the
                            # last kernel producing flow `flow_index`'s parameter-network gradients has been enqueued;
                            # final=True when every one of them went straight into .grad (direct accumulation)


def set_direct_grad_accumulation(on):
    """on: the flow stack's backward adds the weight_v / weight_g gradients directly into the parameters' .grad
    tensors (when they exist, are fp32 and contiguous) and returns no gradient for them to autograd.  Equivalent to
    what AccumulateGrad would do, minus one elementwise launch per parameter; hooks on those parameters (DDP) do not
    fire, so only trainers that reduce gradients themselves (radtts_b200.trainer) switch it on."""
    global _direct_grad
    _direct_grad = bool(on)


class trainer_scope:
    """Context manager a trainer wraps around ITS forward + backward: direct gradient accumulation and the per-flow
    `flow_grads_ready` callback are process-global switches, so they are set on entry and restored on exit -- another
    TrainStep, or any other user of the flow stack in the same process, never inherits them."""

    def __init__(self, direct_grad, on_flow_grads=None, grads_zeroed=False):
        """grads_zeroed: the trainer guarantees that the .grad buffers are zero when backward starts and that there is ONE
        backward per optimizer step (TrainStep with the fused optimizer: the update zeroes them in the pass that consumes
        them).  The weight-norm backward then STORES weight_v / weight_g gradients instead of read-modify-writing 118 MB of
        zeros per flow."""
        self.new = (bool(direct_grad), on_flow_grads, bool(grads_zeroed) and bool(direct_grad))

    def __enter__(self):
        global _direct_grad, flow_grads_ready, _grads_zeroed
        self.prev = (_direct_grad, flow_grads_ready, _grads_zeroed)
        _direct_grad, flow_grads_ready, _grads_zeroed = self.new
        return self

    def __exit__(self, *exc):
        global _direct_grad, flow_grads_ready, _grads_zeroed
        _direct_grad, flow_grads_ready, _grads_zeroed = self.prev
        return False


def set_precision(p):
    """None: follow torch autocast (bf16 inside autocast, fp32 otherwise); 'fp32' / 'bf16': force."""
    global _forced_precision
    assert p in (None, "fp32", "bf16")
    _forced_precision = p


def current_precision():
    if _forced_precision is not None:
        return PREC_BF16 if _forced_precision == "bf16" else PREC_FP32
    return PREC_BF16 if torch.is_autocast_enabled() else PREC_FP32


def _act_dtype(prec):
    return torch.bfloat16 if prec == PREC_BF16 else torch.float32


# ------------------------------------------------------------------------------------------------------
# ctypes mirrors of the public structs
# ------------------------------------------------------------------------------------------------------
class FlowDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("z_ld", "c_off", "c_active", "n_ctx", "n_ch", "n_layers", "ksize", "partial_padding", "scaling")]


_P = ctypes.c_void_p


class FlowWeights(ctypes.Structure):
    _fields_ = [("w_inv", _P), ("w_start", _P), ("b_start", _P), ("w_in", _P * MAX_LAYERS), ("b_in", _P * MAX_LAYERS),
                ("w_rs", _P * MAX_LAYERS), ("b_rs", _P * MAX_LAYERS), ("w_end", _P), ("b_end", _P),
                ("wg_start", _P), ("wg_in", _P * MAX_LAYERS), ("wg_rs", _P * MAX_LAYERS)]


class FlowWnGrads(ctypes.Structure):
    _fields_ = [("gv_start", _P), ("gg_start", _P), ("gv_in", _P * MAX_LAYERS), ("gg_in", _P * MAX_LAYERS),
                ("gv_rs", _P * MAX_LAYERS), ("gg_rs", _P * MAX_LAYERS)]


class FlowBuffers(ctypes.Structure):
    _fields_ = [(n, _P) for n in ("ctx", "zin", "zmid", "zout", "z0", "x", "r", "params", "log_s")]


class FlowGradBuffers(ctypes.Structure):
    _fields_ = [(n, _P) for n in ("g_zout", "g_log_s", "g_zin", "g_ctx", "g_zmid", "g_params", "g_u", "g_v", "g_x0",
                                  "g_w_inv_full", "g_w_start", "g_b_start")] + \
               [("g_w_in", _P * MAX_LAYERS), ("g_b_in", _P * MAX_LAYERS), ("g_w_rs", _P * MAX_LAYERS),
                ("g_b_rs", _P * MAX_LAYERS), ("g_w_end", _P), ("g_b_end", _P), ("scratch_f32", _P)]


def _p(t):
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------------
# plain data movement used by the module API (views, no kernels needed)
# ------------------------------------------------------------------------------------------------------
def squeeze_time(x, g):
    """z[b, c*g+k, t'] = x[b, c, g*t'+k]  (reference nn.Unfold((g,1), stride g), radtts.py:165-169)."""
    b, c, t = x.shape
    tt = t // g
    return x[:, :, :tt * g].reshape(b, c, tt, g).permute(0, 1, 3, 2).reshape(b, c * g, tt)


def unsqueeze_time(z, g):
    b, cg, tt = z.shape
    return z.reshape(b, cg // g, g, tt).permute(0, 1, 3, 2).reshape(b, cg // g, tt * g)


# ------------------------------------------------------------------------------------------------------
# frame plan + pack / unpack
# ------------------------------------------------------------------------------------------------------
class FramePlan:
    """Device-side description of the packed frame layout (see csrc/frameplan.cuh)."""

    def __init__(self, lens, divisor, tmax, geom_lens=None):
        _lib.require_cuda(lens)
        L = _lib.lib()
        self.B = int(lens.shape[0])
        self.Tmax = int(tmax)
        self.divisor = int(divisor)
        self.rows = int(L.radtts_frameplan_rows(self.B, self.Tmax))
        nbytes = int(L.radtts_frameplan_bytes(self.B, self.Tmax))
        self.buf = torch.empty(nbytes // 4, dtype=torch.int32, device=lens.device)
        self.lens = lens.to(torch.int64).contiguous()
        self.geom_lens = None if geom_lens is None else geom_lens.to(device=lens.device, dtype=torch.int64).contiguous()
        _lib.check(L.radtts_frameplan_build(_lib.ptr(self.lens), _lib.ptr(self.geom_lens), self.divisor, self.B,
                                            self.Tmax, _lib.ptr(self.buf), _lib.stream_of(lens)),
                   "radtts_frameplan_build")

    @property
    def ptr(self):
        return _lib.ptr(self.buf)

    def frame_lens(self):
        return torch.div(self.lens, self.divisor, rounding_mode="floor").clamp(0, self.Tmax)


def pack_frames(src, plan, g, dtype=torch.float32, ld=None, col_off=0, ncols_pad=None, out=None, valid_only=False):
    """(B, C, T) fp32 -> packed [rows][ld]; columns [col_off, col_off + ncols_pad) are written."""
    _lib.require_cuda(src)
    src = src.float().contiguous()
    B, C, T = src.shape
    ncols = C * g
    ncols_pad = ncols if ncols_pad is None else ncols_pad
    ld = col_off + ncols_pad if ld is None else ld
    if out is None:
        out = torch.zeros((plan.rows, ld), dtype=dtype, device=src.device) if ld != col_off + ncols_pad or col_off \
            else torch.empty((plan.rows, ld), dtype=dtype, device=src.device)
    L = _lib.lib()
    _lib.check(L.radtts_pack_frames(_lib.ptr(src), B, C, T, g, plan.ptr, plan.Tmax, _lib.ptr(out),
                                    int(out.dtype == torch.bfloat16), ld, col_off, ncols_pad, int(valid_only),
                                    _lib.stream_of(src)),
               "radtts_pack_frames")
    return out


def unpack_frames(packed, plan, C, g, col_off=0):
    """packed fp32 [rows][ld] -> (B, C, Tmax*g) fp32 (zeros beyond each utterance)."""
    _lib.require_cuda(packed)
    assert packed.dtype == torch.float32 and packed.is_contiguous()
    out = torch.empty((plan.B, C, plan.Tmax * g), dtype=torch.float32, device=packed.device)
    L = _lib.lib()
    _lib.check(L.radtts_unpack_frames(_lib.ptr(packed), packed.shape[1], col_off, plan.ptr, plan.B, plan.Tmax, C, g,
                                      _lib.ptr(out), _lib.stream_of(packed)), "radtts_unpack_frames")
    return out


class _PackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, plan, g, dtype, ld, col_off, ncols_pad):
        ctx.plan, ctx.g, ctx.col_off, ctx.C = plan, g, col_off, src.shape[1]
        ctx.T = src.shape[2]
        return pack_frames(src, plan, g, dtype, ld, col_off, ncols_pad)

    @staticmethod
    def backward(ctx, grad):
        g = unpack_frames(grad.float().contiguous(), ctx.plan, ctx.C, ctx.g, ctx.col_off)
        if g.shape[2] != ctx.T:
            g = torch.nn.functional.pad(g, (0, ctx.T - g.shape[2])) if g.shape[2] < ctx.T else g[:, :, :ctx.T]
        return g, None, None, None, None, None, None


class _UnpackFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed, plan, C, g, col_off):
        ctx.plan, ctx.g, ctx.col_off, ctx.ld, ctx.C = plan, g, col_off, packed.shape[1], C
        return unpack_frames(packed, plan, C, g, col_off)

    @staticmethod
    def backward(ctx, grad):
        out = pack_frames(grad.contiguous(), ctx.plan, ctx.g, torch.float32, ctx.ld, ctx.col_off, ctx.C * ctx.g)
        return out, None, None, None, None


def pack(src, plan, g=1, dtype=torch.float32, ld=None, col_off=0, ncols_pad=None):
    return _PackFn.apply(src, plan, g, dtype, ld, col_off, ncols_pad)


def unpack(packed, plan, C, g=1, col_off=0):
    return _UnpackFn.apply(packed, plan, C, g, col_off)


# ------------------------------------------------------------------------------------------------------
# flow step
# ------------------------------------------------------------------------------------------------------
def _flow_dims(flow, z_ld, c_active):
    wn = flow.affine_tfn.affine_param_predictor
    scaling = flow.affine_tfn.scaling_fn
    if isinstance(scaling, list) or scaling not in _SCALING:
        raise NotImplementedError("per-channel scaling_fn lists are outside the fused flow-step kernel")
    return FlowDims(z_ld=z_ld, c_off=z_ld - c_active, c_active=c_active, n_ctx=wn.n_context_dim, n_ch=wn.n_channels,
                    n_layers=wn.n_layers, ksize=wn.kernel_size, partial_padding=int(wn.use_partial_padding),
                    scaling=_SCALING[scaling])


def _v_and_g(conv):
    """(weight or weight_v, weight_g or None) of a possibly weight-normed conv (torch.nn.utils.weight_norm, dim 0):
    the library applies g / ||v|| itself while re-laying the weight out, and derives both gradients."""
    if hasattr(conv, "weight_v"):
        return conv.weight_v, conv.weight_g
    return conv.weight, None


def _flow_weight_list(flow, inverse, w_inv=None):
    """Flat list of fp32 weight tensors in the order FlowWeights expects (3 + 4 n_layers + 2 entries: weight_v where a
    conv is weight-normed), followed by the 1 + 2 n_layers weight_g tensors (None where it is not).  w_inv: the 1x1
    matrix when the caller has already composed it (lus_compose_stack)."""
    wn = flow.affine_tfn.affine_param_predictor
    inv = flow.invtbl_conv
    if w_inv is not None:
        pass
    elif hasattr(inv, "lower"):
        w_inv = inv.inverse_weight() if inverse else inv.weight()
    else:
        w_inv = inv.inverse_weight() if inverse else inv.conv.weight.squeeze(-1)
    v_start, g_start = _v_and_g(wn.start)
    ws = [w_inv.float(), v_start, wn.start.bias]
    gs = [g_start]
    for i in range(wn.n_layers):
        v, g = _v_and_g(wn.in_layers[i].conv)
        ws += [v, wn.in_layers[i].conv.bias]
        gs.append(g)
    g_rs = []
    for i in range(wn.n_layers):
        v, g = _v_and_g(wn.res_skip_layers[i])
        ws += [v, wn.res_skip_layers[i].bias]
        g_rs.append(g)
    ws += [wn.end.weight, wn.end.bias]
    return ws + gs + g_rs


def n_weight_tensors(n_layers):
    return 3 + 4 * n_layers + 2


def _weights_struct(ws, n_layers):
    ws = [None if w is None else w.detach().float().contiguous() for w in ws]
    s = FlowWeights()
    s.w_inv, s.w_start, s.b_start = _p(ws[0]), _p(ws[1]), _p(ws[2])
    for i in range(n_layers):
        s.w_in[i], s.b_in[i] = _p(ws[3 + 2 * i]), _p(ws[4 + 2 * i])
        s.w_rs[i], s.b_rs[i] = _p(ws[3 + 2 * n_layers + 2 * i]), _p(ws[4 + 2 * n_layers + 2 * i])
    s.w_end, s.b_end = _p(ws[3 + 4 * n_layers]), _p(ws[4 + 4 * n_layers])
    nw = n_weight_tensors(n_layers)
    s.wg_start = _p(ws[nw])
    for i in range(n_layers):
        s.wg_in[i] = _p(ws[nw + 1 + i])
        s.wg_rs[i] = _p(ws[nw + 1 + n_layers + i])
    return s, ws  # keep `ws` alive until the kernels have been enqueued on the stream


def prepare_flow(dims, ws, prec, want_backward, device):
    L = _lib.lib()
    nbytes = int(L.radtts_flow_prepared_bytes(ctypes.byref(dims), prec, int(want_backward)))
    if nbytes == 0:
        raise _lib.RadttsB200Error("unsupported flow-step dimensions for the CUDA kernels")
    blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
    s, keep = _weights_struct(ws, dims.n_layers)
    _lib.check(L.radtts_flow_prepare(ctypes.byref(dims), ctypes.byref(s), 0, prec, int(want_backward),
                                     _lib.ptr(blob), ctypes.c_size_t(nbytes),
                                     ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)),
               "radtts_flow_prepare")
    return blob


class _Pool:
    """Zero-initialised scratch tensors, recycled across steps.  The kernels rely on every packed buffer being
    FINITE beyond the rows they write (stale values are multiplied by zero rows; NaN garbage would not be).

    Bounded: free buffers are kept per exact shape in least-recently-used order and the oldest shapes are dropped once
    the free bytes exceed `cap_bytes` (RADTTS_POOL_CAP_MB, default 8 GB of the 180 GB) -- variable-length training sees a
    new `rows` for every distinct padded batch length.  Buffers created while a CUDA graph is being captured live in the
    graph's private memory pool, warm eager buffers a capture takes over are pinned for the graph's lifetime, and both
    go to a separate free list that `end_capture()` drops, so eager steps never write into memory a captured graph
    replays on."""

    def __init__(self, cap_bytes=None):
        import collections
        self.free = collections.OrderedDict()
        self.cap_free = {}
        self.pinned = []
        self.free_bytes = 0
        if cap_bytes is None:
            cap_bytes = int(os.environ.get("RADTTS_POOL_CAP_MB", "8192")) << 20
        self.cap_bytes = cap_bytes

    @staticmethod
    def _key(shape, dtype, device):
        return (tuple(shape), dtype, device.type, device.index)

    def get(self, shape, dtype, device):
        key = self._key(shape, dtype, device)
        if device.type == "cuda" and torch.cuda.is_current_stream_capturing():
            lst = self.cap_free.get(key)
            if lst:
                return lst.pop()
            lst = self.free.get(key)
            if lst:
                # a warm eager buffer (already zero-initialised: no memset node in the graph); from now on it belongs
                # to captured graphs only and is kept alive for them
                t = lst.pop()
                self.free_bytes -= t.numel() * t.element_size()
                if not lst:
                    del self.free[key]
                self.pinned.append(t)
            else:
                t = torch.zeros(shape, dtype=dtype, device=device)
            t._rb_captured = True
            return t
        lst = self.free.get(key)
        if lst:
            t = lst.pop()
            self.free_bytes -= t.numel() * t.element_size()
            if not lst:
                del self.free[key]
            else:
                self.free.move_to_end(key)
            return t
        return torch.zeros(shape, dtype=dtype, device=device)

    def put(self, t):
        key = self._key(t.shape, t.dtype, t.device)
        if getattr(t, "_rb_captured", False):
            self.cap_free.setdefault(key, []).append(t)
            return
        self.free.setdefault(key, []).append(t)
        self.free.move_to_end(key)
        self.free_bytes += t.numel() * t.element_size()
        while self.free_bytes > self.cap_bytes and len(self.free) > 1:
            old_key = next(iter(self.free))
            if old_key == key:
                break
            for old in self.free.pop(old_key):
                self.free_bytes -= old.numel() * old.element_size()

    def end_capture(self):
        self.cap_free.clear()

    def clear(self):
        self.free.clear()
        self.cap_free.clear()
        self.pinned = []
        self.free_bytes = 0


POOL = _Pool()


class _Lease:
    """A set of pooled tensors that goes back to the pool when released or garbage collected."""

    def __init__(self):
        self.t = {}

    def take(self, name, shape, dtype, device):
        self.t[name] = POOL.get(shape, dtype, device)
        return self.t[name]

    def release(self):
        for v in self.t.values():
            POOL.put(v)
        self.t = {}

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _fwd_buffers(dims, plan, prec, device, zin, ctx_packed, lease, keep_params):
    act = _act_dtype(prec)
    rows = plan.rows
    zout = torch.zeros((rows, dims.z_ld), dtype=torch.float32, device=device)
    log_s = torch.zeros((rows, dims.z_ld // 2), dtype=torch.float32, device=device)
    b = {
        "zin": zin, "ctx": ctx_packed, "zout": zout, "log_s": log_s,
        "zmid": lease.take("zmid", (rows, dims.z_ld), torch.float32, device),
        "z0": lease.take("z0", (rows, 128), act, device),
        "x": lease.take("x", (dims.n_layers + 1, rows, dims.n_ch), act, device),
        "r": lease.take("r", (rows, dims.n_layers * dims.n_ch), act, device),
        "params": lease.take("params", (rows, dims.z_ld), torch.float32, device) if keep_params else None,
    }
    s = FlowBuffers(ctx=_p(ctx_packed), zin=_p(zin), zmid=_p(b["zmid"]), zout=_p(zout), z0=_p(b["z0"]),
                    x=_p(b["x"]), r=_p(b["r"]), params=_p(b["params"]), log_s=_p(log_s))
    return b, s


def run_flowstep(dims, blob, plan, zin, ctx_packed, prec, inverse, lease, keep_params=False):
    """zin fp32 [rows][z_ld], ctx_packed act [rows][ctx_ld].  Returns (zout, log_s or None, buffers, struct)."""
    dev = zin.device
    bufs, s = _fwd_buffers(dims, plan, prec, dev, zin, ctx_packed, lease, keep_params)
    L = _lib.lib()
    fn = L.radtts_flowstep_inverse if inverse else L.radtts_flowstep_forward
    _lib.check(fn(ctypes.byref(dims), _lib.ptr(blob), plan.ptr, plan.B, plan.Tmax, ctypes.byref(s), prec,
                  ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
               "radtts_flowstep_inverse" if inverse else "radtts_flowstep_forward")
    return bufs["zout"], (None if inverse else bufs["log_s"]), bufs, s


_prep_streams = {}


def prep_stream(device):
    """The side stream the flow weight preparation runs on (a trainer queues its deferred flow-parameter update on it,
    ahead of the preparation that reads those parameters)."""
    side = _prep_streams.get(device.index)
    if side is None:
        side = _prep_streams[device.index] = torch.cuda.Stream(device=device)
    return side


def prepare_flows_async(dims_list, ws, n_per_flow, prec, want_backward, device, order=None):
    """radtts_flow_prepare of every flow of a stack on a side stream (forked from the current stream here, joined per
    flow by the returned events): ([blob_i], [event_i])."""
    side = prep_stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    blobs, events = [None] * len(dims_list), [None] * len(dims_list)
    with torch.cuda.stream(side):
        for i in (order if order is not None else range(len(dims_list))):
            blobs[i] = prepare_flow(dims_list[i], ws[i * n_per_flow:(i + 1) * n_per_flow], prec, want_backward, device)
            events[i] = torch.cuda.Event()
            events[i].record(side)
    return blobs, events


class FlowPrep:
    """Result of begin_flow_prep: the stack's weight tensors (autograd-linked), LUS log-dets and the prepared blobs."""
    __slots__ = ("ws", "n_per", "lus_ld", "dims_list", "prec", "prepared")


def begin_flow_prep(flows, z_ld, actives, prec=None):
    """Training direction: composes the LU-parameterised 1x1 matrices (batched, autograd-linked) and starts the weight
    preparation of ALL flows on a side stream.  Called at the very top of RADTTS.forward, it hides the preparation
    underneath the text encoder / attention / context-LSTM stages; flow_stack_packed(prep=...) picks the result up."""
    prec = current_precision() if prec is None else prec
    dev = flows[0].invtbl_conv.parameters().__next__().device
    fp = FlowPrep()
    fp.dims_list = [_flow_dims(f, z_ld, c) for f, c in zip(flows, actives)]
    fp.prec = prec
    lus_w = fp.lus_ld = None
    if len(flows) <= 16 and all(hasattr(f.invtbl_conv, "lower") for f in flows):
        lus_w, fp.lus_ld = lus_compose_stack([f.invtbl_conv for f in flows])
    ws = []
    for i, f in enumerate(flows):
        ws += _flow_weight_list(f, False, None if lus_w is None else lus_w[i])
    fp.ws, fp.n_per = ws, len(ws) // len(flows)
    with torch.no_grad():
        fp.prepared = prepare_flows_async(fp.dims_list, ws, fp.n_per, prec, True, dev)
    return fp


class _FlowStackFn(torch.autograd.Function):
    """A run of consecutive decoder flows on packed rows (the whole 8-flow stack in RADTTS.forward, a single flow
    for FlowStep.forward).  Inputs: zin [rows][z_ld] fp32, ctx [rows][ctx_ld] act, then the weight tensors of
    every flow as _flow_weight_list orders them (weight_v / weight_g go in as they are: weight norm and its backward run
    inside the library; the LUS composition of the 1x1 matrix is left to autograd).  Returns (zout, log_s_0, ...)."""

    @staticmethod
    def forward(ctx, zin, ctx_packed, plan, dims_list, prec, inverse, n_per_flow, sinks, prepared, *ws):
        need_bwd = (not inverse) and any(ctx.needs_input_grad)
        order = range(len(dims_list))
        if inverse:
            order = reversed(order)
        z = zin.contiguous()
        log_s_all = [None] * len(dims_list)
        saved = [None] * len(dims_list)
        main = torch.cuda.current_stream(z.device)
        if n_per_flow != 0 and prepared is None and len(dims_list) > 1 and not os.environ.get("RADTTS_NO_PREP_STREAM"):
            # weight norm + re-layout of flow i+1 (memory-bound) underneath the GEMMs of flow i (tensor-bound)
            prepared = prepare_flows_async(dims_list, ws, n_per_flow, prec, need_bwd, z.device, list(order))
        for i in order:
            dims = dims_list[i]
            if n_per_flow == 0:
                blob = ws[i]                                   # cached, already prepared (inference)
            elif prepared is not None:
                blob = prepared[0][i]
                main.wait_event(prepared[1][i])
            else:
                blob = prepare_flow(dims, ws[i * n_per_flow:(i + 1) * n_per_flow], prec, need_bwd, z.device)
            lease = _Lease()
            zout, log_s, bufs, _ = run_flowstep(dims, blob, plan, z, ctx_packed, prec, inverse, lease, need_bwd)
            if need_bwd:
                # zout / log_s leave this function as outputs and get this node as grad_fn: keep storage-sharing
                # aliases instead, or ctx -> output -> grad_fn -> ctx is a cycle only the cyclic GC can free
                bufs = dict(bufs, zout=zout.detach(), log_s=log_s.detach())
                saved[i] = (blob, bufs, lease, ws[i * n_per_flow:(i + 1) * n_per_flow],
                            None if sinks is None else sinks[i * n_per_flow:(i + 1) * n_per_flow])
            else:
                lease.release()
            log_s_all[i] = log_s
            z = zout
        if need_bwd:
            ctx.saved = (plan, dims_list, prec, n_per_flow, saved,
                         [None if w is None else tuple(w.shape) for w in ws], ctx_packed.dtype)
        if inverse:
            return z
        return (z,) + tuple(log_s_all)

    @staticmethod
    def backward(ctx, g_zout, *g_log_s):
        from . import ops_backward
        return ops_backward.flow_stack_backward(ctx.saved, g_zout, g_log_s)


def _ptr_array(tensors):
    return (ctypes.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


class _LUSStackFn(torch.autograd.Function):
    """W_k = P_k L_k U_k and log|det W_k| for ALL flows of a stack (Invertible1x1ConvLUS, reference common.py:407-428)
    in three launches, with the closed-form backward in two (csrc/lus.cu) -- instead of ~12 + ~30 tiny torch kernels
    per flow.  Inputs: (lower, upper, upper_diag) x n then the buffers (p, lower_diag) x n; outputs W x n, log_det x n."""

    @staticmethod
    def forward(ctx, n, *ts):
        lower, upper, ud = ts[0:3 * n:3], ts[1:3 * n:3], ts[2:3 * n:3]
        p, ldiag = ts[3 * n::2], ts[3 * n + 1::2]
        dev = lower[0].device
        _lib.require_cuda(*ts)
        cs = [int(t.shape[0]) for t in lower]
        keep = [[t.detach().float().contiguous() for t in grp] for grp in (lower, upper, ud, ldiag, p)]
        tmp = [torch.empty((c, c), dtype=torch.float32, device=dev) for c in cs]
        w = [torch.empty((c, c), dtype=torch.float32, device=dev) for c in cs]
        ld = torch.empty(n, dtype=torch.float32, device=dev)
        lds = [ld[k] for k in range(n)]
        _lib.check(_lib.lib().radtts_lus_compose(n, (ctypes.c_int * n)(*cs), _ptr_array(keep[0]), _ptr_array(keep[1]),
                                                 _ptr_array(keep[2]), _ptr_array(keep[3]), _ptr_array(keep[4]),
                                                 _ptr_array(tmp), _ptr_array(w), _ptr_array(lds),
                                                 _lib.stream_of(lower[0])), "radtts_lus_compose")
        ctx.n, ctx.cs = n, cs
        ctx.save_for_backward(*[t for grp in keep for t in grp])
        return tuple(w) + tuple(lds)

    @staticmethod
    def backward(ctx, *gs):
        n, cs = ctx.n, ctx.cs
        sv = ctx.saved_tensors
        lower, upper, ud, ldiag, p = (sv[i * n:(i + 1) * n] for i in range(5))
        dev = lower[0].device
        g_w, g_ld = list(gs[:n]), list(gs[n:])
        zeros = None
        g_w_c, g_w_ld = [], []
        for k in range(n):
            g = g_w[k]
            if g is None:
                g = torch.zeros((cs[k], cs[k]), dtype=torch.float32, device=dev)
            elif g.dtype != torch.float32 or g.stride(1) != 1:
                g = g.float().contiguous()
            g_w_c.append(g)
            g_w_ld.append(int(g.stride(0)))
        g_ld_c = [None if g is None else g.float().contiguous() for g in g_ld]
        tmp = [torch.empty((c, c), dtype=torch.float32, device=dev) for c in cs]
        gl = [torch.empty((c, c), dtype=torch.float32, device=dev) for c in cs]
        gu = [torch.empty((c, c), dtype=torch.float32, device=dev) for c in cs]
        gd = [torch.empty(c, dtype=torch.float32, device=dev) for c in cs]
        _lib.check(_lib.lib().radtts_lus_backward(n, (ctypes.c_int * n)(*cs), _ptr_array(lower), _ptr_array(upper),
                                                  _ptr_array(ud), _ptr_array(ldiag), _ptr_array(p), _ptr_array(g_w_c),
                                                  (ctypes.c_int * n)(*g_w_ld), _ptr_array(g_ld_c), _ptr_array(tmp),
                                                  _ptr_array(gl), _ptr_array(gu), _ptr_array(gd),
                                                  _lib.stream_of(lower[0])), "radtts_lus_backward")
        out = [None]
        for k in range(n):
            out += [gl[k], gu[k], gd[k]]
        return tuple(out) + (None,) * (2 * n)


def lus_compose_stack(convs):
    """[Invertible1x1ConvLUS] -> ([W_k], [log_det_k]) with autograd links to lower / upper / upper_diag."""
    n = len(convs)
    ts = []
    for c in convs:
        ts += [c.lower, c.upper, c.upper_diag]
    for c in convs:
        ts += [c.p, c.lower_diag]
    out = _LUSStackFn.apply(n, *ts)
    return list(out[:n]), list(out[n:])


def ctx_ld_of(n_ctx):
    return (n_ctx + 63) // 64 * 64


def _cached_blob(flow, dims, prec, inverse, device):
    """Prepared-weight blob of a frozen flow (inference).  The cache lives ON the module (so it dies with it -- a global
    dict keyed by id(flow) can hand a new model the blob of a garbage-collected one) and is invalidated by the
    parameters' version counters."""
    cache = flow.__dict__.setdefault("_radtts_b200_blobs", {})
    key = (prec, bool(inverse), dims.c_off, device.index)
    versions = tuple((id(p), p._version) for p in flow.parameters())
    hit = cache.get(key)
    if hit is not None and hit[0] == versions:
        return hit[1]
    with torch.no_grad():
        blob = prepare_flow(dims, _flow_weight_list(flow, inverse), prec, False, device)
    cache[key] = (versions, blob)
    return blob


def flow_stack_packed(flows, zin, ctx_packed, plan, z_ld, actives, inverse=False, prec=None, prep=None):
    """Runs `flows` (in order; reversed when inverse) on packed rows.  actives[i] = channels flow i transforms.
    prep: a FlowPrep from begin_flow_prep (training direction, same flows / precision), or None."""
    prec = current_precision() if prec is None else prec
    dims_list = [_flow_dims(f, z_ld, c) for f, c in zip(flows, actives)]
    lus_ld = None
    if not torch.is_grad_enabled():
        # inference: weights are frozen between calls -> the re-laid-out blobs are cached per flow (keyed on the
        # parameters' version counters), and neither weight norm nor W^-1 nor the re-layout run again
        blobs = [_cached_blob(f, d, prec, inverse, zin.device) for f, d in zip(flows, dims_list)]
        out = _FlowStackFn.apply(zin, ctx_packed, plan, dims_list, prec, inverse, 0, None, None, *blobs)
    elif prep is not None and not inverse and prep.prec == prec and len(prep.dims_list) == len(flows):
        lus_ld = prep.lus_ld
        sinks = [w if (_direct_grad and isinstance(w, torch.nn.Parameter)) else None for w in prep.ws] if _direct_grad else None
        out = _FlowStackFn.apply(zin, ctx_packed, plan, dims_list, prec, inverse, prep.n_per, sinks, prep.prepared, *prep.ws)
    else:
        # training direction with LU-parameterised 1x1 convs: W = P L U and log|det| of the whole stack in one batched
        # Function (csrc/lus.cu) instead of ~40 tiny torch kernels per flow
        lus_w = None
        if not inverse and len(flows) <= 16 and zin.is_cuda and all(hasattr(f.invtbl_conv, "lower") for f in flows):
            lus_w, lus_ld = lus_compose_stack([f.invtbl_conv for f in flows])
        ws = []
        for i, f in enumerate(flows):
            ws += _flow_weight_list(f, inverse, None if lus_w is None else lus_w[i])
        n_per = len(ws) // len(flows)
        # gradient sinks: with direct accumulation on, the weight-norm backward kernel adds grad_v / grad_g straight
        # into the parameters' existing .grad (e.g. views of the optimizer's flat buffer) instead of handing 144 tensors
        # to autograd's AccumulateGrad
        sinks = [w if (_direct_grad and isinstance(w, torch.nn.Parameter)) else None for w in ws] if _direct_grad else None
        out = _FlowStackFn.apply(zin, ctx_packed, plan, dims_list, prec, inverse, n_per, sinks, None, *ws)
    if inverse:
        return out
    zout, log_s = out[0], list(out[1:])
    if lus_ld is not None:
        return zout, lus_ld, log_s
    log_dets = []
    for f in flows:
        inv = f.invtbl_conv
        log_dets.append(inv.log_det() if hasattr(inv, "log_det")
                        else torch.logdet(inv.conv.weight.squeeze(-1)).clone())
    return zout, log_dets, log_s


def flow_step_packed(flow, zin, ctx_packed, plan, z_ld, c_active, inverse=False, prec=None):
    out = flow_stack_packed([flow], zin, ctx_packed, plan, z_ld, [c_active], inverse, prec)
    if inverse:
        return out
    return out[0], out[1][0], out[2][0]


def _is_fused_flow(flow):
    wn = getattr(flow.affine_tfn, "affine_param_predictor", None)
    return flow.affine_tfn.affine_model == "wavenet" and wn is not None and wn.affine_activation == "softplus"


def flow_step(flow, z, context, inverse=False, seq_lens=None):
    """FlowStep.forward on reference-shaped tensors: z (B,C,T'), context (B,n_ctx,T')."""
    _lib.require_cuda(z, context)
    if not _is_fused_flow(flow):
        raise NotImplementedError("decoder FlowStep supports affine_model='wavenet' with softplus activations")
    B, C, T = z.shape
    if seq_lens is None:
        seq_lens = torch.full((B,), T, dtype=torch.int64, device=z.device)
    prec = current_precision()
    plan = FramePlan(seq_lens.to(z.device), 1, T)
    z_ld = (C + 15) // 16 * 16
    n_ctx = context.shape[1]
    zin = pack(z, plan, 1, torch.float32, z_ld, z_ld - C, C)
    ctxp = pack(context, plan, 1, _act_dtype(prec), ctx_ld_of(n_ctx), 0, ctx_ld_of(n_ctx))
    if inverse:
        zout = flow_step_packed(flow, zin, ctxp, plan, z_ld, C, True, prec)
        return unpack(zout, plan, C, 1, z_ld - C)
    zout, log_det, log_s = flow_step_packed(flow, zin, ctxp, plan, z_ld, C, False, prec)
    return unpack(zout, plan, C, 1, z_ld - C), log_det, unpack(log_s, plan, C // 2, 1, 0)


def begin_decoder_prep(model):
    """RADTTS.forward calls this first (training on CUDA): see begin_flow_prep."""
    flows = list(model.flows)
    if not flows or not all(_is_fused_flow(f) for f in flows):
        return None
    z_ld = model.n_mel_channels * model.n_group_size
    return begin_flow_prep(flows, z_ld, _active_channels(model, z_ld))


def _active_channels(model, z_ld):
    actives, c = [], z_ld
    for i in range(len(model.flows)):
        if i in model.exit_steps:
            c -= model.n_early_size
        actives.append(c)
    return actives


def decoder_forward(model, mel, context, out_lens, prep=None):
    """Training direction of the decoder loop (reference radtts.py:414,431-444) on packed frames: one pack,
    n_flows fused flow steps operating in place on the column suffix that is still active, one unpack."""
    _lib.require_cuda(mel, context)
    g = model.n_group_size
    prec = current_precision()
    z_ld = model.n_mel_channels * g
    plan = FramePlan(out_lens.to(mel.device), g, mel.shape[2] // g)
    z = pack(mel, plan, g, torch.float32, z_ld, 0, z_ld)
    n_ctx = context.shape[1]
    ctxp = pack(context, plan, 1, _act_dtype(prec), ctx_ld_of(n_ctx), 0, ctx_ld_of(n_ctx))
    actives = _active_channels(model, z_ld)
    z, log_det_list, log_s = flow_stack_packed(list(model.flows), z, ctxp, plan, z_ld, actives, False, prec, prep)
    log_s_list = [unpack(ls, plan, c // 2, 1, 0) for ls, c in zip(log_s, actives)]
    z_mel = unpack(z, plan, z_ld, 1, 0)
    # the packed originals ride along (rows of gaps / beyond the lengths are exactly zero in all of them), so that the
    # flow loss can sum them directly instead of masking and reducing the padded (B, C, T') copies -- and its gradient
    # then enters the flow stack's backward without the unpack / pack round trip (loss.RADTTSLoss)
    z_mel._rb_packed = (z, list(log_s), list(log_s_list), out_lens)
    return z_mel, log_det_list, log_s_list


def decoder_inverse(model, residual, context, out_lens):
    """Sampling direction (reference radtts.py:656-677): residual (B, 80*g, T') -> mel (B, 80, T'*g)."""
    _lib.require_cuda(residual, context)
    g = model.n_group_size
    prec = current_precision()
    z_ld = residual.shape[1]
    plan = FramePlan(out_lens.to(residual.device), g, residual.shape[2])
    z = pack(residual, plan, 1, torch.float32, z_ld, 0, z_ld)
    n_ctx = context.shape[1]
    ctxp = pack(context, plan, 1, _act_dtype(prec), ctx_ld_of(n_ctx), 0, ctx_ld_of(n_ctx))
    z = flow_stack_packed(list(model.flows), z, ctxp, plan, z_ld, _active_channels(model, z_ld), True, prec)
    return unpack(z, plan, z_ld // g, g, 0)


# ------------------------------------------------------------------------------------------------------
# remaining hot-path entry points (filled in as their kernels land)
# ------------------------------------------------------------------------------------------------------
def _needs_grad(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


# Training direction of the BGAP attribute flows: the parameter networks run on the packed-row conv kernels in
# both directions (_ConvStackFn), the spline / affine transforms and the small 1x1 convs have closed-form backward kernels.
class _SplineForwardFn(torch.autograd.Function):
    """Forward-direction spline coupling apply (radtts_rqspline_apply) with the closed-form backward kernel
    (radtts_rqspline_backward): z (B, C, T), params (B, h (2 nb + 1), T) -> (y, log_s)."""

    @staticmethod
    def forward(ctx, z, params, n_bins, left, right, bottom, top):
        z = z.float().contiguous()
        params = params.float().contiguous()
        B, C, T = z.shape
        y = torch.empty_like(z)
        log_s = torch.empty((B, 1, T), dtype=torch.float32, device=z.device)
        _lib.check(_lib.lib().radtts_rqspline_apply(_lib.ptr(z), _lib.ptr(params), B, C, T, n_bins, 0,
                                                    ctypes.c_float(left), ctypes.c_float(right), ctypes.c_float(bottom),
                                                    ctypes.c_float(top), _lib.ptr(y), _lib.ptr(log_s), _lib.stream_of(z)),
                   "radtts_rqspline_apply")
        ctx.save_for_backward(z, params)
        ctx.consts = (n_bins, left, right, bottom, top)
        return y, log_s

    @staticmethod
    def backward(ctx, g_y, g_log_s):
        z, params = ctx.saved_tensors
        n_bins, left, right, bottom, top = ctx.consts
        B, C, T = z.shape
        g_y = None if g_y is None else g_y.float().contiguous()
        g_log_s = None if g_log_s is None else g_log_s.float().contiguous()
        g_z = torch.empty_like(z)
        g_p = torch.empty_like(params)
        _lib.check(_lib.lib().radtts_rqspline_backward(_lib.ptr(z), _lib.ptr(params), _lib.ptr(g_y), _lib.ptr(g_log_s), B, C,
                                                       T, n_bins, ctypes.c_float(left), ctypes.c_float(right),
                                                       ctypes.c_float(bottom), ctypes.c_float(top), _lib.ptr(g_z),
                                                       _lib.ptr(g_p), _lib.stream_of(z)), "radtts_rqspline_backward")
        return g_z, g_p, None, None, None, None, None


class _AffineForwardFn(torch.autograd.Function):
    """Forward-direction affine coupling apply (radtts_affine_apply) with its closed-form backward kernel
    (radtts_affine_backward): z (B, C, T), params (B, C, T) = [raw scale | translation] -> (y, log_s (B, C/2, T))."""

    @staticmethod
    def forward(ctx, z, params, scaling):
        z = z.float().contiguous()
        params = params.float().contiguous()
        B, C, T = z.shape
        y = torch.empty_like(z)
        log_s = torch.empty((B, C // 2, T), dtype=torch.float32, device=z.device)
        _lib.check(_lib.lib().radtts_affine_apply(_lib.ptr(z), _lib.ptr(params), B, C, T, scaling, 0, _lib.ptr(y),
                                                  _lib.ptr(log_s), _lib.stream_of(z)), "radtts_affine_apply")
        ctx.save_for_backward(z, params)
        ctx.scaling = scaling
        return y, log_s

    @staticmethod
    def backward(ctx, g_y, g_log_s):
        z, params = ctx.saved_tensors
        B, C, T = z.shape
        g_y = None if g_y is None else g_y.float().contiguous()
        g_log_s = None if g_log_s is None else g_log_s.float().contiguous()
        g_z = torch.empty_like(z)
        g_p = torch.empty_like(params)
        _lib.check(_lib.lib().radtts_affine_backward(_lib.ptr(z), _lib.ptr(params), _lib.ptr(g_y), _lib.ptr(g_log_s), B, C, T,
                                                     ctx.scaling, _lib.ptr(g_z), _lib.ptr(g_p), _lib.stream_of(z)),
                   "radtts_affine_backward")
        return g_z, g_p, None


def _training_direction_only(what):
    raise NotImplementedError("%s: gradients are implemented for the training (forward) direction only; run the sampling "
                              "direction under torch.no_grad()" % what)


class _PointwiseSmallFn(torch.autograd.Function):
    """y[b,:,t] = W x[b,:,t] for the small dense 1x1 convs of the attribute flows (C <= 16), forward and backward kernels."""

    @staticmethod
    def forward(ctx, x, w):
        x = x.float().contiguous()
        w = w.float().contiguous()
        B, C, T = x.shape
        y = torch.empty_like(x)
        _lib.check(_lib.lib().radtts_pointwise_conv_small(_lib.ptr(x), _lib.ptr(w), B, C, T, _lib.ptr(y), _lib.stream_of(x)),
                   "radtts_pointwise_conv_small")
        ctx.save_for_backward(x, w)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, w = ctx.saved_tensors
        B, C, T = x.shape
        g_y = g_y.float().contiguous()
        g_x = torch.empty_like(x)
        g_w = torch.empty_like(w)
        _lib.check(_lib.lib().radtts_pointwise_conv_small_backward(_lib.ptr(x), _lib.ptr(w), _lib.ptr(g_y), B, C, T,
                                                                   _lib.ptr(g_x), _lib.ptr(g_w), _lib.stream_of(x)),
                   "radtts_pointwise_conv_small_backward")
        return g_x, g_w


def pointwise_conv(z, w):
    """y[b,:,t] = W z[b,:,t] (Invertible1x1Conv / Invertible1x1ConvLUS module API on reference-shaped tensors)."""
    _lib.require_cuda(z, w)
    B, C, T = z.shape
    if C > 16:
        # large-C 1x1 convs on the hot path go through ops.flow_step; this module-level call is a plain library GEMM
        return torch.matmul(w.float(), z.float())
    if _needs_grad(z, w):
        return _PointwiseSmallFn.apply(z, w)
    z = z.float().contiguous()
    w = w.detach().float().contiguous()
    y = torch.empty_like(z)
    _lib.check(_lib.lib().radtts_pointwise_conv_small(_lib.ptr(z), _lib.ptr(w), B, C, T, _lib.ptr(y), _lib.stream_of(z)),
               "radtts_pointwise_conv_small")
    return y


def wn_forward(wn, z, context, seq_lens):
    raise NotImplementedError("WN runs fused inside FlowStep (ops.flow_step); standalone WN.forward is not exposed")


def _prepared_conv(conv, c_in_pad, prec):
    """Re-laid-out weight blob of a Conv1d, cached ON the module per parameter version (inference weights are frozen).
    Not a global dict keyed by id(conv): ids are reused after garbage collection and equal version counters would then
    hand a new model another model's weights."""
    w, b = conv.weight, conv.bias
    cache = conv.__dict__.setdefault("_radtts_b200_blobs", {})
    slot = (c_in_pad, prec, w.device.index)
    key = (id(w), w._version, None if b is None else (id(b), b._version))
    hit = cache.get(slot)
    if hit is not None and hit[0] == key:
        return hit[1]
    L = _lib.lib()
    c_out, c_in, k = w.shape
    nbytes = int(L.radtts_conv_prepared_bytes(c_out, c_in_pad, k, prec))
    blob = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    wf = w.detach().float().contiguous()
    bf = None if b is None else b.detach().float().contiguous()
    _lib.check(L.radtts_conv_prepare(_lib.ptr(wf), _lib.ptr(bf), c_out, c_in, c_in_pad, k, prec, _lib.ptr(blob),
                                     ctypes.c_size_t(nbytes), _lib.stream_of(w)), "radtts_conv_prepare")
    cache[slot] = (key, blob)
    return blob


def _pad64(n):
    return (n + 63) // 64 * 64


_ACT_CODES = {"none": 0, "softplus": 1, "relu": 2}


class ConvStackSpec:
    """Static description of a stack of ConvNorm layers on packed rows: per layer (ksize, dilation, act code, partial,
    mask_rows), whether every utterance physically keeps all T rows (plain, non-partial stacks read the padded region,
    reference common.py:145-154) and whether the packed input is zeroed past the valid length (partial convs)."""

    def __init__(self, layers, geom_full, valid_only):
        self.layers, self.geom_full, self.valid_only = tuple(layers), bool(geom_full), bool(valid_only)


def _prepare_conv_raw(w, b, c_in_pad, prec):
    L = _lib.lib()
    c_out, c_in, k = w.shape
    nbytes = int(L.radtts_conv_prepared_bytes(c_out, c_in_pad, k, prec))
    blob = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    wf = w.detach().float().contiguous()
    bf = None if b is None else b.detach().float().contiguous()
    _lib.check(L.radtts_conv_prepare(_lib.ptr(wf), _lib.ptr(bf), c_out, c_in, c_in_pad, k, prec, _lib.ptr(blob),
                                     ctypes.c_size_t(nbytes), _lib.stream_of(w)), "radtts_conv_prepare")
    return blob


class _ConvStackFn(torch.autograd.Function):
    """x (B, C, T) -> (B, C_out, T) through a stack of conv layers, every layer ONE row-GEMM launch with a fused
    bias / partial-conv ratio / activation epilogue (radtts_conv_rows) -- and the backward on the same engines
    (radtts_conv_rows_backward: transposed-tap dgrad GEMM, one tcgen05 weight-gradient problem per tap with K = packed
    rows, streaming bias column sums).  params = (w_0, b_0, w_1, b_1, ...) in the reference's (c_out, c_in, k) layout."""

    @staticmethod
    def forward(ctx, x, seq_lens, spec, prec, *params):
        B, C, T = x.shape
        dev = x.device
        act = _act_dtype(prec)
        lens = seq_lens if seq_lens is not None else torch.full((B,), T, dtype=torch.int64, device=dev)
        geom = torch.full((B,), T, dtype=torch.int64, device=dev) if spec.geom_full else None
        plan = FramePlan(lens.to(dev), 1, T, geom)
        L = _lib.lib()
        stream = _lib.stream_of(x)
        lease = _Lease()
        cur = pack_frames(x, plan, 1, act, _pad64(C), 0, _pad64(C), valid_only=spec.valid_only)
        acts = [cur]
        for i, (k, dil, act_code, partial, mask_rows) in enumerate(spec.layers):
            w, b = params[2 * i], params[2 * i + 1]
            c_out, c_in, kk = w.shape
            cur_w = acts[-1].shape[1]
            blob = _prepare_conv_raw(w, b, cur_w, prec)
            out_w = _pad64(c_out)
            out = lease.take("y%d" % i, (plan.rows, out_w), act, dev)
            _lib.check(L.radtts_conv_rows(_lib.ptr(blob), c_out, cur_w, kk, dil, act_code, partial, mask_rows,
                                          _lib.ptr(acts[-1]), cur_w, _lib.ptr(out), out_w, 0, plan.ptr, plan.B, plan.Tmax,
                                          prec, stream), "radtts_conv_rows")
            acts.append(out)
        c_last = params[2 * (len(spec.layers) - 1)].shape[0]
        last = acts[-1]
        res = unpack_frames(last.float() if last.dtype != torch.float32 else last, plan, c_last, 1, 0)
        if any(ctx.needs_input_grad):
            ctx.saved = (plan, spec, prec, acts, lease, params, (B, C, T))
        else:
            lease.release()
        return res

    @staticmethod
    def backward(ctx, g):
        plan, spec, prec, acts, lease, params, (B, C, T) = ctx.saved
        ctx.saved = None
        dev = g.device
        act = _act_dtype(prec)
        L = _lib.lib()
        stream = _lib.stream_of(g)
        scratch = _Lease()
        n_layers = len(spec.layers)
        c_last = params[2 * (n_layers - 1)].shape[0]
        g_y = pack_frames(g.contiguous(), plan, 1, act, _pad64(c_last), 0, _pad64(c_last))
        grads = [None] * len(params)
        need_x = ctx.needs_input_grad[0]
        for i in reversed(range(n_layers)):
            k, dil, act_code, partial, mask_rows = spec.layers[i]
            w, b = params[2 * i], params[2 * i + 1]
            c_out, c_in, kk = w.shape
            x_i, y_i = acts[i], acts[i + 1]
            c_in_pad = x_i.shape[1]
            nbytes = int(L.radtts_conv_backward_prepared_bytes(c_out, c_in_pad, kk, prec))
            blob_t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            wf = w.detach().float().contiguous()
            _lib.check(L.radtts_conv_prepare_backward(_lib.ptr(wf), c_out, c_in, c_in_pad, kk, prec, _lib.ptr(blob_t),
                                                      ctypes.c_size_t(nbytes), stream), "radtts_conv_prepare_backward")
            g_pre = scratch.take("g_pre%d" % i, (plan.rows, _pad64(c_out)), act, dev)
            g_x = scratch.take("g_x%d" % i, (plan.rows, c_in_pad), act, dev) if (i > 0 or need_x) else None
            g_w = torch.empty((c_out, c_in, kk), dtype=torch.float32, device=dev)
            g_b = None if b is None else torch.empty((c_out + 15) // 16 * 16, dtype=torch.float32, device=dev)
            _lib.check(L.radtts_conv_rows_backward(_lib.ptr(blob_t), c_out, c_in, c_in_pad, kk, dil, act_code, partial,
                                                   mask_rows, _lib.ptr(x_i), c_in_pad, _lib.ptr(y_i), y_i.shape[1], 0,
                                                   _lib.ptr(g_y), g_y.shape[1], _lib.ptr(g_pre), _lib.ptr(g_x), c_in_pad,
                                                   _lib.ptr(g_w), _lib.ptr(g_b), plan.ptr, plan.B, plan.Tmax, prec, stream),
                       "radtts_conv_rows_backward")
            grads[2 * i] = g_w.to(w.dtype)
            if b is not None:
                grads[2 * i + 1] = g_b[:c_out].to(b.dtype)
            g_y = g_x
        g_in = None
        if need_x:
            gx = g_y.float() if g_y.dtype != torch.float32 else g_y
            g_in = unpack_frames(gx.contiguous(), plan, C, 1, 0)
            if g_in.shape[2] != T:
                g_in = torch.nn.functional.pad(g_in, (0, T - g_in.shape[2]))
        scratch.release()
        lease.release()
        return (g_in, None, None, None) + tuple(grads)


def conv_stack(x, seq_lens, spec, params, prec=None):
    prec = current_precision() if prec is None else prec
    return _ConvStackFn.apply(x.float(), seq_lens, spec, prec, *params)


def _simple_conv_net_spec(net):
    layers = [(layer.conv.kernel_size[0], layer.dilation, _ACT_CODES["relu"], int(net.use_partial_padding), 1)
              for layer in net.layers]
    layers.append((1, 1, _ACT_CODES["none"], 0, 0))
    # The reference runs these nets over the whole zero-padded batch: plain (non-partial) ConvNorm layers read the padded
    # region in their first layer and `last_layer` is not masked (common.py:145-154,514) -> every utterance keeps all T rows
    return ConvStackSpec(layers, geom_full=True, valid_only=bool(net.use_partial_padding))


def simple_conv_net(net, x, seq_lens):
    """SimpleConvNet.forward (reference common.py:503-515) on packed rows: n_layers x [ConvNorm (+partial padding) ->
    ReLU] then the 1x1 `last_layer`; every layer is one row-GEMM launch with a fused epilogue, in both directions."""
    _lib.require_cuda(x)
    params = []
    for layer in net.layers:
        params += [layer.conv.weight, layer.conv.bias]
    params += [net.last_layer.weight, net.last_layer.bias]
    if not _needs_grad(x, *params):
        return _simple_conv_net_cached(net, x, seq_lens)
    return conv_stack(x, seq_lens, _simple_conv_net_spec(net), params)


def _simple_conv_net_cached(net, x, seq_lens):
    """Inference: as conv_stack's forward, with the re-laid-out weights cached on the modules (frozen weights)."""
    B, C, T = x.shape
    prec = current_precision()
    act = _act_dtype(prec)
    dev = x.device
    if seq_lens is None:
        seq_lens = torch.full((B,), T, dtype=torch.int64, device=dev)
    geom = torch.full((B,), T, dtype=torch.int64, device=dev)
    plan = FramePlan(seq_lens.to(dev), 1, T, geom)
    L = _lib.lib()
    stream = _lib.stream_of(x)
    lease = _Lease()
    cur = pack_frames(x, plan, 1, act, _pad64(C), 0, _pad64(C), valid_only=bool(net.use_partial_padding))
    cur_w = _pad64(C)
    convs = [(layer.conv, layer.dilation, 2, int(net.use_partial_padding)) for layer in net.layers]
    convs.append((net.last_layer, 1, 0, 0))
    for i, (conv, dil, act_code, partial) in enumerate(convs):
        mask_rows = 0 if i == len(convs) - 1 else 1
        c_out, c_in, k = conv.weight.shape
        blob = _prepared_conv(conv, cur_w, prec)
        out_w = _pad64(c_out)
        out = lease.take("y%d" % i, (plan.rows, out_w), act, dev)
        _lib.check(L.radtts_conv_rows(_lib.ptr(blob), c_out, cur_w, k, dil, act_code, partial, mask_rows, _lib.ptr(cur), cur_w,
                                      _lib.ptr(out), out_w, 0, plan.ptr, plan.B, plan.Tmax, prec, stream),
                   "radtts_conv_rows")
        cur, cur_w = out, out_w
    res = unpack_frames(cur.float() if cur.dtype != torch.float32 else cur, plan, convs[-1][0].weight.shape[0], 1, 0)
    lease.release()
    return res


def affine_coupling(layer, z, context, inverse, seq_lens):
    """AffineTransformationLayer.forward (reference common.py:810-832) on reference-shaped tensors."""
    if layer.affine_model == "wavenet":
        raise NotImplementedError("the WN-based coupling runs fused inside FlowStep (ops.flow_step)")
    _lib.require_cuda(z, context)
    scaling = layer.scaling_fn
    if isinstance(scaling, list) or scaling not in _SCALING:
        raise NotImplementedError("per-channel scaling_fn lists are not supported")
    z = z.float().contiguous()
    B, C, T = z.shape
    h = C // 2
    params = simple_conv_net(layer.affine_param_predictor, torch.cat((z[:, :h], context.float()), 1), seq_lens)
    if _needs_grad(z, params):
        if inverse:
            _training_direction_only("affine coupling")
        # the parameter net (conv_stack) reaches z[:, :h] and the context through autograd; the Function adds the direct path
        return _AffineForwardFn.apply(z, params, _SCALING[scaling])
    y = torch.empty_like(z)
    log_s = None if inverse else torch.empty((B, h, T), dtype=torch.float32, device=z.device)
    _lib.check(_lib.lib().radtts_affine_apply(_lib.ptr(z), _lib.ptr(params.contiguous()), B, C, T, _SCALING[scaling],
                                              int(bool(inverse)), _lib.ptr(y), _lib.ptr(log_s), _lib.stream_of(z)),
               "radtts_affine_apply")
    return y if inverse else (y, log_s)


def spline_coupling(layer, z, context, inverse, seq_lens):
    """SplineTransformationLayer.forward with use_quadratic=True (reference common.py:694-743, splines.py:221-319)."""
    _lib.require_cuda(z, context)
    z = z.float().contiguous()
    B, C, T = z.shape
    h = layer.half_mel_channels
    params = simple_conv_net(layer.param_predictor, torch.cat((z[:, :h], context.float()), 1), seq_lens).contiguous()
    n_bins = layer.n_bins // 2
    if _needs_grad(z, params):
        if inverse:
            _training_direction_only("spline coupling")
        return _SplineForwardFn.apply(z, params, n_bins, float(layer.left), float(layer.right), float(layer.bottom),
                                      float(layer.top))
    y = torch.empty_like(z)
    log_s = None if inverse else torch.empty((B, 1, T), dtype=torch.float32, device=z.device)
    _lib.check(_lib.lib().radtts_rqspline_apply(_lib.ptr(z), _lib.ptr(params), B, C, T, n_bins, int(bool(inverse)),
                                                ctypes.c_float(layer.left), ctypes.c_float(layer.right),
                                                ctypes.c_float(layer.bottom), ctypes.c_float(layer.top), _lib.ptr(y),
                                                _lib.ptr(log_s), _lib.stream_of(z)), "radtts_rqspline_apply")
    return y if inverse else (y, log_s)


class _ConvAttnFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q_enc, k_enc, prior, key_lens, temp):
        q_enc, k_enc = q_enc.float().contiguous(), k_enc.float().contiguous()
        B, C, T1 = q_enc.shape
        T2 = k_enc.shape[2]
        dev = q_enc.device
        attn = torch.empty((B, 1, T1, T2), dtype=torch.float32, device=dev)
        logprob = torch.empty_like(attn)
        lse = torch.empty((B, T1), dtype=torch.float32, device=dev)
        prior_c = None if prior is None else prior.float().contiguous()
        lens = None if key_lens is None else key_lens.to(device=dev, dtype=torch.int64).contiguous()
        L = _lib.lib()
        _lib.check(L.radtts_convattn_forward(_lib.ptr(q_enc), _lib.ptr(k_enc), _lib.ptr(prior_c), _lib.ptr(lens), B, C,
                                             T1, T2, ctypes.c_float(temp), _lib.ptr(attn), _lib.ptr(logprob),
                                             _lib.ptr(lse), _lib.stream_of(q_enc)), "radtts_convattn_forward")
        ctx.save_for_backward(q_enc, k_enc, lse, attn)
        ctx.has_prior = prior is not None
        ctx.temp = temp
        return attn, logprob

    @staticmethod
    def backward(ctx, g_attn, g_logprob):
        q_enc, k_enc, lse, attn = ctx.saved_tensors
        B, C, T1 = q_enc.shape
        T2 = k_enc.shape[2]
        g_attn = None if g_attn is None else g_attn.float().contiguous()
        g_logprob = None if g_logprob is None else g_logprob.float().contiguous()
        gd = torch.empty((B, T1, T2), dtype=torch.float32, device=q_enc.device)
        g_q = torch.empty_like(q_enc)
        g_k = torch.empty_like(k_enc)
        L = _lib.lib()
        _lib.check(L.radtts_convattn_backward(_lib.ptr(q_enc), _lib.ptr(k_enc), _lib.ptr(lse), _lib.ptr(attn),
                                              _lib.ptr(g_attn), _lib.ptr(g_logprob), int(ctx.has_prior), B, C, T1, T2,
                                              ctypes.c_float(ctx.temp), _lib.ptr(gd), _lib.ptr(g_q), _lib.ptr(g_k),
                                              _lib.stream_of(q_enc)), "radtts_convattn_backward")
        return g_q, g_k, None, None, None


def _projection_stack(seq, x):
    """An nn.Sequential of ConvNorm / ReLU modules (ConvAttention.key_proj / query_proj, reference common.py:843-858)
    as one conv_stack call: plain zero-padded convs over the whole padded batch (no length masks, as in the reference)."""
    layers, params = [], []
    mods = list(seq)
    for i, m in enumerate(mods):
        if isinstance(m, torch.nn.ReLU):
            continue
        conv = m.conv
        if hasattr(conv, "weight_v") or conv.stride[0] != 1 or conv.padding[0] != conv.dilation[0] * (conv.kernel_size[0] - 1) // 2:
            return None
        relu = i + 1 < len(mods) and isinstance(mods[i + 1], torch.nn.ReLU)
        layers.append((conv.kernel_size[0], conv.dilation[0], _ACT_CODES["relu" if relu else "none"], 0, 0))
        params += [conv.weight, conv.bias]
    return conv_stack(x, None, ConvStackSpec(layers, geom_full=True, valid_only=False), params)


def conv_attention(att, queries, keys, mask, key_lens, attn_prior):
    """ConvAttention.forward (reference common.py:886-924), all of it in the library: the key / query projections are
    conv_stack calls (row GEMMs on the tcgen05 engine under autocast, exactly where the reference's convs would run in
    half precision; fp32 SIMT otherwise) with their backward on the same engines, and everything after them -- the part that
    dominates memory traffic in the reference -- is the fused distance / log-softmax / prior / softmax kernel."""
    _lib.require_cuda(queries, keys)
    k_enc = _projection_stack(att.key_proj, keys)
    q_enc = _projection_stack(att.query_proj, queries)
    if k_enc is None or q_enc is None:    # a projection this build does not recognise (weight norm, strides): library convs
        k_enc = att.key_proj(keys).float()
        q_enc = att.query_proj(queries).float()
    if key_lens is None and mask is not None:
        key_lens = (~mask.squeeze(-1)).sum(1)
    if mask is None:
        key_lens = None
    return _ConvAttnFn.apply(q_enc, k_enc, attn_prior, key_lens, 0.0005)


class _EncNormFn(torch.autograd.Function):
    """Partial-conv renormalisation + masked InstanceNorm1d(affine) + ReLU + dropout + length mask of one text-encoder conv
    block (reference common.py:348-356), fused: radtts_encnorm_forward / _backward."""

    @staticmethod
    def forward(ctx, raw, conv_bias, lens, gamma, beta, drop, drop_scale, ksize, eps):
        raw = raw.float().contiguous()
        B, C, T = raw.shape
        dev = raw.device
        out = torch.empty_like(raw)
        mean = torch.empty((B, C), dtype=torch.float32, device=dev)
        rstd = torch.empty((B, C), dtype=torch.float32, device=dev)
        lens = lens.to(device=dev, dtype=torch.int64).contiguous()
        cb = None if conv_bias is None else conv_bias.detach().float().contiguous()
        ga = None if gamma is None else gamma.detach().float().contiguous()
        be = None if beta is None else beta.detach().float().contiguous()
        _lib.check(_lib.lib().radtts_encnorm_forward(_lib.ptr(raw), _lib.ptr(cb), _lib.ptr(lens), _lib.ptr(ga), _lib.ptr(be),
                                                     _lib.ptr(drop), ctypes.c_float(drop_scale), B, C, T, ksize,
                                                     ctypes.c_float(eps), _lib.ptr(out), _lib.ptr(mean), _lib.ptr(rstd),
                                                     _lib.stream_of(raw)), "radtts_encnorm_forward")
        ctx.save_for_backward(raw, lens, mean, rstd)
        ctx.aux = (cb, ga, be, drop, drop_scale, ksize, conv_bias is not None, gamma is not None, beta is not None)
        return out

    @staticmethod
    def backward(ctx, g_out):
        raw, lens, mean, rstd = ctx.saved_tensors
        cb, ga, be, drop, drop_scale, ksize, has_cb, has_ga, has_be = ctx.aux
        B, C, T = raw.shape
        dev = raw.device
        g_out = g_out.float().contiguous()
        g_raw = torch.empty_like(raw)
        g_ga = torch.empty(C, dtype=torch.float32, device=dev) if has_ga else None
        g_be = torch.empty(C, dtype=torch.float32, device=dev) if has_be else None
        g_cb = torch.empty(C, dtype=torch.float32, device=dev) if has_cb else None
        _lib.check(_lib.lib().radtts_encnorm_backward(_lib.ptr(raw), _lib.ptr(cb), _lib.ptr(lens), _lib.ptr(ga), _lib.ptr(be),
                                                      _lib.ptr(drop), ctypes.c_float(drop_scale), _lib.ptr(mean),
                                                      _lib.ptr(rstd), _lib.ptr(g_out), B, C, T, ksize, _lib.ptr(g_raw),
                                                      _lib.ptr(g_ga), _lib.ptr(g_be), _lib.ptr(g_cb), _lib.stream_of(raw)),
                   "radtts_encnorm_backward")
        return g_raw, g_cb, None, g_ga, g_be, None, None, None, None


def encoder_conv_block(conv, norm, x_masked, lens, p_drop, training):
    """One block of the text encoder's conv stack on the whole padded batch (reference common.py:348-356): the k-tap
    convolution is a library call, everything after it one fused kernel.  x_masked must be zero beyond each length."""
    raw = torch.nn.functional.conv1d(x_masked, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation)
    drop, scale = None, 1.0
    if training and p_drop > 0:
        drop = torch.empty_like(raw).bernoulli_(1.0 - p_drop)
        scale = 1.0 / (1.0 - p_drop)
    return _EncNormFn.apply(raw, conv.bias, lens, norm.weight, norm.bias, drop, scale, conv.kernel_size[0], norm.eps)


class _AttnCTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, attn_logprob, in_lens, out_lens, blank_logprob):
        x = attn_logprob.float().contiguous()
        B, _, T1, T2 = x.shape
        dev = x.device
        il = in_lens.to(device=dev, dtype=torch.int64).contiguous()
        ol = out_lens.to(device=dev, dtype=torch.int64).contiguous()
        losses = torch.empty(B, dtype=torch.float32, device=dev)
        grad = torch.empty_like(x)
        L = _lib.lib()
        nws = int(L.radtts_attn_ctc_workspace_bytes(B, T1, T2))
        ws = _lib.scratch(dev, nws)
        _lib.check(L.radtts_attn_ctc(_lib.ptr(x), _lib.ptr(il), _lib.ptr(ol), B, T1, T2, ctypes.c_float(blank_logprob),
                                     _lib.ptr(losses), _lib.ptr(grad), _lib.ptr(ws), ctypes.c_size_t(ws.numel()),
                                     _lib.stream_of(x)), "radtts_attn_ctc")
        ctx.save_for_backward(grad)
        return (losses / il.clamp(min=1).to(losses.dtype)).mean()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None


def attention_ctc_loss(attn_logprob, in_lens, out_lens, blank_logprob=-1.0):
    """AttentionCTCLoss.forward (reference loss.py:118-135) as one fused CUDA launch, no host synchronisation."""
    _lib.require_cuda(attn_logprob)
    return _AttnCTCFn.apply(attn_logprob, in_lens, out_lens, float(blank_logprob))


class _ContextGatherFn(torch.autograd.Function):
    """context = bmm(text_enc, attn_hard^T) for a binarized alignment (reference radtts.py:399), as a gather by the
    frame -> token index of the MAS kernel; backward = per-token segment sum.  No gradient reaches the alignment."""

    @staticmethod
    def forward(ctx, text_enc, f2t):
        text = text_enc.float().contiguous()
        B, C, T2 = text.shape
        T1 = f2t.shape[1]
        out = torch.empty((B, C, T1), dtype=torch.float32, device=text.device)
        _lib.check(_lib.lib().radtts_context_gather(_lib.ptr(text), _lib.ptr(f2t), B, C, T1, T2, _lib.ptr(out),
                                                    _lib.stream_of(text)), "radtts_context_gather")
        ctx.save_for_backward(f2t)
        ctx.shape = (B, C, T1, T2)
        ctx.in_dtype = text_enc.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        (f2t,) = ctx.saved_tensors
        B, C, T1, T2 = ctx.shape
        g = g.float().contiguous()
        out = torch.empty((B, C, T2), dtype=torch.float32, device=g.device)
        _lib.check(_lib.lib().radtts_context_scatter(_lib.ptr(g), _lib.ptr(f2t), B, C, T1, T2, _lib.ptr(out),
                                                     _lib.stream_of(g)), "radtts_context_scatter")
        return out.to(ctx.in_dtype), None


def hard_attention_context(text_enc, frame_to_token):
    """text_enc (B, C, T2), frame_to_token (B, T1) int32 from alignment.mas_forward(..., return_indices=True)."""
    _lib.require_cuda(text_enc, frame_to_token)
    return _ContextGatherFn.apply(text_enc, frame_to_token.contiguous())
