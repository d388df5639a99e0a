"""Backward orchestration of the decoder flow stack (host plumbing for radtts_flowstep_backward)."""
import ctypes
import os

import torch

from . import _lib
from . import ops


_aux = {}


def _aux_stream(dev):
    s = _aux.get(dev.index)
    if s is None:
        s = _aux[dev.index] = torch.cuda.Stream(device=dev)
    return s


def _all_direct(ws, sinks, nl):
    """True when every weight_v / weight_g gradient of this flow can be accumulated straight into an existing .grad."""
    if sinks is None:
        return False
    nw = ops.n_weight_tensors(nl)
    idx = [1] + [3 + 2 * j for j in range(nl)] + [3 + 2 * nl + 2 * j for j in range(nl)] + \
          [nw] + [nw + 1 + j for j in range(nl)] + [nw + 1 + nl + j for j in range(nl)]
    for j in idx:
        if ws[j] is None:
            continue               # nothing to write for an absent weight_g
        t = sinks[j]
        gr = None if t is None else t.grad
        if gr is None or gr.dtype != torch.float32 or not gr.is_contiguous() or gr.shape != t.shape:
            return False
    return True


def flow_stack_backward(saved, g_zout, g_log_s):
    plan, dims_list, prec, n_per, saved_list, w_shapes, ctx_dtype = saved
    dev = plan.buf.device
    rows = plan.rows
    act = ops._act_dtype(prec)
    L = _lib.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    z_ld = dims_list[0].z_ld
    ctx_ld = ops.ctx_ld_of(dims_list[0].n_ctx)
    g_z = torch.zeros((rows, z_ld), dtype=torch.float32, device=dev) if g_zout is None else g_zout.float().contiguous()
    g_ctx = torch.zeros((rows, ctx_ld), dtype=torch.float32, device=dev)
    grads = [None] * len(w_shapes)
    first = True
    # Auxiliary stream (direct gradient accumulation + bf16 only): per flow, the helpers that are not on the dgrad chain
    # -- 1x1-conv weight gradient, bias column sums, weight-norm backward, the small accumulations into .grad -- run on
    # `aux` underneath the NEXT flow's tensor-bound GEMMs.  Scratch and weight-gradient buffers then alternate between
    # two pooled sets: flow i-2 may only reuse flow i's set once aux is done with it (ev_aux).
    main = torch.cuda.current_stream(dev)
    use_aux = (prec == ops.PREC_BF16 and not os.environ.get("RADTTS_NO_AUX_STREAM")
               and all(_all_direct(s[3], s[4], d.n_layers) for s, d in zip(saved_list, dims_list)))
    aux = _aux_stream(dev) if use_aux else None
    aux_ptr = ctypes.c_void_p(aux.cuda_stream) if aux is not None else ctypes.c_void_p(0)
    held = []          # [(lease, event)] of the previous flows whose aux work may still be running
    for i in reversed(range(len(dims_list))):
        dims = dims_list[i]
        blob, bufs, lease, ws, sinks = saved_list[i]
        nl, nc, k = dims.n_layers, dims.n_ch, dims.ksize
        h = dims.c_active // 2
        gl = g_log_s[i] if i < len(g_log_s) else None
        gl = None if gl is None else gl.float().contiguous()
        if len(held) >= 2:
            old_lease, old_ev = held.pop(0)
            main.wait_event(old_ev)
            old_lease.release()
        scratch = ops._Lease()
        kpad = (z_ld + 63) // 64 * 64
        g_zin = torch.zeros((rows, z_ld), dtype=torch.float32, device=dev)
        fwd = ops.FlowBuffers(ctx=ops._p(bufs["ctx"]), zin=ops._p(bufs["zin"]), zmid=ops._p(bufs["zmid"]),
                              zout=ops._p(bufs["zout"]), z0=ops._p(bufs["z0"]), x=ops._p(bufs["x"]), r=ops._p(bufs["r"]),
                              params=ops._p(bufs["params"]), log_s=ops._p(bufs["log_s"]))
        gw_inv_full = torch.empty((z_ld, z_ld), dtype=torch.float32, device=dev)
        if aux is not None:
            new = lambda name, *shape: scratch.take(name, shape, torch.float32, dev)     # noqa: E731
        else:
            new = lambda name, *shape: torch.empty(shape, dtype=torch.float32, device=dev)   # noqa: E731
        gw_start = new("gw_start", nc, h + dims.n_ctx)
        gb_start = new("gb_start", nc)
        gw_in = [new("gw_in%d" % j, k, nc, nc) for j in range(nl)]
        gb_in = [new("gb_in%d" % j, nc) for j in range(nl)]
        gw_rs = [new("gw_rs%d" % j, nc, nc) for j in range(nl)]
        gb_rs = [new("gb_rs%d" % j, nc) for j in range(nl)]
        gw_end = new("gw_end", 2 * h, nc)
        gb_end = new("gb_end", 2 * h)
        g = ops.FlowGradBuffers(
            g_zout=ops._p(g_z), g_log_s=ops._p(gl), g_zin=ops._p(g_zin), g_ctx=ops._p(g_ctx),
            g_zmid=ops._p(scratch.take("g_zmid", (rows, z_ld), torch.float32, dev)),
            g_params=ops._p(scratch.take("g_params", (rows, kpad), act, dev)),
            g_u=ops._p(scratch.take("g_u", (nl, rows, nc), act, dev)),
            g_v=ops._p(scratch.take("g_v", (nl, rows, nc), act, dev)),
            g_x0=ops._p(scratch.take("g_x0", (rows, nc), act, dev)),
            g_w_inv_full=ops._p(gw_inv_full), g_w_start=ops._p(gw_start), g_b_start=ops._p(gb_start),
            g_w_end=ops._p(gw_end), g_b_end=ops._p(gb_end),
            scratch_f32=ops._p(scratch.take("scratch_f32", (z_ld, nc + 1), torch.float32, dev)))
        for j in range(nl):
            g.g_w_in[j], g.g_b_in[j] = ops._p(gw_in[j]), ops._p(gb_in[j])
            g.g_w_rs[j], g.g_b_rs[j] = ops._p(gw_rs[j]), ops._p(gb_rs[j])
        _lib.check(L.radtts_flowstep_backward_ex(ctypes.byref(dims), _lib.ptr(blob), plan.ptr, plan.B, plan.Tmax,
                                                 ctypes.byref(fwd), ctypes.byref(g), 0 if first else 1, prec, stream,
                                                 aux_ptr), "radtts_flowstep_backward")
        first = False
        if aux is None:
            scratch.release()
        lease.release()
        saved_list[i] = None
        tail_stream = stream if aux is None else aux_ptr
        base = i * n_per
        nw = ops.n_weight_tensors(nl)
        # weight-norm backward (and the tap-major -> (out, in, tap) re-layout of the in_layer gradients), one launch
        wstruct, keep = ops._weights_struct(ws, nl)
        idx_v = [1] + [3 + 2 * j for j in range(nl)] + [3 + 2 * nl + 2 * j for j in range(nl)]
        idx_g = [nw] + [nw + 1 + j for j in range(nl)] + [nw + 1 + nl + j for j in range(nl)]

        direct = _all_direct(ws, sinks, nl)
        outs = {}
        for j in idx_v + idx_g:
            if ws[j] is None:
                outs[j] = None
            elif direct:
                outs[j] = sinks[j].grad
            else:
                outs[j] = torch.empty(w_shapes[base + j], dtype=torch.float32, device=dev)
        wg = ops.FlowWnGrads(gv_start=ops._p(outs[idx_v[0]]), gg_start=ops._p(outs[idx_g[0]]))
        for j in range(nl):
            wg.gv_in[j], wg.gg_in[j] = ops._p(outs[idx_v[1 + j]]), ops._p(outs[idx_g[1 + j]])
            wg.gv_rs[j], wg.gg_rs[j] = ops._p(outs[idx_v[1 + nl + j]]), ops._p(outs[idx_g[1 + nl + j]])
        _lib.check(L.radtts_flow_weight_norm_backward(ctypes.byref(dims), ctypes.byref(wstruct), ctypes.byref(g),
                                                      ctypes.byref(wg), 1 if (direct and not ops._grads_zeroed) else 0,
                                                      tail_stream),
                   "radtts_flow_weight_norm_backward")
        out = [None] * n_per
        out[0] = gw_inv_full[dims.c_off:, dims.c_off:]
        out[2] = gb_start
        for j in range(nl):
            out[4 + 2 * j] = gb_in[j]
            out[4 + 2 * nl + 2 * j] = gb_rs[j]
        out[3 + 4 * nl] = gw_end
        out[4 + 4 * nl] = gb_end
        if not direct:
            for j in idx_v + idx_g:
                out[j] = outs[j]
        else:
            # the small WN gradients (biases, `end`) take the same route, so that every gradient of the parameter
            # network is final in .grad when this function returns (the data-parallel trainer starts their all-reduce
            # right here, underneath the rest of the backward pass); only the 1x1 conv's LUS factors go through autograd
            with torch.cuda.stream(aux if aux is not None else main):
                for j in range(1, n_per):
                    t, sk = out[j], sinks[j]
                    if t is not None and sk is not None and sk.grad is not None and sk.grad.dtype == t.dtype:
                        sk.grad.add_(t.reshape(sk.grad.shape))
                        out[j] = None
        final = direct
        for j, t in enumerate(out):
            if t is not None:
                assert t.numel() == int(torch.Size(w_shapes[base + j]).numel()), (j, t.shape, w_shapes[base + j])
                grads[base + j] = t.reshape(w_shapes[base + j])
                if j > 0:
                    final = False      # a parameter-network gradient still has to go through AccumulateGrad
        if aux is not None and not final:
            # a gradient of this flow goes back through autograd as a tensor from the pooled scratch: it must be final
            # before the lease can be recycled -> no deferral for this stack (never the case with a fused trainer)
            main.wait_stream(aux)
        if ops.flow_grads_ready is not None:
            # flows finish in the order n-1 .. 0: a data-parallel trainer starts this flow's all-reduce right here
            with torch.cuda.stream(aux if aux is not None else main):
                ops.flow_grads_ready(i, final)
        if aux is not None:
            ev = torch.cuda.Event()
            ev.record(aux)
            held.append((scratch, ev))
        g_z = g_zin
    if aux is not None:
        main.wait_stream(aux)          # join: everything below (LUS backward, clip, optimizer) sees final gradients
        for lease_, _ in held:
            lease_.release()
    g_ctx_out = g_ctx if ctx_dtype == torch.float32 else g_ctx.to(ctx_dtype)
    return (g_z, g_ctx_out, None, None, None, None, None, None, None) + tuple(grads)
