"""Times the persistent BiLSTM recurrence kernels alone (CUDA events): context LSTM (T'=400, B=32, H=520, bf16 operands)
and text-encoder LSTM (T=150, B=32, H=256, split-bf16).  RADTTS_LSTM_CLUSTER=0 selects the cooperative (L2-exchange) kernels,
the default the cluster / DSMEM kernels (csrc/lstm_cluster.cuh)."""
import ctypes
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from radtts_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda", 0)


def run(T, B, H, prec, iters=10):
    g = torch.Generator(device=dev).manual_seed(0)
    gx = torch.randn((2, T, B, 4 * H), device=dev, generator=g) * 0.5
    whh = torch.randn((2, 4 * H, H), device=dev, generator=g) / H ** 0.5
    lens = torch.randint(T // 2, T + 1, (B,), device=dev, generator=g).int()
    lens[0] = T
    h_all = torch.empty((T, B, 2 * H), device=dev)
    gates = torch.empty((2, T, B, 4 * H), device=dev)
    cs = torch.empty((2, T, B, H), device=dev)
    dh = torch.randn((T, B, 2 * H), device=dev, generator=g)
    dg = torch.empty((2, T, B, 4 * H), device=dev)
    nws = int(L.radtts_lstm_workspace_bytes(B, H))
    ws = torch.empty(nws, dtype=torch.uint8, device=dev)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def fwd():
        _lib.check(L.radtts_lstm_forward(_lib.ptr(gx), _lib.ptr(whh), _lib.ptr(lens), T, B, H, _lib.ptr(h_all), _lib.ptr(gates),
                                         _lib.ptr(cs), _lib.ptr(ws), ctypes.c_size_t(nws), prec, st), "fwd")

    def bwd():
        _lib.check(L.radtts_lstm_backward(_lib.ptr(dh), _lib.ptr(whh), _lib.ptr(lens), _lib.ptr(gates), _lib.ptr(cs), T, B, H,
                                          _lib.ptr(dg), _lib.ptr(ws), ctypes.c_size_t(nws), prec, st), "bwd")
    out = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / iters
    fwd()
    dbg = (ctypes.c_ulonglong * 16)()
    L.radtts_lstm_debug_timeline(dbg)
    if dbg[5]:
        print("   fwd cycles/step: wait %.0f  mma %.0f  gates %.0f  send+stores %.0f  gx-load issue %.0f  back edge %.0f" %
              tuple(dbg[i] / dbg[5] for i in (0, 1, 2, 4, 3, 6)))
    return out, float(h_all.double().abs().sum()), float(dg.double().abs().sum())


for (T, B, H, prec, tag) in ((400, 32, 520, 1, "context LSTM bf16"), (150, 32, 256, 2, "text LSTM split-bf16"), (400, 16, 520, 1, "context LSTM bf16 B=16")):
    r, hs, ds = run(T, B, H, prec)
    print("%-28s cluster=%s  fwd %.3f ms (%.2f us/step)  bwd %.3f ms (%.2f us/step)  checksums %.6e %.6e" % (
        tag, os.environ.get("RADTTS_LSTM_CLUSTER", "1"), r["fwd"], r["fwd"] / T * 1e3, r["bwd"], r["bwd"] / T * 1e3, hs, ds))
