"""Fused RAdam (reference radam.py) on a flat parameter buffer.

All trainable parameters become views into ONE contiguous fp32 buffer and their gradients views into a second one,
so gradient clipping is one norm over a flat tensor and the optimizer is one kernel launch
(csrc/optim.cu: radtts_radam_step).  Sync-free and CUDA-graph capturable (step counter on the device)."""
import ctypes

import torch

from . import _lib


class FusedRAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params), \
            "FusedRAdam needs fp32 CUDA parameters (no CPU fallback)"
        dev = self.params[0].device
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]          # keep every view 16-byte aligned
        total = sum(sizes)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        off = 0
        with torch.no_grad():
            for p, n in zip(self.params, sizes):
                view = self.flat[off:off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
                p.grad = self.grad[off:off + p.numel()].view_as(p)
                off += n

    def zero_grad(self, set_to_none=False):
        """Gradients stay views into the flat buffer (autograd accumulates into them in place)."""
        self.grad.zero_()
        off = 0
        for p in self.params:
            n = (p.numel() + 3) // 4 * 4
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + off * 4:
                p.grad = self.grad[off:off + p.numel()].view_as(p)
            off += n

    def clip_coefficient(self, max_norm):
        """torch.nn.utils.clip_grad_norm_ semantics: min(1, max_norm / (||g||_2 + 1e-6)), kept on the device."""
        total_norm = torch.linalg.vector_norm(self.grad)
        return (max_norm / (total_norm + 1e-6)).clamp(max=1.0).reshape(1)

    def check_views(self):
        """Every parameter must still be a view of the flat buffer (nn.LSTM.flatten_parameters() and `p.data = ...`
        re-point storage silently; the fused kernel would then update memory nobody reads)."""
        off = 0
        for p in self.params:
            if p.data_ptr() != self.flat.data_ptr() + off * 4:
                raise _lib.RadttsB200Error("FusedRAdam: a parameter of shape %s no longer lives in the flat buffer "
                                           "(was flatten_parameters() or `.data =` applied to it?)" % (tuple(p.shape),))
            off += (p.numel() + 3) // 4 * 4

    def step(self, grad_scale=None):
        self.step_range(0, self.flat.numel(), grad_scale)

    def step_range(self, lo, hi, grad_scale=None, enable=None, zero_grad=False, bump=True):
        """The update on flat-buffer elements [lo, hi) (lo a multiple of 4).  enable: device int32 tensor, 0 makes the call a
        no-op; zero_grad: write zeros over the gradients in the same pass; bump: advance the step counter afterwards.  All
        ranges of one optimizer step must run before the bump of that step (radtts_radam_step_ex)."""
        if hi <= lo:
            return
        assert lo % 4 == 0
        if not torch.cuda.is_current_stream_capturing():
            self.check_views()
        L = _lib.lib()
        b1, b2 = self.betas
        f = lambda t: ctypes.c_void_p(t.data_ptr() + 4 * lo)      # noqa: E731
        _lib.check(L.radtts_radam_step_ex(f(self.flat), f(self.grad), f(self.exp_avg), f(self.exp_avg_sq),
                                          ctypes.c_size_t(hi - lo), ctypes.c_float(self.lr), ctypes.c_float(b1),
                                          ctypes.c_float(b2), ctypes.c_float(self.eps), ctypes.c_float(self.weight_decay),
                                          _lib.ptr(self.step_dev), _lib.ptr(grad_scale), _lib.ptr(enable), int(bool(zero_grad)),
                                          int(bool(bump)), ctypes.c_void_p(torch.cuda.current_stream(self.flat.device).cuda_stream)),
                   "radtts_radam_step_ex")
