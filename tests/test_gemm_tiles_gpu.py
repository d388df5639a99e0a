"""Device-side tile-width selection of the tcgen05 row GEMM (csrc/rowgemm_tc.cuh): the kernel picks 256 / 208 / 176 / 144
output columns per tile from the number of row tiles the frame plan holds.  Every width accumulates each output element in
the same order, so the decoder flows must give BIT-IDENTICAL results with the selection on and off -- at the benchmarked
shape (85 row tiles: 208 wins), at shapes where other widths win, forward and inverse."""
import pytest
import torch

from radtts_b200 import _lib, configs, ops, synth
from radtts_b200.radtts import RADTTS

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,T1", [(32, 800), (12, 800), (32, 430), (21, 640)])
def test_tile_width_selection_is_bit_identical(cuda_lib, B, T1):
    torch.manual_seed(0)
    model = RADTTS(**configs.model_config("radtts")).eval()
    synth.load_synth(model, seed=1234)
    model = model.cuda()
    batch = synth.synth_batch(B, T1, 150, seed=11)
    mel, out_lens = batch["mel"].cuda(), batch["out_lens"].cuda()
    ctx = torch.randn(B, 1040, T1 // 2, device="cuda") * 0.5
    L = _lib.lib()
    ops.set_precision("bf16")
    outs = {}
    try:
        with torch.no_grad():
            for tag, on in (("off", 0), ("on", 1)):
                old = L.radtts_set_gemm_tile_select(on)
                try:
                    z, _, log_s = ops.decoder_forward(model, mel, ctx, out_lens)
                    x = ops.decoder_inverse(model, z, ctx, out_lens)
                    torch.cuda.synchronize()
                finally:
                    L.radtts_set_gemm_tile_select(old)
                outs[tag] = (z.clone(), [s.clone() for s in log_s], x.clone())
    finally:
        ops.set_precision(None)
    z0, ls0, x0 = outs["off"]
    z1, ls1, x1 = outs["on"]
    assert torch.isfinite(z0).all() and torch.isfinite(x0).all()
    assert torch.equal(z0, z1)
    assert torch.equal(x0, x1)
    for a, b in zip(ls0, ls1):
        assert torch.equal(a, b)
