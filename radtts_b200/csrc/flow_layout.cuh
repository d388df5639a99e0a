// Byte layout of the "prepared weights" blob of one flow step (see radtts_flow_prepare in the public header).
// The forward section comes first and does not depend on want_backward.
#pragma once
#include "common.cuh"

namespace rb {

struct FlowLayout {
  int ctx_ld;     // round_up(n_ctx, 64)
  int end_kpad;   // K of the `end` dgrad GEMM: round_up(z_ld, 64)
  size_t w_inv, w_start, b_start, w_in[RADTTS_MAX_LAYERS], b_in[RADTTS_MAX_LAYERS], w_rs[RADTTS_MAX_LAYERS],
      b_rs[RADTTS_MAX_LAYERS], w_end, b_end;
  size_t scales;  // float [(1 + 2 n_layers)][n_ch]: g / ||v|| of start, in_layers, res_skip_layers (1 where not weight-normed)
  // backward section
  size_t w_inv_t, w_end_t, w_dg[RADTTS_MAX_LAYERS], w_dg0, w_start_t;
  int dg_k[RADTTS_MAX_LAYERS];
  size_t fwd_total, total;
};

inline FlowLayout flow_layout(const radtts_flow_dims& d, int precision, int want_backward) {
  FlowLayout L{};
  const size_t es = precision == RADTTS_PREC_FP32 ? 4 : 2;
  const size_t nc = d.n_ch, k = d.ksize, nl = d.n_layers;
  L.ctx_ld = round_up(d.n_ctx, 64);
  L.end_kpad = round_up(d.z_ld, 64);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += round_up(bytes, (size_t)1024); return o; };
  L.w_inv = take((size_t)d.z_ld * d.z_ld * 4);
  L.w_start = take(nc * (L.ctx_ld + 128) * es);
  L.b_start = take(nc * 4);
  for (size_t i = 0; i < nl; ++i) {
    L.w_in[i] = take(nc * k * nc * es);
    L.b_in[i] = take(nc * 4);
    L.w_rs[i] = take(nc * nc * es);
    L.b_rs[i] = take(nc * 4);
  }
  L.w_end = take((size_t)d.z_ld * nl * nc * es);
  L.b_end = take((size_t)d.z_ld * 4);
  L.scales = take((1 + 2 * nl) * nc * 4);
  L.fwd_total = off;
  if (want_backward) {
    L.w_inv_t = take((size_t)d.z_ld * d.z_ld * 4);
    L.w_end_t = take(nc * L.end_kpad * es);
    for (size_t i = 0; i < nl; ++i) {
      L.dg_k[i] = (int)(nc + (i + 1 < nl ? k * nc : 0));
      L.w_dg[i] = take(nc * (size_t)L.dg_k[i] * es);
    }
    L.w_dg0 = take(nc * k * nc * es);
    L.w_start_t = take((size_t)(L.ctx_ld + 128) * nc * es);
  }
  L.total = off;
  return L;
}

}  // namespace rb
