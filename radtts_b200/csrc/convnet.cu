// Generic packed-row conv layer (one row-GEMM launch): the building block of SimpleConvNet (reference
// common.py:475-515) used by the BGAP attribute flows, and of any other ConvNorm stack on packed frames.
//   y[r, n] = act( ratio(r) * sum_{t,c} x[r + (t - k/2) * dil, c] * w[n, c, t]  + bias[n] )   on valid rows, 0 on gaps
#include <cuda_bf16.h>

#include "flow_common.cuh"

namespace rb {

template <typename T>
__global__ void conv_prep_kernel(const float* __restrict__ w, const float* __restrict__ b, int c_out, int c_in,
                                 int c_in_pad, int k, int n_pad, T* __restrict__ dst, float* __restrict__ bdst) {
  const size_t ldw = (size_t)k * c_in_pad;
  const size_t total = (size_t)n_pad * ldw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c_in_pad);
    const int t = (int)((i / c_in_pad) % k);
    const int n = (int)(i / ldw);
    const float v = (n < c_out && c < c_in) ? w[((size_t)n * c_in + c) * k + t] : 0.f;
    if (sizeof(T) == 4) reinterpret_cast<float*>(dst)[i] = v;
    else reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16(v);
    if (c == 0 && t == 0) bdst[n] = (n < c_out && b) ? b[n] : 0.f;
  }
}

static inline size_t conv_w_bytes(int n_pad, int c_in_pad, int k, int precision) {
  return round_up((size_t)n_pad * k * c_in_pad * (precision == RADTTS_PREC_FP32 ? 4 : 2), (size_t)256);
}

template <typename T>
static int conv_rows_impl(const uint8_t* prepared, int c_out, int c_in_pad, int k, int dil_log2, int act, int partial,
                          int mask_rows, const void* x, int ld_x, void* y, int ld_y, int y_col_off, const PlanView& pv,
                          cudaStream_t st) {
  const int n_pad = round_up(c_out, 16);
  const size_t wbytes = conv_w_bytes(n_pad, c_in_pad, k, sizeof(T) == 4 ? RADTTS_PREC_FP32 : RADTTS_PREC_BF16);
  RowMeta meta{pv.pos(), pv.rem()};
  GemmDesc g{};
  g.rows_alloc = pv.rows_alloc;
  g.plan = pv.hdr();
  g.nseg = k;
  for (int t = 0; t < k; ++t) g.seg[t] = Seg{x, ld_x, (t - k / 2) << dil_log2, 0, c_in_pad};
  g.w = prepared; g.ldw = k * c_in_pad; g.N = n_pad;
  EpiBiasAct<T> e{reinterpret_cast<T*>(y), ld_y, y_col_off, reinterpret_cast<const float*>(prepared + wbytes), meta, act,
                  partial, dil_log2, k, mask_rows};
  return run_gemm<T>(g, e, st);
}

}  // namespace rb

using namespace rb;

extern "C" size_t radtts_conv_prepared_bytes(int c_out, int c_in_pad, int ksize, int precision) {
  if (c_out <= 0 || c_in_pad <= 0 || ksize <= 0) return 0;
  const int n_pad = round_up(c_out, 16);
  return conv_w_bytes(n_pad, c_in_pad, ksize, precision) + round_up((size_t)n_pad * 4, (size_t)256);
}

extern "C" int radtts_conv_prepare(const float* w, const float* bias, int c_out, int c_in, int c_in_pad, int ksize,
                                   int precision, void* prepared, size_t prepared_bytes, void* stream) {
  if (!w || !prepared || c_out <= 0 || c_in <= 0 || c_in_pad < c_in || c_in_pad % 64 || ksize <= 0 || ksize % 2 == 0)
    return RADTTS_ERR_INVALID_ARG;
  if (prepared_bytes < radtts_conv_prepared_bytes(c_out, c_in_pad, ksize, precision)) return RADTTS_ERR_WORKSPACE;
  const int n_pad = round_up(c_out, 16);
  uint8_t* base = reinterpret_cast<uint8_t*>(prepared);
  float* bdst = reinterpret_cast<float*>(base + conv_w_bytes(n_pad, c_in_pad, ksize, precision));
  const size_t total = (size_t)n_pad * ksize * c_in_pad;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RADTTS_PREC_FP32)
    conv_prep_kernel<float><<<grid_for(total), 256, 0, st>>>(w, bias, c_out, c_in, c_in_pad, ksize, n_pad,
                                                            reinterpret_cast<float*>(base), bdst);
  else if (precision == RADTTS_PREC_BF16)
    conv_prep_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>(w, bias, c_out, c_in, c_in_pad, ksize, n_pad,
                                                                    reinterpret_cast<__nv_bfloat16*>(base), bdst);
  else
    return RADTTS_ERR_INVALID_ARG;
  return after_launch();
}

extern "C" int radtts_conv_rows(const void* prepared, int c_out, int c_in_pad, int ksize, int dilation, int act,
                                int partial, int mask_rows, const void* x, int ld_x, void* y, int ld_y, int y_col_off,
                                const void* plan, int B, int Tmax, int precision, void* stream) {
  if (!prepared || !x || !y || !plan || c_out <= 0 || c_in_pad <= 0 || c_in_pad % 64 || ksize <= 0 || ksize % 2 == 0 ||
      ksize > kMaxSeg || dilation <= 0 || (dilation & (dilation - 1)))
    return RADTTS_ERR_INVALID_ARG;
  if ((ksize / 2) * dilation > kGap) return RADTTS_ERR_UNSUPPORTED;
  if (ld_x < c_in_pad || ld_y < y_col_off + round_up(c_out, 16)) return RADTTS_ERR_INVALID_ARG;
  int dl = 0;
  while ((1 << dl) < dilation) ++dl;
  PlanView pv = make_plan_view(plan, B, Tmax);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(prepared);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RADTTS_PREC_FP32)
    return conv_rows_impl<float>(base, c_out, c_in_pad, ksize, dl, act, partial, mask_rows, x, ld_x, y, ld_y, y_col_off,
                                 pv, st);
  if (precision == RADTTS_PREC_BF16)
    return conv_rows_impl<__nv_bfloat16>(base, c_out, c_in_pad, ksize, dl, act, partial, mask_rows, x, ld_x, y, ld_y,
                                         y_col_off, pv, st);
  return RADTTS_ERR_INVALID_ARG;
}
