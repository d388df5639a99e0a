// Generic packed-row conv layer (one row-GEMM launch): the building block of SimpleConvNet (reference
// common.py:475-515) used by the BGAP attribute flows, and of any other ConvNorm stack on packed frames.
//   y[r, n] = act( ratio(r) * sum_{t,c} x[r + (t - k/2) * dil, c] * w[n, c, t]  + bias[n] )   on valid rows, 0 on gaps
#include <cuda_bf16.h>

#include "flow_common.cuh"
#include "wgrad.cuh"

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

// ---- backward -----------------------------------------------------------------------------------------------
// transposed weights for the input-gradient GEMM: dst[c][t * n64 + n] = w[n][c][t]  (zeros in the padding)
template <typename T>
__global__ void conv_prep_t_kernel(const float* __restrict__ w, int c_out, int c_in, int c_in_pad, int k, int n64,
                                   T* __restrict__ dst) {
  const size_t ldw = (size_t)k * n64;
  const size_t total = (size_t)c_in_pad * ldw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % n64);
    const int t = (int)((i / n64) % k);
    const int c = (int)(i / ldw);
    const float v = (n < c_out && c < c_in) ? w[((size_t)n * c_in + c) * k + t] : 0.f;
    if (sizeof(T) == 4) reinterpret_cast<float*>(dst)[i] = v;
    else reinterpret_cast<__nv_bfloat16*>(dst)[i] = __float2bfloat16(v);
  }
}

// gradient of the pre-activation conv output: g_pre = g_y * act'(y) * ratio on the rows the forward wrote, 0 elsewhere
// (gap rows, rows past the valid length of a masked layer, padding columns)
template <typename T>
__global__ void __launch_bounds__(256) conv_bwd_pre_kernel(const T* __restrict__ g_y, int ld_gy, const T* __restrict__ y,
                                                           int ld_y, int y_col_off, int c_out, int n64, int act, int partial,
                                                           int log2d, int ksize, int mask_rows, RowMeta meta,
                                                           const int* __restrict__ tpos, const int* __restrict__ plan,
                                                           T* __restrict__ g_pre) {
  const int rows_used = (plan[0] + 127) / 128 * 128;
  const size_t total = (size_t)rows_used * n64;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % n64);
    const int row = (int)(i / n64);
    float v = 0.f;
    const bool ok = mask_rows ? meta.valid(row) : tpos[row] >= 0;
    if (ok && n < c_out) {
      v = Act<T>::ld(g_y + (size_t)row * ld_gy + n);
      if (act != ACT_NONE) {
        const float yv = Act<T>::ld(y + (size_t)row * ld_y + y_col_off + n);
        v *= act == ACT_RELU ? (yv > 0.f ? 1.f : 0.f) : softplus_grad_t<T>(yv);
      }
      if (partial) v *= meta.ratio(row, log2d, ksize);
    }
    if (sizeof(T) == 4) reinterpret_cast<float*>(g_pre)[i] = v;
    else reinterpret_cast<__nv_bfloat16*>(g_pre)[i] = __float2bfloat16(v);
  }
}

template <typename T>
static int conv_rows_backward_impl(const void* prepared_t, int c_out, int c_in, int c_in_pad, int k, int dl, int act,
                                   int partial, int mask_rows, const void* x, int ld_x, const void* y, int ld_y,
                                   int y_col_off, const void* g_y, int ld_gy, void* g_pre, void* g_x, int ld_gx, float* g_w,
                                   float* g_b, const PlanView& pv, cudaStream_t st) {
  const int n64 = round_up(c_out, 64);
  const int rows = pv.rows_alloc;
  RowMeta meta{pv.pos(), pv.rem()};
  conv_bwd_pre_kernel<T><<<grid_for((size_t)rows * n64), 256, 0, st>>>(
      reinterpret_cast<const T*>(g_y), ld_gy, reinterpret_cast<const T*>(y), ld_y, y_col_off, c_out, n64, act, partial, dl,
      k, mask_rows, meta, pv.tpos(), pv.hdr(), reinterpret_cast<T*>(g_pre));
  RB_TRY(after_launch());
  if (g_x) {
    // g_x[r, c] = sum_t sum_n g_pre[r - (t - k/2) d, n] w[n, c, t]; rows outside the layer's input span get zeros:
    // a partial (masked-input) layer only fed its valid rows, a plain layer its whole geometric span
    GemmDesc g{};
    g.rows_alloc = rows;
    g.plan = pv.hdr();
    g.nseg = k;
    for (int t = 0; t < k; ++t) g.seg[t] = Seg{g_pre, n64, -((t - k / 2) << dl), 0, n64};
    g.w = prepared_t; g.ldw = k * n64; g.N = c_in_pad;
    RowMeta span = partial ? meta : RowMeta{pv.tpos(), pv.tpos()};
    EpiDgradAct<T, ACT_NONE> e{nullptr, 0, reinterpret_cast<T*>(g_x), ld_gx, span, ACT_NONE, 0, 0, k};
    RB_TRY((run_gemm<T>(g, e, st)));
  }
  // weight gradient in the reference layout (c_out, c_in, k): one problem per tap, K = packed rows
  if constexpr (sizeof(T) == 2) {
    WgradBatch batch;
    batch.rows_alloc = rows;
    bool all = true;
    for (int t = 0; t < k && all; ++t) {
      WgradProb p{g_pre, n64, 0, c_out, x, ld_x, 0, c_in, (t - k / 2) << dl, g_w + t, (long)c_in * k, k};
      all = batch.add(p, false);
    }
    if (!all) return RADTTS_ERR_UNSUPPORTED;
    RB_TRY(batch.launch(pv.hdr(), st));
  } else {
    RB_CUDA(cudaMemsetAsync(g_w, 0, (size_t)c_out * c_in * k * sizeof(float), st));
    for (int t = 0; t < k; ++t) {
      WgradProb p{g_pre, n64, 0, c_out, x, ld_x, 0, c_in, (t - k / 2) << dl, g_w + t, (long)c_in * k, k};
      RB_TRY((launch_wgrad_simt<T, T>(p, pv.hdr(), rows, st)));
    }
  }
  if (g_b) {
    // the bias sits outside the partial-conv renormalisation (partialconv1d.py:63): g_b = sum_r g_pre / ratio
    ColsumBatch<T> sums;
    RB_TRY(sums.add(reinterpret_cast<const T*>(g_pre), n64, 0, round_up(c_out, 4), partial, dl, g_b, 1, true, st));
    RB_TRY(sums.launch(meta, k, pv.hdr(), rows, st));
  }
  return 0;
}

}  // namespace rb

using namespace rb;

extern "C" size_t radtts_conv_backward_prepared_bytes(int c_out, int c_in_pad, int ksize, int precision) {
  if (c_out <= 0 || c_in_pad <= 0 || ksize <= 0) return 0;
  return round_up((size_t)c_in_pad * ksize * round_up(c_out, 64) * (precision == RADTTS_PREC_FP32 ? 4 : 2), (size_t)256);
}

extern "C" int radtts_conv_prepare_backward(const float* w, int c_out, int c_in, int c_in_pad, int ksize, int precision,
                                            void* prepared_t, size_t prepared_bytes, void* stream) {
  if (!w || !prepared_t || c_out <= 0 || c_in <= 0 || c_in_pad < c_in || c_in_pad % 64 || ksize <= 0 || ksize % 2 == 0)
    return RADTTS_ERR_INVALID_ARG;
  if (prepared_bytes < radtts_conv_backward_prepared_bytes(c_out, c_in_pad, ksize, precision)) return RADTTS_ERR_WORKSPACE;
  const int n64 = round_up(c_out, 64);
  const size_t total = (size_t)c_in_pad * ksize * n64;
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RADTTS_PREC_FP32)
    conv_prep_t_kernel<float><<<grid_for(total), 256, 0, st>>>(w, c_out, c_in, c_in_pad, ksize, n64,
                                                              reinterpret_cast<float*>(prepared_t));
  else if (precision == RADTTS_PREC_BF16)
    conv_prep_t_kernel<__nv_bfloat16><<<grid_for(total), 256, 0, st>>>(w, c_out, c_in, c_in_pad, ksize, n64,
                                                                      reinterpret_cast<__nv_bfloat16*>(prepared_t));
  else
    return RADTTS_ERR_INVALID_ARG;
  return after_launch();
}

extern "C" int radtts_conv_rows_backward(const void* prepared_t, int c_out, int c_in, int c_in_pad, int ksize, int dilation,
                                         int act, int partial, int mask_rows, const void* x, int ld_x, const void* y,
                                         int ld_y, int y_col_off, const void* g_y, int ld_gy, void* g_pre, void* g_x,
                                         int ld_gx, float* g_w, float* g_b, const void* plan, int B, int Tmax,
                                         int precision, void* stream) {
  if (!prepared_t || !x || !g_y || !g_pre || !g_w || !plan || c_out <= 0 || c_in <= 0 || c_in_pad < c_in ||
      c_in_pad % 64 || ksize <= 0 || ksize % 2 == 0 || ksize > kMaxSeg || dilation <= 0 || (dilation & (dilation - 1)))
    return RADTTS_ERR_INVALID_ARG;
  if (act != ACT_NONE && !y) return RADTTS_ERR_INVALID_ARG;
  if ((ksize / 2) * dilation > kGap) return RADTTS_ERR_UNSUPPORTED;
  if (ld_x < c_in_pad || ld_gy < c_out || (g_x && ld_gx < c_in_pad)) return RADTTS_ERR_INVALID_ARG;
  int dl = 0;
  while ((1 << dl) < dilation) ++dl;
  PlanView pv = make_plan_view(plan, B, Tmax);
  cudaStream_t st = (cudaStream_t)stream;
  if (precision == RADTTS_PREC_FP32)
    return conv_rows_backward_impl<float>(prepared_t, c_out, c_in, c_in_pad, ksize, dl, act, partial, mask_rows, x, ld_x, y,
                                          ld_y, y_col_off, g_y, ld_gy, g_pre, g_x, ld_gx, g_w, g_b, pv, st);
  if (precision == RADTTS_PREC_BF16)
    return conv_rows_backward_impl<__nv_bfloat16>(prepared_t, c_out, c_in, c_in_pad, ksize, dl, act, partial, mask_rows, x,
                                                  ld_x, y, ld_y, y_col_off, g_y, ld_gy, g_pre, g_x, ld_gx, g_w, g_b, pv, st);
  return RADTTS_ERR_INVALID_ARG;
}

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
