// Attribute-flow pieces (BGAP, reference attribute_prediction_model.py:120-224) that are not GEMMs:
//   * rational-quadratic spline coupling transform  (splines.py:221-319 + common.py:699-743), both directions
//   * affine coupling apply for the simple_conv flows (common.py:782-784,821-832)
//   * plain invertible 1x1 conv on tiny channel counts (common.py:431-472: 4x4 for F0, 8x8 for energy)
// All operate on reference-shaped (B, C, T) float32 tensors; one thread per (b, t), coalesced along t.
#include <math_constants.h>

#include "common.cuh"

namespace rb {

constexpr int kMaxSplineBins = 32;  // n_bins (w) <= 32, v has n_bins + 1 entries

// params: (B, h * (2*nb + 1), T), channel index c * (2*nb+1) + i   (common.py:708-711 reshape)
// x:      (B, C, T) whole coupling input; the transformed half is channels [h, 2h)
__global__ void __launch_bounds__(128) rqspline_kernel(const float* __restrict__ x, const float* __restrict__ params, int B,
                                                       int C, int h, int T, int nb, int inverse, float left, float right,
                                                       float bottom, float top, float* __restrict__ y,
                                                       float* __restrict__ log_s) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  const float eps = 1.1920929e-07f;
  const int np = 2 * nb + 1;
  const float in_lo = inverse ? bottom : left, in_hi = inverse ? top : right;
  const float out_lo = inverse ? left : bottom, out_hi = inverse ? right : top;
  float ls_total = 0.f;
  // untransformed half passes through (torch.cat((z_0, z_1)) common.py:731,735)
  for (int c = 0; c < h; ++c) y[((size_t)b * C + c) * T + t] = x[((size_t)b * C + c) * T + t];
  for (int c = 0; c < h; ++c) {
    const size_t xi = ((size_t)b * C + h + c) * T + t;
    const float xn = (x[xi] - in_lo) / (in_hi - in_lo);
    float outn = xn, lj = 0.f;
    if (xn >= 0.f && xn < 1.f) {
      const float* p = params + ((size_t)b * h * np + (size_t)c * np) * T + t;
      float w[kMaxSplineBins], v[kMaxSplineBins + 1];
      float m = -CUDART_INF_F;
      for (int i = 0; i < nb; ++i) { w[i] = p[(size_t)i * T]; m = fmaxf(m, w[i]); }
      float s = 0.f;
      for (int i = 0; i < nb; ++i) { w[i] = expf(w[i] - m); s += w[i]; }
      for (int i = 0; i < nb; ++i) w[i] /= s;
      float vm = -CUDART_INF_F;
      for (int i = 0; i <= nb; ++i) { v[i] = p[(size_t)(nb + i) * T]; vm = fmaxf(vm, v[i]); }
      for (int i = 0; i <= nb; ++i) v[i] = expf(v[i] - vm) + 1e-8f;
      float vs = 0.f;
      for (int i = 0; i < nb; ++i) vs += (v[i] + v[i + 1]) / 2.f * w[i];
      for (int i = 0; i <= nb; ++i) v[i] /= vs;
      // walk the knots: wc = cumsum(w) (last forced to 1), cdf = cumsum(area) (last forced to 1); searchsorted(left)
      float wc = 0.f, cdf = 0.f, wc_prev = 0.f, cdf_prev = 0.f;
      int bin = nb - 1;
      float w_lo = 0.f, c_lo = 0.f;
      bool found = false;
      for (int i = 0; i < nb; ++i) {
        wc_prev = wc; cdf_prev = cdf;
        wc += w[i];
        cdf += (v[i + 1] + v[i]) / 2.f * w[i];
        const float wck = (i == nb - 1) ? 1.f : wc;
        const float cdk = (i == nb - 1) ? 1.f : cdf;
        const float knot = inverse ? cdk : wck;
        if (!found && knot >= xn) { found = true; bin = i; w_lo = wc_prev; c_lo = cdf_prev; }
      }
      if (!found) { bin = nb - 1; w_lo = wc_prev; c_lo = cdf_prev; }
      const float w_b = w[bin], v_b = v[bin], v_n = v[bin + 1];
      if (!inverse) {
        const float alpha = (xn - w_lo) / fmaxf(w_b, eps);
        float cval = alpha * alpha / 2.f * (v_n - v_b) * w_b + alpha * v_b * w_b + c_lo;
        const float dens = alpha < 0.5f ? v_b + alpha * (v_n - v_b) : v_n - (v_n - v_b) * (1.f - alpha);  // torch.lerp
        lj = logf(fmaxf(dens, eps));
        outn = fminf(fmaxf(cval, eps), 1.f - eps);
      } else {
        const float qa = (v_n - v_b) * w_b / 2.f;
        const float qb = v_b * w_b;
        const float qc = c_lo - xn;
        const float alpha = (-qb + sqrtf(qb * qb - 4.f * qa * qc)) / (2.f * qa);
        outn = fminf(fmaxf(alpha * w_b + w_lo, eps), 1.f - eps);
      }
    }
    y[xi] = outn * (out_hi - out_lo) + out_lo;
    ls_total += lj;
  }
  if (log_s && !inverse) log_s[(size_t)b * T + t] = ls_total + (float)h * (logf(top - bottom) - logf(right - left));
}

// Backward of the forward-direction spline coupling (training the attribute flows): closed-form gradients, term by term
// as in oracle/spline_grad.py (which is pinned against autograd).  One thread per (b, t); for every transformed channel
// the forward quantities are recomputed from the 2 nb + 1 raw parameters, then
//   g_x (B, C, T): pass-through half = g_y, transformed half = d y / d x (+ d log_s / d x)
//   g_params (B, h (2 nb + 1), T): gradients of the raw widths / heights (zero where x is outside the spline's interval)
__global__ void __launch_bounds__(128) rqspline_bwd_kernel(const float* __restrict__ x, const float* __restrict__ params,
                                                           const float* __restrict__ g_y, const float* __restrict__ g_log_s,
                                                           int B, int C, int h, int T, int nb, float left, float right,
                                                           float bottom, float top, float* __restrict__ g_x,
                                                           float* __restrict__ g_params) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  const float eps = 1.1920929e-07f;
  const int np = 2 * nb + 1;
  const float gl = g_log_s ? g_log_s[(size_t)b * T + t] : 0.f;     // d L / d log_j, the same for every channel
  for (int c = 0; c < h; ++c) {
    const size_t i0 = ((size_t)b * C + c) * T + t;
    g_x[i0] = g_y ? g_y[i0] : 0.f;
  }
  for (int c = 0; c < h; ++c) {
    const size_t xi = ((size_t)b * C + h + c) * T + t;
    const float xn = (x[xi] - left) / (right - left);
    const float g_out = (g_y ? g_y[xi] : 0.f) * (top - bottom);    // y = out_n (top - bottom) + bottom
    float* gp = g_params + ((size_t)b * h * np + (size_t)c * np) * T + t;
    if (!(xn >= 0.f && xn < 1.f)) {                                // identity outside the interval
      g_x[xi] = g_out / (right - left);
      for (int i = 0; i < np; ++i) gp[(size_t)i * T] = 0.f;
      continue;
    }
    const float* p = params + ((size_t)b * h * np + (size_t)c * np) * T + t;
    float w[kMaxSplineBins], v[kMaxSplineBins + 1], e[kMaxSplineBins + 1], gw[kMaxSplineBins], gv[kMaxSplineBins + 1];
    float m = -CUDART_INF_F;
    for (int i = 0; i < nb; ++i) { w[i] = p[(size_t)i * T]; m = fmaxf(m, w[i]); }
    float s = 0.f;
    for (int i = 0; i < nb; ++i) { w[i] = expf(w[i] - m); s += w[i]; }
    for (int i = 0; i < nb; ++i) w[i] /= s;
    float vm = -CUDART_INF_F;
    int am = 0;
    for (int i = 0; i <= nb; ++i) { v[i] = p[(size_t)(nb + i) * T]; if (v[i] > vm) { vm = v[i]; am = i; } }
    for (int i = 0; i <= nb; ++i) { e[i] = expf(v[i] - vm); v[i] = e[i] + 1e-8f; }
    float S = 0.f;
    for (int i = 0; i < nb; ++i) S += (v[i] + v[i + 1]) / 2.f * w[i];
    for (int i = 0; i <= nb; ++i) v[i] /= S;
    // bin of xn in wc = cumsum(w) (last knot forced to 1), searchsorted(left)
    float wc = 0.f, cdf = 0.f, w_lo = 0.f, c_lo = 0.f;
    int bin = nb - 1;
    bool found = false;
    for (int i = 0; i < nb; ++i) {
      const float wc_prev = wc, cdf_prev = cdf;
      wc += w[i];
      cdf += (v[i + 1] + v[i]) / 2.f * w[i];
      const float knot = (i == nb - 1) ? 1.f : wc;
      if (!found && knot >= xn) { found = true; bin = i; w_lo = wc_prev; c_lo = cdf_prev; }
      if (!found && i == nb - 1) { w_lo = wc_prev; c_lo = cdf_prev; }
    }
    const float w_b = w[bin], v_b = v[bin], v_n = v[bin + 1];
    const float alpha = (xn - w_lo) / fmaxf(w_b, eps);
    const float D = v_n - v_b, L = v_b + alpha * D;
    const float cval = alpha * alpha / 2.f * D * w_b + alpha * v_b * w_b + c_lo;
    const float gy = (cval > eps && cval < 1.f - eps) ? g_out : 0.f;         // the output clamp's gradient
    const float g_alpha = gy * w_b * L + gl * D / L;
    g_x[xi] = g_alpha / w_b / (right - left);
    for (int i = 0; i < nb; ++i) gw[i] = 0.f;
    for (int i = 0; i <= nb; ++i) gv[i] = 0.f;
    gw[bin] += gy * (alpha * alpha / 2.f * D + alpha * v_b) - g_alpha * alpha / w_b;
    gv[bin] += gy * (alpha - alpha * alpha / 2.f) * w_b + gl * (1.f - alpha) / L;
    gv[bin + 1] += gy * alpha * alpha / 2.f * w_b + gl * alpha / L;
    for (int j = 0; j < bin; ++j) {                                            // the prefix sums wc0_b and cdf0_b
      gw[j] += -g_alpha / w_b + gy * (v[j] + v[j + 1]) / 2.f;
      gv[j] += gy * w[j] / 2.f;
      gv[j + 1] += gy * w[j] / 2.f;
    }
    float dot = 0.f;
    for (int i = 0; i <= nb; ++i) dot += gv[i] * v[i];
    for (int i = 0; i < nb; ++i) gw[i] -= dot * (v[i] + v[i + 1]) / 2.f;     // S depends on w
    float gsum = 0.f;
    for (int i = 0; i <= nb; ++i) {                                            // v = u / S, u = e + 1e-8
      const float wl = i > 0 ? w[i - 1] : 0.f, wr = i < nb ? w[i] : 0.f;
      const float gu = (gv[i] - dot * (wl + wr) / 2.f) / S;
      gv[i] = gu * e[i];                                                       // d / d v~_i (before the max's share)
      gsum += gv[i];
    }
    gv[am] -= gsum;                                                            // the max's sub-gradient, as autograd
    float wdot = 0.f;
    for (int i = 0; i < nb; ++i) wdot += gw[i] * w[i];
    for (int i = 0; i < nb; ++i) gp[(size_t)i * T] = w[i] * (gw[i] - wdot);    // softmax
    for (int i = 0; i <= nb; ++i) gp[(size_t)(nb + i) * T] = gv[i];
  }
}

// Affine coupling apply: params (B, 2h, T) = [raw scale | translation]; z (B, 2h, T).
__global__ void __launch_bounds__(256) affine_apply_kernel(const float* __restrict__ z, const float* __restrict__ params,
                                                           int h, int T, int scaling, int inverse, float* __restrict__ y,
                                                           float* __restrict__ log_s) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const size_t base = (size_t)b * 2 * h * T;
  const size_t i0 = base + (size_t)c * T + t, i1 = base + (size_t)(h + c) * T + t;
  y[i0] = z[i0];
  const float xr = params[i0], tr = params[i1];
  float s, ls;
  if (scaling == 0) { s = (tanhf(xr) + 1.f) + 1e-6f; ls = logf(s); }
  else if (scaling == 1) { s = expf(xr); ls = xr; }
  else if (scaling == 2) { s = 1.f / (1.f + expf(-(xr + 10.f))) + 1e-6f; ls = logf(s); }
  else { s = 1.f; ls = 0.f; }
  if (inverse) y[i1] = (z[i1] - tr) / s;
  else { y[i1] = s * z[i1] + tr; if (log_s) log_s[base / 2 + (size_t)c * T + t] = ls; }
}

// Backward of the forward-direction affine coupling (what autograd derives for common.py:782-784,821-832):
//   y1 = s z1 + b, log_s = log s, s = f(x)  ->  g_z0 = g_y0, g_z1 = s g_y1, g_b = g_y1, g_x = (g_y1 z1 + g_log_s / s) ds/dx
__global__ void __launch_bounds__(256) affine_bwd_kernel(const float* __restrict__ z, const float* __restrict__ params,
                                                         const float* __restrict__ g_y, const float* __restrict__ g_log_s,
                                                         int h, int T, int scaling, float* __restrict__ g_z,
                                                         float* __restrict__ g_params) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const size_t base = (size_t)b * 2 * h * T;
  const size_t i0 = base + (size_t)c * T + t, i1 = base + (size_t)(h + c) * T + t;
  const float gy0 = g_y ? g_y[i0] : 0.f, gy1 = g_y ? g_y[i1] : 0.f;
  const float gl = g_log_s ? g_log_s[base / 2 + (size_t)c * T + t] : 0.f;
  const float x = params[i0];
  float s, ds, dls;
  if (scaling == 0) { const float th = tanhf(x); s = (th + 1.f) + 1e-6f; ds = 1.f - th * th; dls = ds / s; }
  else if (scaling == 1) { s = expf(x); ds = s; dls = 1.f; }
  else if (scaling == 2) { const float sg = 1.f / (1.f + expf(-(x + 10.f))); s = sg + 1e-6f; ds = sg * (1.f - sg); dls = ds / s; }
  else { s = 1.f; ds = 0.f; dls = 0.f; }
  g_z[i0] = gy0;
  g_z[i1] = s * gy1;
  g_params[i0] = gy1 * z[i1] * ds + gl * dls;
  g_params[i1] = gy1;
}

// Backward of the small dense 1x1 conv y[b,:,t] = W x[b,:,t] (C <= 16):  g_x = W^T g_y;  g_W[c][j] = sum_{b,t} g_y[c] x[j]
__global__ void __launch_bounds__(256) pointwise_small_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                  const float* __restrict__ g_y, int C, int T,
                                                                  float* __restrict__ g_x, float* __restrict__ g_w) {
  __shared__ float ws[16 * 16];
  __shared__ float acc_s[16 * 16];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) { ws[i] = w[i]; acc_s[i] = 0.f; }
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  float xv[16], gv[16];
  const bool ok = t < T;
  for (int j = 0; j < C; ++j) {
    xv[j] = ok ? x[((size_t)b * C + j) * T + t] : 0.f;
    gv[j] = ok ? g_y[((size_t)b * C + j) * T + t] : 0.f;
  }
  if (ok) {
    for (int j = 0; j < C; ++j) {
      float a = 0.f;
      for (int c = 0; c < C; ++c) a = fmaf(ws[c * C + j], gv[c], a);
      g_x[((size_t)b * C + j) * T + t] = a;
    }
  }
  for (int c = 0; c < C; ++c)
    for (int j = 0; j < C; ++j) {
      float v = gv[c] * xv[j];
#pragma unroll
      for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&acc_s[c * C + j], v);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) atomicAdd(&g_w[i], acc_s[i]);
}

// out[b][c][t] = sum_j W[c][j] x[b][j][t], C <= 16
__global__ void __launch_bounds__(256) pointwise_small_kernel(const float* __restrict__ x, const float* __restrict__ w, int C,
                                                              int T, float* __restrict__ y) {
  __shared__ float ws[16 * 16];
  for (int i = threadIdx.x; i < C * C; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  float v[16];
  for (int j = 0; j < C; ++j) v[j] = x[((size_t)b * C + j) * T + t];
  for (int c = 0; c < C; ++c) {
    float acc = 0.f;
    for (int j = 0; j < C; ++j) acc = fmaf(ws[c * C + j], v[j], acc);
    y[((size_t)b * C + c) * T + t] = acc;
  }
}

}  // namespace rb

using namespace rb;

extern "C" int radtts_rqspline_apply(const float* x, const float* params, int B, int C, int T, int n_bins, int inverse,
                                     float left, float right, float bottom, float top, float* y, float* log_s,
                                     void* stream) {
  if (!x || !params || !y || B <= 0 || C <= 0 || C % 2 || T <= 0 || n_bins <= 0) return RADTTS_ERR_INVALID_ARG;
  if (n_bins > kMaxSplineBins) return RADTTS_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(T, 128), B);
  rqspline_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, params, B, C, C / 2, T, n_bins, inverse, left, right, bottom,
                                                         top, y, log_s);
  return after_launch();
}

extern "C" int radtts_affine_apply(const float* z, const float* params, int B, int C, int T, int scaling, int inverse,
                                   float* y, float* log_s, void* stream) {
  if (!z || !params || !y || B <= 0 || C <= 0 || C % 2 || T <= 0) return RADTTS_ERR_INVALID_ARG;
  dim3 grid(ceil_div(T, 256), C / 2, B);
  affine_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, params, C / 2, T, scaling, inverse, y, log_s);
  return after_launch();
}

extern "C" int radtts_pointwise_conv_small(const float* x, const float* w, int B, int C, int T, float* y, void* stream) {
  if (!x || !w || !y || B <= 0 || C <= 0 || T <= 0) return RADTTS_ERR_INVALID_ARG;
  if (C > 16) return RADTTS_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(T, 256), B);
  pointwise_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, C, T, y);
  return after_launch();
}

extern "C" int radtts_rqspline_backward(const float* x, const float* params, const float* g_y, const float* g_log_s, int B,
                                        int C, int T, int n_bins, float left, float right, float bottom, float top,
                                        float* g_x, float* g_params, void* stream) {
  if (!x || !params || !g_x || !g_params || B <= 0 || C <= 0 || C % 2 || T <= 0 || n_bins <= 0) return RADTTS_ERR_INVALID_ARG;
  if (n_bins > kMaxSplineBins) return RADTTS_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(T, 128), B);
  rqspline_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(x, params, g_y, g_log_s, B, C, C / 2, T, n_bins, left, right,
                                                             bottom, top, g_x, g_params);
  return after_launch();
}

extern "C" int radtts_affine_backward(const float* z, const float* params, const float* g_y, const float* g_log_s, int B,
                                      int C, int T, int scaling, float* g_z, float* g_params, void* stream) {
  if (!z || !params || !g_z || !g_params || B <= 0 || C <= 0 || C % 2 || T <= 0) return RADTTS_ERR_INVALID_ARG;
  dim3 grid(ceil_div(T, 256), C / 2, B);
  affine_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(z, params, g_y, g_log_s, C / 2, T, scaling, g_z, g_params);
  return after_launch();
}

extern "C" int radtts_pointwise_conv_small_backward(const float* x, const float* w, const float* g_y, int B, int C, int T,
                                                    float* g_x, float* g_w, void* stream) {
  if (!x || !w || !g_y || !g_x || !g_w || B <= 0 || C <= 0 || T <= 0) return RADTTS_ERR_INVALID_ARG;
  if (C > 16) return RADTTS_ERR_UNSUPPORTED;
  RB_CUDA(cudaMemsetAsync(g_w, 0, (size_t)C * C * sizeof(float), (cudaStream_t)stream));
  dim3 grid(ceil_div(T, 256), B);
  pointwise_small_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, w, g_y, C, T, g_x, g_w);
  return after_launch();
}
