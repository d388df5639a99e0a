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
