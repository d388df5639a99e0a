// Text-encoder conv block epilogue (reference common.py:348-356 run per utterance; Encoder.convolutions = ConvNorm with
// partial padding + InstanceNorm1d(affine) + ReLU + dropout): everything AFTER the k-tap convolution, for the whole padded
// batch in one launch per direction, instead of ~15 element-wise / reduction launches forward and ~30 backward:
//   y   = ((raw - b) * k / (cnt + 1e-6) + b)            t <  len   (partialconv1d.py:51-66; cnt = taps inside [0, len))
//       = 0                                             t >= len
//   z   = (y - mean) * rstd,  mean / var over the len valid frames of (utterance, channel)      (InstanceNorm1d, biased var)
//   out = relu(z * gamma + beta) * drop * [t < len]     drop = keep mask (0 / 1) scaled by 1 / (1 - p), or absent
// One warp per (utterance, channel) row; T is small (text length), the row is read twice from L1/L2.
#include "common.cuh"

namespace rb {

__device__ __forceinline__ float en_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float en_ratio(int t, int len, int ksize) {
  const int half = ksize >> 1;
  const int cnt = 1 + min(half, t) + min(half, len - 1 - t);
  return (float)ksize / ((float)cnt + 1e-6f);
}

__global__ void __launch_bounds__(256) encnorm_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ cbias,
                                                          const int64_t* __restrict__ lens, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ drop,
                                                          float drop_scale, int B, int C, int T, int ksize, float eps,
                                                          float* __restrict__ out, float* __restrict__ mean_out,
                                                          float* __restrict__ rstd_out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= B * C) return;
  const int lane = threadIdx.x & 31;
  const int b = row / C, c = row - b * C;
  const int len = (int)min((long long)lens[b], (long long)T);
  const float* r = raw + (size_t)row * T;
  const float cb = cbias ? cbias[c] : 0.f;
  float s = 0.f;
  for (int t = lane; t < len; t += 32) s += (r[t] - cb) * en_ratio(t, len, ksize) + cb;
  s = en_warp_sum(s);
  const float n = (float)max(len, 1);
  const float mean = s / n;
  float v = 0.f;
  for (int t = lane; t < len; t += 32) {
    const float d = ((r[t] - cb) * en_ratio(t, len, ksize) + cb) - mean;
    v = fmaf(d, d, v);
  }
  v = en_warp_sum(v);
  const float rstd = rsqrtf(v / n + eps);
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  const float g = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  float* o = out + (size_t)row * T;
  const float* dr = drop ? drop + (size_t)row * T : nullptr;
  for (int t = lane; t < T; t += 32) {
    float y = 0.f;
    if (t < len) {
      const float z = (((r[t] - cb) * en_ratio(t, len, ksize) + cb) - mean) * rstd;
      y = fmaxf(fmaf(z, g, be), 0.f);
      if (dr) y *= dr[t] * drop_scale;
    }
    o[t] = y;
  }
}

// g_raw (B, C, T); g_gamma / g_beta / g_cbias (C).  One CTA per channel, its eight warps stride over the utterances and the
// per-channel sums are combined in a fixed order: the conv-bias gradient is analytically ZERO (the instance norm removes any
// per-channel constant), so what is returned is rounding noise and has to be at least reproducible from run to run.
__global__ void __launch_bounds__(256) encnorm_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ cbias,
                                                          const int64_t* __restrict__ lens, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ drop,
                                                          float drop_scale, const float* __restrict__ mean_in,
                                                          const float* __restrict__ rstd_in, const float* __restrict__ g_out,
                                                          int B, int C, int T, int ksize, float* __restrict__ g_raw,
                                                          float* __restrict__ g_gamma, float* __restrict__ g_beta,
                                                          float* __restrict__ g_cbias) {
  __shared__ float part[3][8];
  const int c = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float cb = cbias ? cbias[c] : 0.f;
  const float g = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
  float acc_gamma = 0.f, acc_beta = 0.f, acc_cb = 0.f;
  for (int b = warp; b < B; b += 8) {
    const int row = b * C + c;
    const int len = (int)min((long long)lens[b], (long long)T);
    const float* r = raw + (size_t)row * T;
    const float* go = g_out + (size_t)row * T;
    const float* dr = drop ? drop + (size_t)row * T : nullptr;
    const float mean = mean_in[row], rstd = rstd_in[row];
    // pass 1: sums of g_a, g_a z  (g_a = gradient at the affine output, through dropout and ReLU)
    float s1 = 0.f, s2 = 0.f;
    for (int t = lane; t < len; t += 32) {
      const float z = (((r[t] - cb) * en_ratio(t, len, ksize) + cb) - mean) * rstd;
      float ga = go[t];
      if (dr) ga *= dr[t] * drop_scale;
      if (fmaf(z, g, be) <= 0.f) ga = 0.f;
      s1 += ga;
      s2 = fmaf(ga, z, s2);
    }
    s1 = en_warp_sum(s1);
    s2 = en_warp_sum(s2);
    acc_beta += s1;
    acc_gamma += s2;
    // instance norm: g_y = rstd * gamma * (g_a - mean_t(g_a) - z * mean_t(g_a z)); then the partial-conv epilogue:
    // g_raw = g_y * ratio,  direct bias term g_b += g_y * (1 - ratio)
    const float n = (float)max(len, 1);
    const float m1 = s1 / n, m2 = s2 / n;
    float gb = 0.f;
    float* gr = g_raw + (size_t)row * T;
    for (int t = lane; t < T; t += 32) {
      float v = 0.f;
      if (t < len) {
        const float rt = en_ratio(t, len, ksize);
        const float z = (((r[t] - cb) * rt + cb) - mean) * rstd;
        float ga = go[t];
        if (dr) ga *= dr[t] * drop_scale;
        if (fmaf(z, g, be) <= 0.f) ga = 0.f;
        const float gy = rstd * g * (ga - m1 - z * m2);
        v = gy * rt;
        gb += gy * (1.f - rt);
      }
      gr[t] = v;
    }
    acc_cb += en_warp_sum(gb);
  }
  if (lane == 0) { part[0][warp] = acc_gamma; part[1][warp] = acc_beta; part[2][warp] = acc_cb; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[threadIdx.x][w];
    float* dst = threadIdx.x == 0 ? g_gamma : (threadIdx.x == 1 ? g_beta : g_cbias);
    if (dst) dst[c] = s;
  }
}

}  // namespace rb

using namespace rb;

extern "C" int radtts_encnorm_forward(const float* raw, const float* conv_bias, const int64_t* lens, const float* gamma,
                                      const float* beta, const float* drop, float drop_scale, int B, int C, int T,
                                      int ksize, float eps, float* out, float* mean, float* rstd, void* stream) {
  if (!raw || !lens || !out || !mean || !rstd || B <= 0 || C <= 0 || T <= 0 || ksize <= 0 || ksize % 2 == 0)
    return RADTTS_ERR_INVALID_ARG;
  encnorm_fwd_kernel<<<ceil_div(B * C, 8), 256, 0, (cudaStream_t)stream>>>(raw, conv_bias, lens, gamma, beta, drop, drop_scale,
                                                                         B, C, T, ksize, eps, out, mean, rstd);
  return after_launch();
}

extern "C" int radtts_encnorm_backward(const float* raw, const float* conv_bias, const int64_t* lens, const float* gamma,
                                       const float* beta, const float* drop, float drop_scale, const float* mean,
                                       const float* rstd, const float* g_out, int B, int C, int T, int ksize, float* g_raw,
                                       float* g_gamma, float* g_beta, float* g_conv_bias, void* stream) {
  if (!raw || !lens || !mean || !rstd || !g_out || !g_raw || B <= 0 || C <= 0 || T <= 0 || ksize <= 0 || ksize % 2 == 0)
    return RADTTS_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  encnorm_bwd_kernel<<<C, 256, 0, st>>>(raw, conv_bias, lens, gamma, beta, drop, drop_scale, mean, rstd, g_out,
                                                        B, C, T, ksize, g_raw, g_gamma, g_beta, g_conv_bias);
  return after_launch();
}
