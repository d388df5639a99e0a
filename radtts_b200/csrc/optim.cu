// Fused RAdam over one flat fp32 parameter buffer (the reference's radam.py walks the parameter
// list in Python, ~10 elementwise launches per tensor).  One pass reads p, g, m, v and writes p, m, v.
// Update rule = reference radam.py:76-116 (rectified Adam with the N_sma >= 5 switch, decoupled-style weight decay
// `p -= wd * lr * p`).  The step counter lives on the device so that the whole train step can sit in a CUDA graph;
// `grad_scale` (device scalar, may be NULL) carries the gradient-clipping coefficient.
#include "common.cuh"

namespace rb {

__global__ void __launch_bounds__(256) radam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, size_t n, float lr, float beta1, float beta2,
                                                    float eps, float wd, const long long* __restrict__ step_dev,
                                                    const float* __restrict__ grad_scale) {
  const double t = (double)(step_dev[0] + 1);
  const double b2t = pow((double)beta2, t), b1t = pow((double)beta1, t);
  const double nmax = 2.0 / (1.0 - beta2) - 1.0;
  const double nsma = nmax - 2.0 * t * b2t / (1.0 - b2t);
  const bool rect = nsma >= 5.0;
  const float step_size = rect ? (float)(lr * sqrt((1.0 - b2t) * (nsma - 4.0) / (nmax - 4.0) * (nsma - 2.0) / nsma * nmax /
                                                   (nmax - 2.0)) / (1.0 - b1t))
                               : (float)(lr / (1.0 - b1t));
  const float gs = grad_scale ? grad_scale[0] : 1.f;
  const float decay = 1.f - wd * lr;
  const size_t n4 = n / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pa = &pp.x; const float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = ga[k] * gs;
      va[k] = beta2 * va[k] + (1.f - beta2) * gk * gk;
      ma[k] = beta1 * ma[k] + (1.f - beta1) * gk;
      float x = pa[k] * decay;
      x -= rect ? step_size * ma[k] / (sqrtf(va[k]) + eps) : step_size * ma[k];
      pa[k] = x;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gk = g[i] * gs;
    v[i] = beta2 * v[i] + (1.f - beta2) * gk * gk;
    m[i] = beta1 * m[i] + (1.f - beta1) * gk;
    float x = p[i] * decay;
    x -= rect ? step_size * m[i] / (sqrtf(v[i]) + eps) : step_size * m[i];
    p[i] = x;
  }
}
__global__ void radam_bump_kernel(long long* step_dev) { step_dev[0] += 1; }

}  // namespace rb

using namespace rb;

extern "C" int radtts_radam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, long long* step_dev,
                                 const float* grad_scale, void* stream) {
  if (!p || !g || !m || !v || !step_dev || n == 0) return RADTTS_ERR_INVALID_ARG;
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) return RADTTS_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  radam_kernel<<<kNumSMs * 8, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale);
  RB_TRY(after_launch());
  radam_bump_kernel<<<1, 1, 0, st>>>(step_dev);
  return after_launch();
}
