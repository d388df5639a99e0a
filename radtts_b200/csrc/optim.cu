// Fused RAdam over one flat fp32 parameter buffer (the reference's radam.py walks the parameter
// list in Python, ~10 elementwise launches per tensor).  One pass reads p, g, m, v and writes p, m, v.
// Update rule = reference radam.py:76-116 (rectified Adam with the N_sma >= 5 switch, decoupled-style weight decay
// `p -= wd * lr * p`).  The step counter lives on the device so that the whole train step can sit in a CUDA graph;
// `grad_scale` (device scalar, may be NULL) carries the gradient-clipping coefficient.
#include <cstdlib>

#include "common.cuh"

namespace rb {

__global__ void __launch_bounds__(256) radam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, size_t n, float lr, float beta1, float beta2,
                                                    float eps, float wd, const long long* __restrict__ step_dev,
                                                    const float* __restrict__ grad_scale, const int* __restrict__ enable,
                                                    int zero_grad) {
  if (enable && enable[0] == 0) return;
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
  // Two independent float4 quadruples per thread and iteration (8 x 16-byte loads in flight): the grid is kept to 4 CTAs
  // per SM so that a deferred update leaves half of every SM's thread slots to the kernels it runs underneath, and the
  // bytes in flight (148 x 4 x 256 x 8 x 16 B = 19 MB) still cover HBM latency x bandwidth.
  auto update4 = [&](float4& pp, const float4& gg, float4& mm, float4& vv) {
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
  };
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* g4 = reinterpret_cast<float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {
    const size_t j = i + stride;
    float4 pa = p4[i], pb = p4[j];
    const float4 ga = g4[i], gb = g4[j];
    float4 ma = m4[i], mb = m4[j];
    float4 va = v4[i], vb = v4[j];
    update4(pa, ga, ma, va);
    update4(pb, gb, mb, vb);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
    p4[j] = pb; m4[j] = mb; v4[j] = vb;
    if (zero_grad) { g4[i] = zero4; g4[j] = zero4; }
  }
  if (i < n4) {
    float4 pa = p4[i];
    const float4 ga = g4[i];
    float4 ma = m4[i], va = v4[i];
    update4(pa, ga, ma, va);
    p4[i] = pa; m4[i] = ma; v4[i] = va;
    if (zero_grad) g4[i] = zero4;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gk = g[i] * gs;
    v[i] = beta2 * v[i] + (1.f - beta2) * gk * gk;
    m[i] = beta1 * m[i] + (1.f - beta1) * gk;
    float x = p[i] * decay;
    x -= rect ? step_size * m[i] / (sqrtf(v[i]) + eps) : step_size * m[i];
    p[i] = x;
    if (zero_grad) g[i] = 0.f;
  }
}
__global__ void radam_bump_kernel(long long* step_dev, const int* __restrict__ enable) {
  if (enable && enable[0] == 0) return;
  step_dev[0] += 1;
}

}  // namespace rb

using namespace rb;

extern "C" int radtts_radam_step_ex(float* p, float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                                    float eps, float weight_decay, long long* step_dev, const float* grad_scale,
                                    const int* enable, int zero_grad, int bump, void* stream) {
  if (!p || !g || !m || !v || !step_dev || n == 0) return RADTTS_ERR_INVALID_ARG;
  if (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) return RADTTS_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // A pending (pipelined) update runs UNDERNEATH other kernels: 2 CTAs per SM (35 K registers, 512 threads) leave room for
  // them on every SM; an update nothing can overlap takes 6 CTAs per SM for the bytes in flight HBM wants.
  static const int kPipelinedCtas = [] { const char* e = std::getenv("RADTTS_RADAM_PIPELINED_CTAS"); return e ? atoi(e) : 2; }();
  // An SM changes its L1 / shared-memory split only when it is empty.  This kernel uses no shared memory, so by default its
  // CTAs would configure every SM for the largest L1 -- and every kernel of the step's front end that needs a few KB of
  // shared memory would wait for the whole update to drain (measured: the first front-end kernel started 1 ms late).  Ask
  // for the largest shared-memory carve-out instead; a streaming kernel has no use for L1.
  static bool carveout_set = false;
  if (!carveout_set) {
    RB_CUDA(cudaFuncSetAttribute(radam_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    carveout_set = true;
  }
  static const int kPipelinedGrid = [] { const char* e = std::getenv("RADTTS_RADAM_PIPELINED_GRID"); return e ? atoi(e) : 0; }();
  const int ctas_per_sm = enable ? kPipelinedCtas : 6;
  const int grid = (enable && kPipelinedGrid > 0) ? kPipelinedGrid : kNumSMs * ctas_per_sm;
  radam_kernel<<<grid, 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale,
                                                      enable, zero_grad);
  RB_TRY(after_launch());
  if (bump) {
    radam_bump_kernel<<<1, 1, 0, st>>>(step_dev, enable);
    RB_TRY(after_launch());
  }
  return 0;
}

extern "C" int radtts_radam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, long long* step_dev,
                                 const float* grad_scale, void* stream) {
  return radtts_radam_step_ex(p, const_cast<float*>(g), m, v, n, lr, beta1, beta2, eps, weight_decay, step_dev, grad_scale,
                              nullptr, 0, 1, stream);
}
