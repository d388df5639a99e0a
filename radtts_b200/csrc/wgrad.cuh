// Weight-gradient contraction over the packed row axis:
//
//     out[n * so_n + c * so_c] += sum_r  G[r, gcol + n] * X[r + shift, xcol + c]        n < N, c < C
//
// (G = gradient w.r.t. a conv's pre-activation output, X = that conv's input; one call per tap, `shift` is the
// tap offset, so_n / so_c let the result land directly in PyTorch's (out, in, tap) weight layout.)
// K is the row axis, so both operands are "MN-major" (channels contiguous, rows strided).
//   * wgrad_simt  -- fp32 accumulate SIMT engine (operands fp32 or bf16), split over row chunks + atomics
//   * wgrad_tc    -- tcgen05 engine with MN-major smem descriptors (wgrad_tc.cuh)
// plus the bias-gradient column sums.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "rowgemm.cuh"

namespace rb {

struct WgradProb {
  const void* G; int ldg, gcol, N;
  const void* X; int ldx, xcol, C;
  int shift;
  float* out; long so_n; int so_c;
};

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]) { Act<T>::template ldv<8>(p, v); }

template <typename TG, typename TX>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradProb p, const int* __restrict__ plan, int rows_alloc,
                                                         int chunk) {
  const int rows_used = plan ? plan[0] : rows_alloc;
  const int tiles_c = (p.C + 127) / 128;
  const int n0 = (blockIdx.x / tiles_c) * 128, c0 = (blockIdx.x % tiles_c) * 128;
  const int r_begin = blockIdx.y * chunk;
  const int r_end = min(r_begin + chunk, rows_used);
  if (r_begin >= r_end) return;

  __shared__ __align__(16) float Gs[2][16][132];
  __shared__ __align__(16) float Xs[2][16][132];
  const int tid = threadIdx.x;
  const int lrow = tid >> 4, lcol = (tid & 15) * 8;
  const int ty = tid >> 4, tx = tid & 15;
  const TG* G = reinterpret_cast<const TG*>(p.G);
  const TX* X = reinterpret_cast<const TX*>(p.X);
  const bool g_ok = p.gcol + n0 + lcol + 8 <= p.ldg;
  const bool x_ok = p.xcol + c0 + lcol + 8 <= p.ldx;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float rg[8], rx[8];
  auto fetch = [&](int r) {
    const int gr = r + lrow;
    const int xr = gr + p.shift;
    if (g_ok && gr < r_end) load8<TG>(G + (size_t)gr * p.ldg + p.gcol + n0 + lcol, rg);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) rg[i] = 0.f;
    }
    if (x_ok && gr < r_end && xr >= 0 && xr < rows_used) load8<TX>(X + (size_t)xr * p.ldx + p.xcol + c0 + lcol, rx);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) rx[i] = 0.f;
    }
  };
  auto stage = [&](int buf) {
    *reinterpret_cast<float4*>(&Gs[buf][lrow][lcol]) = make_float4(rg[0], rg[1], rg[2], rg[3]);
    *reinterpret_cast<float4*>(&Gs[buf][lrow][lcol + 4]) = make_float4(rg[4], rg[5], rg[6], rg[7]);
    *reinterpret_cast<float4*>(&Xs[buf][lrow][lcol]) = make_float4(rx[0], rx[1], rx[2], rx[3]);
    *reinterpret_cast<float4*>(&Xs[buf][lrow][lcol + 4]) = make_float4(rx[4], rx[5], rx[6], rx[7]);
  };
  fetch(r_begin);
  stage(0);
  __syncthreads();
  int buf = 0;
  for (int r = r_begin; r < r_end; r += 16, buf ^= 1) {
    const bool more = r + 16 < r_end;
    if (more) fetch(r + 16);
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&Gs[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&Gs[buf][kk][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Xs[buf][kk][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Xs[buf][kk][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      stage(buf ^ 1);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int n = n0 + ty * 8 + i;
    if (n >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + tx * 8 + j;
      if (c < p.C) atomicAdd(p.out + (size_t)n * p.so_n + (size_t)c * p.so_c, acc[i][j]);
    }
  }
}

template <typename TG, typename TX>
inline int launch_wgrad_simt(const WgradProb& p, const int* plan, int rows_alloc, cudaStream_t st) {
  const int tiles = ceil_div(p.N, 128) * ceil_div(p.C, 128);
  int nsplit = ceil_div(4 * kNumSMs, tiles);
  const int max_split = ceil_div(rows_alloc, 256);
  if (nsplit > max_split) nsplit = max_split;
  if (nsplit < 1) nsplit = 1;
  const int chunk = round_up(ceil_div(rows_alloc, nsplit), 16);
  dim3 grid(tiles, ceil_div(rows_alloc, chunk));
  wgrad_simt_kernel<TG, TX><<<grid, 256, 0, st>>>(p, plan, rows_alloc, chunk);
  return after_launch();
}

// out[n * ostride] += sum_r G[r, gcol + n] / (partial ? ratio(r) : 1)
// Block = 128 columns x 8 row lanes, 4 consecutive columns per thread (one 8/16-byte load), 256-row chunks:
// every warp reads whole 256/512-byte row segments, so the pass streams at HBM/L2 speed.
template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 t = *reinterpret_cast<const uint2*>(p);
  v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
  v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
}

constexpr int kColsumMaxItems = 2 * RADTTS_MAX_LAYERS + 2;
struct ColsumItem {
  const void* G;
  float* out;
  int ldg, gcol, N, partial, log2d, ostride;
};
struct ColsumParams {
  ColsumItem it[kColsumMaxItems];
  RowMeta meta;
  const int* plan;
  int rows_alloc, chunk, ksize;
};
// All bias-gradient column sums of a flow in one launch: blockIdx.z = problem.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const ColsumParams p) {
  const ColsumItem& it = p.it[blockIdx.z];
  const int N = it.N, ldg = it.ldg;
  if ((int)blockIdx.x * 128 >= N) return;
  const T* G = reinterpret_cast<const T*>(it.G);
  const int rows_used = p.plan ? p.plan[0] : p.rows_alloc;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 128 + tx * 4;
  const int r_begin = blockIdx.y * p.chunk, r_end = min(r_begin + p.chunk, rows_used);
  float s[4] = {0.f, 0.f, 0.f, 0.f};
  if (n0 < N) {   // N is a multiple of 4 for every caller (channel counts are multiples of 16)
#pragma unroll 4
    for (int r = r_begin + ty; r < r_end; r += 8) {
      float v[4];
      load4<T>(G + (size_t)r * ldg + it.gcol + n0, v);
      float sc = 1.f;
      if (it.partial) sc = p.meta.valid(r) ? 1.f / p.meta.ratio(r, it.log2d, p.ksize) : 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) s[i] = fmaf(v[i], sc, s[i]);
    }
  }
  __shared__ float red[8][128 + 4];
#pragma unroll
  for (int i = 0; i < 4; ++i) red[ty][tx * 4 + i] = s[i];
  __syncthreads();
  if (threadIdx.x < 128 && r_begin < r_end) {
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n < N) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
      atomicAdd(it.out + (size_t)n * it.ostride, t);
    }
  }
}

// Collects the column-sum problems of a flow; launch() runs them all at once.
template <typename T>
struct ColsumBatch {
  ColsumParams p{};
  int n = 0, max_n = 0;
  int add(const T* G, int ldg, int gcol, int N, int partial, int log2d, float* out, int ostride, bool zero_first,
          cudaStream_t st) {
    if (n == kColsumMaxItems || (N % 4) || (ldg % 4) || (gcol % 4)) return RADTTS_ERR_INVALID_ARG;
    if (zero_first) RB_CUDA(cudaMemsetAsync(out, 0, (size_t)N * ostride * sizeof(float), st));
    p.it[n++] = ColsumItem{G, out, ldg, gcol, N, partial, log2d, ostride};
    if (N > max_n) max_n = N;
    return 0;
  }
  int launch(RowMeta meta, int ksize, const int* plan, int rows_alloc, cudaStream_t st) {
    if (n == 0) return 0;
    p.meta = meta; p.plan = plan; p.rows_alloc = rows_alloc; p.chunk = 256; p.ksize = ksize;
    dim3 grid(ceil_div(max_n, 128), ceil_div(rows_alloc, p.chunk), n);
    colsum_kernel<T><<<grid, 256, 0, st>>>(p);
    return after_launch();
  }
};

}  // namespace rb

#include "wgrad_tc.cuh"

namespace rb {

// Dispatch: fp32 operands -> SIMT; bf16 operands -> tcgen05.
template <typename TG, typename TX>
inline int launch_wgrad(const WgradProb& p, const int* plan, int rows_alloc, bool zero_first, cudaStream_t st) {
  if (zero_first) RB_CUDA(cudaMemsetAsync(p.out, 0, (size_t)p.N * p.so_n * sizeof(float), st));
  if constexpr (sizeof(TG) == 2 && sizeof(TX) == 2) {
    if (wgrad_tc_supported(p)) return launch_wgrad_tc(p, plan, rows_alloc, st);
  }
  return launch_wgrad_simt<TG, TX>(p, plan, rows_alloc, st);
}

}  // namespace rb
