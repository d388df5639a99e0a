// fp32 kernels for the invertible 1x1 convolution of a flow step on packed rows (reference common.py:407-428: the conv is
// forced to fp32 even under autocast).  The matrix is the C x C weight embedded in the z_ld x z_ld identity (z_ld = 160),
// i.e. 0.1 % of the step's FLOPs -- but as a generic 128x128-tile SIMT GEMM with K = 160 each launch cost ~50 us
// (two waves of latency-bound CTAs, scalar epilogue stores) and its weight gradient ~80 us (split-K atomics with ~100
// adders per address).  Here: W^T stays resident in shared memory, a persistent grid streams 32-row blocks, every
// thread owns a 4 x 4 (rows x columns) register tile; the weight gradient keeps an 8 x 10 tile per thread over a
// contiguous row chunk and issues one red.add per output per CTA.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "rowgemm.cuh"

namespace rb {

constexpr int kInvThreads = 320;   // 40 column groups (4 columns) x 8 row groups
constexpr int kInvTileRows = 10;   // rows per thread: 8 x 10 = 80 rows per pass
constexpr int kInvRows = 8 * kInvTileRows;
constexpr int kInvMaxLd = 160;
constexpr int kInvWtLd = kInvMaxLd + 4;   // padded row of the transposed weight: the transposing stores hit 8 banks, not 1

// y[r][c] = ok(r) ? sum_j w[c][j] x[r][j] : 0      (w row-major [zld][zld])
//   y2 (optional): copy of columns c < c_lim;  z0 (optional, element type T): columns [c_off, c_off + 128) of the result,
//   zero beyond c_off + h
// Every CTA takes ONE contiguous chunk of ceil(rows / grid) rows (74 at the bench shape: a single 80-row pass, no second
// round for a few CTAs) and every thread a 10 x 4 register tile: per four input channels 10 broadcast LDS.128 of x and 4
// LDS.128 of W^T feed 160 FMAs, so the loop is bound by the FMA pipe and not by shared-memory wavefronts (the 4 x 4 tile
// this replaces needed 20 wavefronts per 64 FMAs and ran at 17 % of the fp32 peak: 44 us per launch, 16 launches a step).
template <typename T>
__global__ void __launch_bounds__(kInvThreads, 1)
invconv_rows_kernel(const float* __restrict__ x, const float* __restrict__ w, int zld, const int* __restrict__ plan,
                    int rows_alloc, RowMeta meta, int mask_rows, float* __restrict__ y, float* __restrict__ y2, int c_lim,
                    T* __restrict__ z0, int c_off, int h) {
  extern __shared__ __align__(16) float sm_inv[];
  float* wt = sm_inv;                               // [zld (j)][kInvWtLd (c)]
  float* xs = sm_inv + (size_t)zld * kInvWtLd;      // [80][zld]
  const int rows_used = plan ? plan[0] : rows_alloc;
  int chunk = (rows_used + (int)gridDim.x - 1) / (int)gridDim.x;
  chunk = (chunk + 3) / 4 * 4;
  const int r_begin = blockIdx.x * chunk, r_end = min(r_begin + chunk, rows_used);
  if (r_begin >= r_end) return;
  const int tid = threadIdx.x;
  for (int i = tid; i < zld * zld; i += kInvThreads) {
    const int c = i / zld, j = i - c * zld;
    wt[(size_t)j * kInvWtLd + c] = w[i];
  }
  const int ncg = zld / 4;
  const int cg = tid % 40, rg = tid / 40;
  const bool active = cg < ncg;
  const int zld4 = zld / 4;
  for (int r0 = r_begin; r0 < r_end; r0 += kInvRows) {
    __syncthreads();                           // previous pass's xs fully consumed (and wt staged, first time)
    const int nrows = min(kInvRows, r_end - r0);
    for (int i = tid; i < kInvRows * zld4; i += kInvThreads) {
      const int r = i / zld4, q = i - r * zld4;
      reinterpret_cast<float4*>(xs)[(size_t)r * zld4 + q] =
          r < nrows ? *reinterpret_cast<const float4*>(x + (size_t)(r0 + r) * zld + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (!active || kInvTileRows * rg >= nrows) continue;
    float acc[kInvTileRows][4];
#pragma unroll
    for (int i = 0; i < kInvTileRows; ++i)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[i][k] = 0.f;
    const float* xr = xs + (size_t)(kInvTileRows * rg) * zld;
    for (int j = 0; j < zld; j += 4) {
      float xv[kInvTileRows][4];
#pragma unroll
      for (int i = 0; i < kInvTileRows; ++i) {
        const float4 t = *reinterpret_cast<const float4*>(xr + (size_t)i * zld + j);
        xv[i][0] = t.x; xv[i][1] = t.y; xv[i][2] = t.z; xv[i][3] = t.w;
      }
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float4 w4 = *reinterpret_cast<const float4*>(wt + (size_t)(j + jj) * kInvWtLd + 4 * cg);
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int i = 0; i < kInvTileRows; ++i)
#pragma unroll
          for (int k = 0; k < 4; ++k) acc[i][k] = fmaf(wv[k], xv[i][jj], acc[i][k]);
      }
    }
#pragma unroll
    for (int i = 0; i < kInvTileRows; ++i) {
      const int row = r0 + kInvTileRows * rg + i;
      if (row >= r_end) break;
      const bool ok = !mask_rows || meta.valid(row);
      float v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = ok ? acc[i][k] : 0.f;
      *reinterpret_cast<float4*>(y + (size_t)row * zld + 4 * cg) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = 4 * cg + k;
        if (y2 && c < c_lim) y2[(size_t)row * zld + c] = v[k];
        if (z0 && c >= c_off && c < c_off + 128) {
          const float zv = c < c_off + h ? v[k] : 0.f;
          if (sizeof(T) == 4) reinterpret_cast<float*>(z0)[(size_t)row * 128 + (c - c_off)] = zv;
          else reinterpret_cast<__nv_bfloat16*>(z0)[(size_t)row * 128 + (c - c_off)] = __float2bfloat16(zv);
        }
      }
    }
  }
}

template <typename T>
inline int launch_invconv_rows(const float* x, const float* w, int zld, const int* plan, int rows_alloc, RowMeta meta,
                               int mask_rows, float* y, float* y2, int c_lim, T* z0, int c_off, int h, cudaStream_t st) {
  if (zld % 4 || zld > kInvMaxLd) return RADTTS_ERR_UNSUPPORTED;
  const size_t smem = ((size_t)zld * kInvWtLd + (size_t)kInvRows * zld) * sizeof(float);
  static bool configured = false;
  if (!configured) {
    RB_CUDA(cudaFuncSetAttribute(invconv_rows_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(((size_t)kInvMaxLd * kInvWtLd + (size_t)kInvRows * kInvMaxLd) * sizeof(float))));
    configured = true;
  }
  const int blocks = ceil_div(rows_alloc, 8);      // tiny inputs: at least 8 rows per CTA
  const int grid = blocks < kNumSMs ? blocks : kNumSMs;
  invconv_rows_kernel<T><<<grid, kInvThreads, smem, st>>>(x, w, zld, plan, rows_alloc, meta, mask_rows, y, y2, c_lim, z0,
                                                         c_off, h);
  return after_launch();
}

// out[c][j] += sum_r g[r][c] x[r][j]   (zld == 160; out zeroed by the caller).  Thread tile 8 (c) x 10 (j).
constexpr int kInvWgRows = 16;
static __global__ void __launch_bounds__(kInvThreads, 1)
invconv_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ x, const int* __restrict__ plan, int rows_alloc,
                     float* __restrict__ out) {
  constexpr int LD = kInvMaxLd;
  __shared__ __align__(16) float gs[kInvWgRows][LD];
  __shared__ __align__(16) float xs[kInvWgRows][LD];
  const int rows_used = plan ? plan[0] : rows_alloc;
  int chunk = (rows_used + gridDim.x - 1) / gridDim.x;
  chunk = (chunk + kInvWgRows - 1) / kInvWgRows * kInvWgRows;
  const int r_begin = blockIdx.x * chunk, r_end = min(r_begin + chunk, rows_used);
  if (r_begin >= r_end) return;
  const int tid = threadIdx.x;
  const int cgp = tid % 20, jgp = tid / 20;    // columns 8 cgp .. +7 of g, columns 10 jgp .. +9 of x
  float acc[8][10];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 10; ++k) acc[i][k] = 0.f;
  for (int r0 = r_begin; r0 < r_end; r0 += kInvWgRows) {
    __syncthreads();
    for (int i = tid; i < kInvWgRows * (LD / 4); i += kInvThreads) {
      const int r = i / (LD / 4), q = i - r * (LD / 4);
      const bool in = r0 + r < r_end;
      const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
      reinterpret_cast<float4*>(&gs[r][0])[q] = in ? *reinterpret_cast<const float4*>(g + (size_t)(r0 + r) * LD + 4 * q) : z4;
      reinterpret_cast<float4*>(&xs[r][0])[q] = in ? *reinterpret_cast<const float4*>(x + (size_t)(r0 + r) * LD + 4 * q) : z4;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < kInvWgRows; ++r) {
      const float4 g0 = *reinterpret_cast<const float4*>(&gs[r][8 * cgp]);
      const float4 g1 = *reinterpret_cast<const float4*>(&gs[r][8 * cgp + 4]);
      const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float2* xp = reinterpret_cast<const float2*>(&xs[r][10 * jgp]);
      float xv[10];
#pragma unroll
      for (int k = 0; k < 5; ++k) { const float2 t = xp[k]; xv[2 * k] = t.x; xv[2 * k + 1] = t.y; }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int k = 0; k < 10; ++k) acc[i][k] = fmaf(gv[i], xv[k], acc[i][k]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 10; ++k) atomicAdd(out + (size_t)(8 * cgp + i) * LD + 10 * jgp + k, acc[i][k]);
}

inline int launch_invconv_wgrad(const float* g, const float* x, const int* plan, int rows_alloc, float* out, cudaStream_t st) {
  // ~64 rows per CTA at the bench shape: enough CTAs to fill the machine, few enough that each output sees <= 148 adds
  const int grid = kNumSMs;
  invconv_wgrad_kernel<<<grid, kInvThreads, 0, st>>>(g, x, plan, rows_alloc, out);
  return after_launch();
}

}  // namespace rb
