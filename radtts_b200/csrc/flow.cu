// Kernel 2 (host orchestration + small kernels): frame plan, pack / unpack, weight re-layout, and the
// decoder flow step (invertible 1x1 conv -> WN parameter net -> affine coupling) in both directions.
// The contractions run through rowgemm_simt (fp32) or rowgemm_tc (bf16 tcgen05); see rowgemm.cuh.
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "flow_common.cuh"
#include "flow_layout.cuh"
#include "invconv.cuh"
#include "frameplan.cuh"
#include "rowgemm.cuh"
#include "rowgemm_tc.cuh"

namespace rb {

// ====================================================================================================
// frame plan
// ====================================================================================================
__global__ void __launch_bounds__(1024) frameplan_kernel(const int64_t* __restrict__ lens, const int64_t* __restrict__ geom_lens,
                                                         int divisor, int B, int Tmax, int rows_alloc,
                                                         int* __restrict__ plan) {
  int* hdr = plan;
  int* row0 = plan + 8;
  int* len = plan + 8 + B;
  int* glen = plan + 8 + 2 * B;
  int* pos = plan + plan_meta_off(B);
  int* rem = pos + rows_alloc;
  int* utt = rem + rows_alloc;
  int* tpos = utt + rows_alloc;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int r = kGap;
    for (int b = 0; b < B; ++b) {
      long long l = lens[b] / divisor;
      int li = (int)(l < 0 ? 0 : (l > Tmax ? Tmax : l));
      long long gl = geom_lens ? geom_lens[b] / divisor : l;
      int gi = (int)(gl < li ? li : (gl > Tmax ? Tmax : gl));
      row0[b] = r;
      len[b] = li;
      glen[b] = gi;
      r += gi + kGap;
    }
    hdr[0] = r;  // rows in use (ends with a gap)
    hdr[1] = B; hdr[2] = Tmax; hdr[3] = rows_alloc; hdr[4] = hdr[5] = hdr[6] = hdr[7] = 0;
  }
  for (int i = tid; i < rows_alloc; i += blockDim.x) { pos[i] = -1; rem[i] = -1; utt[i] = -1; tpos[i] = -1; }
  __syncthreads();
  for (int b = 0; b < B; ++b) {
    const int r0 = row0[b], l = len[b], gl = glen[b];
    for (int t = tid; t < gl; t += blockDim.x) {
      utt[r0 + t] = b;
      tpos[r0 + t] = t;
      if (t < l) { pos[r0 + t] = t; rem[r0 + t] = l - 1 - t; }
    }
  }
}

// ====================================================================================================
// pack / unpack: (B, C, T) time-contiguous  <->  packed rows, channel-contiguous (32 x 32 smem transpose)
// ====================================================================================================
template <typename T>
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ src, int C, int Tsrc, int g,
                                                   PlanView pv, T* __restrict__ dst, int ld, int col_off,
                                                   int ncols_pad, int valid_only) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;  // all rows_alloc rows are written (zeros off-range)
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ncols = C * g;
  {
    const int row = r0 + tx;
    int b = row < pv.rows_alloc ? pv.utt()[row] : -1;
    if (valid_only && b >= 0 && pv.pos()[row] < 0) b = -1;  // inside the geometric span but past the valid length
    const int t = b >= 0 ? pv.tpos()[row] : 0;
    for (int cc = ty; cc < 32; cc += 8) {
      const int col = c0 + cc;
      float v = 0.f;
      if (b >= 0 && col < ncols) {
        const int c = col / g, k = col - c * g;
        const int ts = g * t + k;
        if (ts < Tsrc) v = src[((size_t)b * C + c) * Tsrc + ts];
      }
      tile[cc][tx] = v;
    }
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int row = r0 + rr, col = c0 + tx;
    if (row < pv.rows_alloc && col < ncols_pad) {
      const float v = tile[tx][rr];
      if (sizeof(T) == 4) reinterpret_cast<float*>(dst)[(size_t)row * ld + col_off + col] = v;
      else reinterpret_cast<__nv_bfloat16*>(dst)[(size_t)row * ld + col_off + col] = __float2bfloat16(v);
    }
  }
}

__global__ void __launch_bounds__(256) unpack_kernel(const float* __restrict__ src, int ld, int col_off, PlanView pv,
                                                     int C, int g, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ncols = C * g;
  const int len = pv.glen()[b], row0 = pv.row0()[b];  // geometric span (== valid length unless geom_lens given)
  const int Tout = pv.Tmax * g;
  for (int rr = ty; rr < 32; rr += 8) {
    const int t = t0 + rr, col = c0 + tx;
    float v = 0.f;
    if (t < len && col < ncols) v = src[(size_t)(row0 + t) * ld + col_off + col];
    tile[rr][tx] = v;
  }
  __syncthreads();
  for (int cc = ty; cc < 32; cc += 8) {
    const int col = c0 + cc, t = t0 + tx;
    if (col < ncols && t < pv.Tmax) {
      const int c = col / g, k = col - c * g;
      dst[((size_t)b * C + c) * Tout + g * t + k] = tile[tx][cc];
    }
  }
}

// ====================================================================================================
// weight re-layout kernels (run once per optimizer step / once per model)
// ====================================================================================================
template <typename T>
__device__ __forceinline__ void put(T* p, float v);
template <>
__device__ __forceinline__ void put<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void put<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// start weight (N, h + n_ctx) [z0 | ctx] -> dst[n][ctx (pad ctx_ld) | z0 (pad 128)]
template <typename T>
__global__ void prep_start_kernel(const float* __restrict__ w, const float* __restrict__ scale, int N, int h, int n_ctx,
                                  int ctx_ld, T* __restrict__ dst) {
  const int ldw = ctx_ld + 128;
  const size_t total = (size_t)N * ldw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % ldw);
    const int n = (int)(i / ldw);
    float v = 0.f;
    if (j < ctx_ld) { if (j < n_ctx) v = w[(size_t)n * (h + n_ctx) + h + j]; }
    else if (j - ctx_ld < h) v = w[(size_t)n * (h + n_ctx) + (j - ctx_ld)];
    put(dst + i, v * scale[n]);
  }
}

// ---- weight norm folded into the re-layout (torch._weight_norm, dim 0: w[n] = v[n] * g[n] / ||v[n]||) ----
constexpr int kMaxWnItems = 1 + 2 * RADTTS_MAX_LAYERS;
struct WnScaleParams {
  const float* v[kMaxWnItems];
  const float* g[kMaxWnItems];   // NULL: the tensor is an effective weight already, scale = 1
  int ck[kMaxWnItems];           // elements per output channel
  int n_items, nc;
  float* scales;                 // [n_items][nc]
};
// one warp per (tensor, output channel)
__global__ void __launch_bounds__(256) wn_scale_kernel(const WnScaleParams p) {
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gw >= p.n_items * p.nc) return;
  const int item = gw / p.nc, n = gw - item * p.nc;
  float s = 1.f;
  if (p.g[item]) {
    const float* v = p.v[item] + (size_t)n * p.ck[item];
    float acc = 0.f;
    for (int i = lane; i < p.ck[item]; i += 32) acc = fmaf(v[i], v[i], acc);
#pragma unroll
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    s = p.g[item][n] / sqrtf(acc);
  }
  if (lane == 0) p.scales[gw] = s;
}

// All k-tap / 1x1 convs of a WN in ONE launch: (N, C, k) fp32 -> dst[n][t * C + c] (tap-major K, scaled) and, for the
// backward pass, the transposed copy the fused dgrad GEMMs read: dstT[c][dcolT + t * N + n].
struct PrepConvItem {
  const float* v;
  const float* scale;
  void* dst;
  void* dstT;   // may be NULL
  int k, lddT, dcolT, pad_;
};
struct PrepConvParams {
  PrepConvItem it[2 * RADTTS_MAX_LAYERS];
  int N, C;
};
// K = tap count as a template parameter: the index arithmetic below divides by it ~40 times per thread
template <typename T, int K>
__global__ void __launch_bounds__(256) prep_convs_kernel(const PrepConvParams p) {
  extern __shared__ float tile[];                  // [k][32][65], tap planes kPlane apart (odd mod 32: the staging
                                                   // loop walks taps fastest and would hit one bank k times otherwise)
  constexpr int kPlane = 32 * 65 + 7;
  const PrepConvItem& it = p.it[blockIdx.z];
  constexpr int k = K;
  const int N = p.N, C = p.C;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
  // Whole tiles with 16-byte aligned rows (always the case for the WN convs: N, C multiples of 64) move as 16-byte
  // vectors: float4 loads, 8 x bf16 / 4 x fp32 stores.  Anything else takes the element-wise path.
  const bool vec = (n0 + 32 <= N) && (c0 + 64 <= C) && (C % 8 == 0) && (N % 8 == 0) && ((it.lddT | it.dcolT) % 8 == 0);
  if (vec) {
    constexpr int kQ = 16 * K;            // float4 per weight row of this tile (64 channels x K taps, contiguous)
    for (int idx = threadIdx.x; idx < 32 * kQ; idx += 256) {
      const int n = idx / kQ, q = idx - n * kQ;
      const float sc = it.scale[n0 + n];
      const float4 v4 = *reinterpret_cast<const float4*>(it.v + ((size_t)(n0 + n) * C + c0) * K + 4 * q);
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = 4 * q + j, c = r / K, t = r - c * K;
        tile[t * kPlane + n * 65 + c] = vv[j] * sc;
      }
    }
    __syncthreads();
    T* dst = reinterpret_cast<T*>(it.dst);
    const size_t ldw = (size_t)K * C;
    constexpr int kPer = 16 / sizeof(T);  // elements per 16-byte store
    // forward layout: (t, n) rows of 64 consecutive channels
    for (int idx = threadIdx.x; idx < K * 32 * (64 / kPer); idx += 256) {
      const int cc = idx % (64 / kPer), n = (idx / (64 / kPer)) & 31, t = idx / (32 * (64 / kPer));
      float vals[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) vals[j] = tile[t * kPlane + n * 65 + cc * kPer + j];
      Act<T>::template stv<kPer>(dst + (size_t)(n0 + n) * ldw + (size_t)t * C + c0 + cc * kPer, vals);
    }
    if (it.dstT) {
      // transposed layout: (t, c) rows of 32 consecutive output channels
      T* dstT = reinterpret_cast<T*>(it.dstT);
      for (int idx = threadIdx.x; idx < K * 64 * (32 / kPer); idx += 256) {
        const int nn = idx % (32 / kPer), c = (idx / (32 / kPer)) & 63, t = idx / (64 * (32 / kPer));
        float vals[kPer];
#pragma unroll
        for (int j = 0; j < kPer; ++j) vals[j] = tile[t * kPlane + (nn * kPer + j) * 65 + c];
        Act<T>::template stv<kPer>(dstT + (size_t)(c0 + c) * it.lddT + it.dcolT + (size_t)t * N + n0 + nn * kPer, vals);
      }
    }
    return;
  }
  const int per_row = 64 * k;
  for (int idx = threadIdx.x; idx < 32 * per_row; idx += 256) {
    const int n = idx / per_row, r = idx - n * per_row;
    const int c = r / k, t = r - c * k;
    float v = 0.f;
    if (n0 + n < N && c0 + c < C) v = it.v[((size_t)(n0 + n) * C + c0 + c) * k + t] * it.scale[n0 + n];
    tile[t * kPlane + n * 65 + c] = v;
  }
  __syncthreads();
  T* dst = reinterpret_cast<T*>(it.dst);
  const size_t ldw = (size_t)k * C;
  for (int idx = threadIdx.x; idx < k * 32 * 64; idx += 256) {
    const int c = idx & 63, n = (idx >> 6) & 31, t = idx >> 11;
    if (n0 + n < N && c0 + c < C) put(dst + (size_t)(n0 + n) * ldw + (size_t)t * C + c0 + c, tile[t * kPlane + n * 65 + c]);
  }
  if (it.dstT) {
    T* dstT = reinterpret_cast<T*>(it.dstT);
    for (int idx = threadIdx.x; idx < k * 64 * 32; idx += 256) {
      const int n = idx & 31, c = (idx >> 5) & 63, t = idx >> 11;
      if (n0 + n < N && c0 + c < C)
        put(dstT + (size_t)(c0 + c) * it.lddT + it.dcolT + (size_t)t * N + n0 + n, tile[t * kPlane + n * 65 + c]);
    }
  }
}
struct CopyVecParams {
  const float* src[kMaxWnItems];
  float* dst[kMaxWnItems];
  int n;
};
__global__ void copy_vecs_kernel(const CopyVecParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p.n) p.dst[blockIdx.y][i] = p.src[blockIdx.y][i];
}
// end weight (2h, n_ch) -> dst[2c + p][l * n_ch + k] replicated over the n_layers K blocks; rows >= 2h zero
template <typename T>
__global__ void prep_end_kernel(const float* __restrict__ w, const float* __restrict__ b, int h, int n_ch,
                                int n_layers, int nrows, T* __restrict__ dst, float* __restrict__ bdst) {
  const int ldw = n_layers * n_ch;
  const size_t total = (size_t)nrows * ldw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % ldw);
    const int r = (int)(i / ldw);
    const int c = r >> 1, p = r & 1;
    float v = 0.f;
    if (c < h) v = w[(size_t)(p * h + c) * n_ch + (kk % n_ch)];
    put(dst + i, v);
    if (kk == 0) bdst[r] = c < h ? b[p * h + c] : 0.f;
  }
}
// dst[col0_dst + r][...]: transposed window copy: dst[j][dcol + n] = src[n][scol + j], n < N, j < ncols
template <typename T>
__global__ void prep_transpose_kernel(const T* __restrict__ src, int lds, int scol, int N, int ncols,
                                      T* __restrict__ dst, int ldd, int dcol) {
  __shared__ T tile[32][33];
  const int n0 = blockIdx.x * 32, j0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const int n = n0 + rr, j = j0 + tx;
    if (n < N && j < ncols) tile[rr][tx] = src[(size_t)n * lds + scol + j];
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int j = j0 + rr, n = n0 + tx;
    if (n < N && j < ncols) dst[(size_t)j * ldd + dcol + n] = tile[tx][rr];
  }
}
template <typename T>
__global__ void fill_zero_kernel(T* p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    put(p + i, 0.f);
}
// embeds W (C x C) into the z_ld x z_ld identity (exited channels pass through) and also writes its transpose
__global__ void prep_inv_kernel(const float* __restrict__ w, int zld, int c_off, float* __restrict__ full,
                                float* __restrict__ full_t) {
  const int C = zld - c_off;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zld * zld; i += gridDim.x * blockDim.x) {
    const int r = i / zld, c = i % zld;
    float v;
    if (r >= c_off && c >= c_off) v = w[(size_t)(r - c_off) * C + (c - c_off)];
    else v = (r == c) ? 1.f : 0.f;
    full[(size_t)r * zld + c] = v;
    if (full_t) full_t[(size_t)c * zld + r] = v;
  }
}
// inverse direction prologue: z0 copy for `start`, and zmid[:, < c_off + h] = zin
template <typename T>
__global__ void extract_z0_kernel(const float* __restrict__ zin, int zld, int c_off, int h, const int* __restrict__ plan,
                                  float* __restrict__ zmid, T* __restrict__ z0) {
  const int rows_used = plan[0];
  const size_t total = (size_t)rows_used * zld;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % zld);
    const size_t row = i / zld;
    const float v = zin[i];
    if (c < c_off + h) zmid[i] = v;
    if (c >= c_off && c < c_off + 128) put(z0 + row * 128 + (c - c_off), c < c_off + h ? v : 0.f);
  }
}

// ====================================================================================================
// prepare
// ====================================================================================================
template <typename T>
static int prepare_impl(const radtts_flow_dims& d, const radtts_flow_weights& w, int want_backward, uint8_t* base,
                        const FlowLayout& L, cudaStream_t st) {
  const int h = d.c_active / 2, nc = d.n_ch, k = d.ksize, nl = d.n_layers;
  const int ldstart = L.ctx_ld + 128;
  float* invf = reinterpret_cast<float*>(base + L.w_inv);
  float* invt = want_backward ? reinterpret_cast<float*>(base + L.w_inv_t) : nullptr;
  prep_inv_kernel<<<grid_for((size_t)d.z_ld * d.z_ld), 256, 0, st>>>(w.w_inv, d.z_ld, d.c_off, invf, invt);
  RB_TRY(after_launch());
  // weight-norm scales g / ||v|| of every conv (1 where the caller passed an effective weight)
  float* scales = reinterpret_cast<float*>(base + L.scales);
  {
    WnScaleParams sp{};
    sp.n_items = 1 + 2 * nl;
    sp.nc = nc;
    sp.scales = scales;
    sp.v[0] = w.w_start; sp.g[0] = w.wg_start; sp.ck[0] = h + d.n_ctx;
    for (int i = 0; i < nl; ++i) {
      sp.v[1 + i] = w.w_in[i]; sp.g[1 + i] = w.wg_in[i]; sp.ck[1 + i] = nc * k;
      sp.v[1 + nl + i] = w.w_rs[i]; sp.g[1 + nl + i] = w.wg_rs[i]; sp.ck[1 + nl + i] = nc;
    }
    wn_scale_kernel<<<ceil_div(sp.n_items * nc, 8), 256, 0, st>>>(sp);
    RB_TRY(after_launch());
  }
  T* ws = reinterpret_cast<T*>(base + L.w_start);
  prep_start_kernel<T><<<grid_for((size_t)nc * ldstart), 256, 0, st>>>(w.w_start, scales, nc, h, d.n_ctx, L.ctx_ld, ws);
  RB_TRY(after_launch());
  {
    CopyVecParams cp{};
    cp.n = nc;
    cp.src[0] = w.b_start; cp.dst[0] = reinterpret_cast<float*>(base + L.b_start);
    for (int i = 0; i < nl; ++i) {
      cp.src[1 + i] = w.b_in[i]; cp.dst[1 + i] = reinterpret_cast<float*>(base + L.b_in[i]);
      cp.src[1 + nl + i] = w.b_rs[i]; cp.dst[1 + nl + i] = reinterpret_cast<float*>(base + L.b_rs[i]);
    }
    copy_vecs_kernel<<<dim3(ceil_div(nc, 256), 1 + 2 * nl), 256, 0, st>>>(cp);
    RB_TRY(after_launch());
  }
  {
    // in_layers and res_skip layers, forward layout + (want_backward) the dgrad layout, one launch:
    //   dgrad GEMM of layer i reads [ rs_i^T | in_{i+1}^T taps ] (w_dg[i]); the first conv's transpose is w_dg0
    PrepConvParams pp{};
    pp.N = nc; pp.C = nc;
    for (int i = 0; i < nl; ++i) {
      PrepConvItem& a = pp.it[i];
      a.v = w.w_in[i]; a.scale = scales + (size_t)(1 + i) * nc; a.dst = base + L.w_in[i]; a.k = k;
      if (want_backward) {
        if (i == 0) { a.dstT = base + L.w_dg0; a.lddT = k * nc; a.dcolT = 0; }
        else { a.dstT = base + L.w_dg[i - 1]; a.lddT = L.dg_k[i - 1]; a.dcolT = nc; }
      }
      PrepConvItem& r = pp.it[nl + i];
      r.v = w.w_rs[i]; r.scale = scales + (size_t)(1 + nl + i) * nc; r.dst = base + L.w_rs[i]; r.k = 1;
      if (want_backward) { r.dstT = base + L.w_dg[i]; r.lddT = L.dg_k[i]; r.dcolT = 0; }
    }
    // two launches (k taps / 1 tap) so that the 1x1 items do not pay for the k-tap tile
    const size_t smem_k = (size_t)k * (32 * 65 + 7) * sizeof(float), smem_1 = (size_t)(32 * 65 + 7) * sizeof(float);
    if (smem_k > 48 * 1024) return RADTTS_ERR_UNSUPPORTED;
    if (k == 5) prep_convs_kernel<T, 5><<<dim3(ceil_div(nc, 32), ceil_div(nc, 64), nl), 256, smem_k, st>>>(pp);
    else if (k == 3) prep_convs_kernel<T, 3><<<dim3(ceil_div(nc, 32), ceil_div(nc, 64), nl), 256, smem_k, st>>>(pp);
    else if (k == 1) prep_convs_kernel<T, 1><<<dim3(ceil_div(nc, 32), ceil_div(nc, 64), nl), 256, smem_k, st>>>(pp);
    else if (k == 7) prep_convs_kernel<T, 7><<<dim3(ceil_div(nc, 32), ceil_div(nc, 64), nl), 256, smem_k, st>>>(pp);
    else return RADTTS_ERR_UNSUPPORTED;
    RB_TRY(after_launch());
    PrepConvParams pr{};
    pr.N = nc; pr.C = nc;
    for (int i = 0; i < nl; ++i) pr.it[i] = pp.it[nl + i];
    prep_convs_kernel<T, 1><<<dim3(ceil_div(nc, 32), ceil_div(nc, 64), nl), 256, smem_1, st>>>(pr);
    RB_TRY(after_launch());
  }
  prep_end_kernel<T><<<grid_for((size_t)d.z_ld * nl * nc), 256, 0, st>>>(
      w.w_end, w.b_end, h, nc, nl, d.z_ld, reinterpret_cast<T*>(base + L.w_end), reinterpret_cast<float*>(base + L.b_end));
  RB_TRY(after_launch());
  if (!want_backward) return 0;

  auto transpose = [&](size_t src_off, int lds, int scol, int N, int ncols, size_t dst_off, int ldd, int dcol) -> int {
    dim3 g(ceil_div(N, 32), ceil_div(ncols, 32));
    prep_transpose_kernel<T><<<g, 256, 0, st>>>(reinterpret_cast<const T*>(base + src_off), lds, scol, N, ncols,
                                                reinterpret_cast<T*>(base + dst_off), ldd, dcol);
    return after_launch();
  };
  // end^T: [n_ch][end_kpad] (zero padded K) from the first K block of w_end
  fill_zero_kernel<T><<<grid_for((size_t)nc * L.end_kpad), 256, 0, st>>>(reinterpret_cast<T*>(base + L.w_end_t),
                                                                        (size_t)nc * L.end_kpad);
  RB_TRY(after_launch());
  RB_TRY(transpose(L.w_end, nl * nc, 0, d.z_ld, nc, L.w_end_t, L.end_kpad, 0));
  RB_TRY(transpose(L.w_start, ldstart, 0, nc, ldstart, L.w_start_t, nc, 0));
  return 0;
}

// ====================================================================================================
// forward / inverse
// ====================================================================================================
template <typename T>
static int wn_and_coupling(const radtts_flow_dims& d, const uint8_t* base, const FlowLayout& L, const PlanView& pv,
                           const radtts_flow_buffers& buf, int inverse, cudaStream_t st) {
  const int h = d.c_active / 2, nc = d.n_ch, k = d.ksize, nl = d.n_layers;
  const int rows = pv.rows_alloc;
  RowMeta meta{pv.pos(), pv.rem()};
  T* x = reinterpret_cast<T*>(buf.x);
  T* r = reinterpret_cast<T*>(buf.r);
  GemmDesc g{};
  g.rows_alloc = rows;
  g.plan = pv.hdr();
  // start
  g.nseg = 2;
  g.seg[0] = Seg{buf.ctx, L.ctx_ld, 0, 0, L.ctx_ld};
  g.seg[1] = Seg{buf.z0, 128, 0, 0, 128};
  g.w = base + L.w_start; g.ldw = L.ctx_ld + 128; g.N = nc;
  {
    EpiBiasAct<T, ACT_NONE> e{x, nc, 0, reinterpret_cast<const float*>(base + L.b_start), meta, ACT_NONE, 0, 0, k, 1};
    RB_TRY((run_gemm<T>(g, e, st)));
  }
  // res_skip(i) and in_layer(i + 1) both only READ x_{i+1} and write disjoint buffers: res_skip goes to a side stream
  // (fork / join with events -- also what a stream capture records), so its CTAs back-fill the SMs that the in_layer
  // GEMM leaves idle in its last wave (340 tiles on 148 SMs = 2.3 waves) instead of waiting for it to drain.
  SideStream* side = side_stream();
  if (!side) return RADTTS_ERR_UNSUPPORTED;
  for (int i = 0; i < nl; ++i) {
    T* xi = x + (size_t)i * rows * nc;
    T* xo = x + (size_t)(i + 1) * rows * nc;
    g.nseg = k;
    for (int t = 0; t < k; ++t) g.seg[t] = Seg{xi, nc, (t - k / 2) << i, 0, nc};
    g.w = base + L.w_in[i]; g.ldw = k * nc; g.N = nc;
    {
      EpiBiasAct<T, ACT_SOFTPLUS> e{xo, nc, 0, reinterpret_cast<const float*>(base + L.b_in[i]), meta, ACT_SOFTPLUS,
                                    d.partial_padding, i, k, 1};
      RB_TRY((run_gemm<T>(g, e, st)));
    }
    RB_CUDA(cudaEventRecord(side->fork[i], st));
    RB_CUDA(cudaStreamWaitEvent(side->stream, side->fork[i], 0));
    GemmDesc gr = g;
    gr.nseg = 1;
    gr.seg[0] = Seg{xo, nc, 0, 0, nc};
    gr.w = base + L.w_rs[i]; gr.ldw = nc; gr.N = nc;
    {
      EpiBiasAct<T, ACT_SOFTPLUS> e{r, nl * nc, i * nc, reinterpret_cast<const float*>(base + L.b_rs[i]), meta, ACT_SOFTPLUS,
                                    0, 0, k, 1};
      RB_TRY((run_gemm<T>(gr, e, side->stream)));
    }
  }
  RB_CUDA(cudaEventRecord(side->join, side->stream));
  RB_CUDA(cudaStreamWaitEvent(st, side->join, 0));
  g.nseg = 1;
  g.seg[0] = Seg{r, nl * nc, 0, 0, nl * nc};
  g.w = base + L.w_end; g.ldw = nl * nc; g.N = d.z_ld;
  {
    const float* b_end = reinterpret_cast<const float*>(base + L.b_end);
    const float* zsrc = inverse ? buf.zin : buf.zmid;
    float* zdst = inverse ? buf.zmid : buf.zout;
    float* ls = inverse ? nullptr : buf.log_s;
    float* pr = inverse ? nullptr : buf.params;
    constexpr bool kFast = sizeof(T) == 2;
    switch (d.scaling) {
      case 0: RB_TRY((run_gemm<T>(g, EpiCouplingT<kFast, 0>{b_end, zsrc, zdst, ls, pr, d.c_off, h, d.z_ld, inverse, meta}, st))); break;
      case 1: RB_TRY((run_gemm<T>(g, EpiCouplingT<kFast, 1>{b_end, zsrc, zdst, ls, pr, d.c_off, h, d.z_ld, inverse, meta}, st))); break;
      case 2: RB_TRY((run_gemm<T>(g, EpiCouplingT<kFast, 2>{b_end, zsrc, zdst, ls, pr, d.c_off, h, d.z_ld, inverse, meta}, st))); break;
      default: RB_TRY((run_gemm<T>(g, EpiCouplingT<kFast, 3>{b_end, zsrc, zdst, ls, pr, d.c_off, h, d.z_ld, inverse, meta}, st))); break;
    }
  }
  return 0;
}

template <typename T>
static int flowstep_impl(const radtts_flow_dims& d, const uint8_t* base, const PlanView& pv,
                         const radtts_flow_buffers& buf, int inverse, cudaStream_t st) {
  FlowLayout L = flow_layout(d, sizeof(T) == 4 ? RADTTS_PREC_FP32 : RADTTS_PREC_BF16, 0);
  const int h = d.c_active / 2;
  RowMeta meta{pv.pos(), pv.rem()};
  GemmDesc g{};
  g.rows_alloc = pv.rows_alloc;
  g.plan = pv.hdr();
  g.nseg = 1;
  g.w = base + L.w_inv; g.ldw = d.z_ld; g.N = d.z_ld;
  if (!inverse) {
    g.seg[0] = Seg{buf.zin, d.z_ld, 0, 0, d.z_ld};
    if (d.z_ld % 4 == 0 && d.z_ld <= kInvMaxLd) {
      RB_TRY(launch_invconv_rows<T>(buf.zin, reinterpret_cast<const float*>(base + L.w_inv), d.z_ld, pv.hdr(),
                                    pv.rows_alloc, meta, 1, buf.zmid, buf.zout, d.c_off + h, reinterpret_cast<T*>(buf.z0),
                                    d.c_off, h, st));
    } else {
      EpiInvConv<T> e{buf.zmid, buf.zout, reinterpret_cast<T*>(buf.z0), d.c_off, h, d.z_ld, meta};
      RB_TRY(launch_rowgemm_simt(g, e, st));
    }
    return wn_and_coupling<T>(d, base, L, pv, buf, 0, st);
  }
  extract_z0_kernel<T><<<grid_for((size_t)pv.rows_alloc * d.z_ld), 256, 0, st>>>(
      buf.zin, d.z_ld, d.c_off, h, pv.hdr(), buf.zmid, reinterpret_cast<T*>(buf.z0));
  RB_TRY(after_launch());
  RB_TRY(wn_and_coupling<T>(d, base, L, pv, buf, 1, st));
  if (d.z_ld % 4 == 0 && d.z_ld <= kInvMaxLd)
    return launch_invconv_rows<float>(buf.zmid, reinterpret_cast<const float*>(base + L.w_inv), d.z_ld, pv.hdr(),
                                      pv.rows_alloc, meta, 1, buf.zout, nullptr, 0, nullptr, 0, 0, st);
  g.seg[0] = Seg{buf.zmid, d.z_ld, 0, 0, d.z_ld};
  EpiStoreF32 e{buf.zout, d.z_ld, meta, 1};
  return launch_rowgemm_simt(g, e, st);
}

}  // namespace rb

using namespace rb;

extern "C" size_t radtts_frameplan_bytes(int B, int Tmax) {
  if (B <= 0 || Tmax <= 0) return 0;
  return plan_ints(B, Tmax) * sizeof(int);
}
extern "C" int radtts_frameplan_rows(int B, int Tmax) {
  if (B <= 0 || Tmax <= 0) return 0;
  return plan_rows_alloc(B, Tmax);
}
extern "C" int radtts_frameplan_build(const int64_t* lens, const int64_t* geom_lens, int divisor, int B, int Tmax,
                                      void* plan, void* stream) {
  if (!lens || !plan || B <= 0 || Tmax <= 0 || divisor <= 0) return RADTTS_ERR_INVALID_ARG;
  frameplan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(lens, geom_lens, divisor, B, Tmax, plan_rows_alloc(B, Tmax),
                                                        reinterpret_cast<int*>(plan));
  return after_launch();
}

extern "C" int radtts_pack_frames(const float* src, int B, int C, int T, int g, const void* plan, int Tmax, void* dst,
                                  int dst_bf16, int ld, int col_off, int ncols_pad, int valid_only, void* stream) {
  if (!src || !plan || !dst || B <= 0 || C <= 0 || T <= 0 || g <= 0 || ncols_pad < C * g) return RADTTS_ERR_INVALID_ARG;
  PlanView pv = make_plan_view(plan, B, Tmax);
  dim3 grid(pv.rows_alloc / 32, ceil_div(ncols_pad, 32));
  if (dst_bf16)
    pack_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(src, C, T, g, pv, reinterpret_cast<__nv_bfloat16*>(dst),
                                                                       ld, col_off, ncols_pad, valid_only);
  else
    pack_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(src, C, T, g, pv, reinterpret_cast<float*>(dst), ld,
                                                               col_off, ncols_pad, valid_only);
  return after_launch();
}
extern "C" int radtts_unpack_frames(const float* src, int ld, int col_off, const void* plan, int B, int Tmax, int C,
                                    int g, float* dst, void* stream) {
  if (!src || !plan || !dst || B <= 0 || C <= 0 || g <= 0) return RADTTS_ERR_INVALID_ARG;
  PlanView pv = make_plan_view(plan, B, Tmax);
  dim3 grid(ceil_div(Tmax, 32), ceil_div(C * g, 32), B);
  unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld, col_off, pv, C, g, dst);
  return after_launch();
}

extern "C" size_t radtts_flow_prepared_bytes(const radtts_flow_dims* dims, int precision, int want_backward) {
  if (check_dims(dims)) return 0;
  return flow_layout(*dims, precision, want_backward).total;
}
extern "C" int radtts_flow_prepare(const radtts_flow_dims* dims, const radtts_flow_weights* w, int inverse,
                                   int precision, int want_backward, void* prepared, size_t prepared_bytes,
                                   void* stream) {
  (void)inverse;  // the caller passes W or W^-1 in w_inv; nothing else depends on the direction
  RB_TRY(check_dims(dims));
  if (!w || !prepared) return RADTTS_ERR_INVALID_ARG;
  FlowLayout L = flow_layout(*dims, precision, want_backward);
  if (prepared_bytes < L.total) return RADTTS_ERR_WORKSPACE;
  if (precision == RADTTS_PREC_FP32)
    return prepare_impl<float>(*dims, *w, want_backward, reinterpret_cast<uint8_t*>(prepared), L, (cudaStream_t)stream);
  if (precision == RADTTS_PREC_BF16)
    return prepare_impl<__nv_bfloat16>(*dims, *w, want_backward, reinterpret_cast<uint8_t*>(prepared), L,
                                       (cudaStream_t)stream);
  return RADTTS_ERR_INVALID_ARG;
}

static int flowstep_entry(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B, int Tmax,
                          const radtts_flow_buffers* buf, int precision, int inverse, void* stream) {
  RB_TRY(check_dims(dims));
  if (!prepared || !plan || !buf || B <= 0 || Tmax <= 0) return RADTTS_ERR_INVALID_ARG;
  if (!buf->ctx || !buf->zin || !buf->zmid || !buf->zout || !buf->z0 || !buf->x || !buf->r) return RADTTS_ERR_INVALID_ARG;
  if (!inverse && !buf->log_s) return RADTTS_ERR_INVALID_ARG;
  PlanView pv = make_plan_view(plan, B, Tmax);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(prepared);
  if (precision == RADTTS_PREC_FP32) return flowstep_impl<float>(*dims, base, pv, *buf, inverse, (cudaStream_t)stream);
  if (precision == RADTTS_PREC_BF16)
    return flowstep_impl<__nv_bfloat16>(*dims, base, pv, *buf, inverse, (cudaStream_t)stream);
  return RADTTS_ERR_INVALID_ARG;
}
// One WN in_layer (dilated partial conv + softplus) in isolation: x[layer] -> x[layer + 1].  Profiling hook for the
// dominant kernel of the step (bench.py times it with CUDA events; ncu captures it by name).
template <typename T>
static int wn_layer_impl(const radtts_flow_dims& d, const uint8_t* base, const PlanView& pv, void* xbuf, int layer,
                         cudaStream_t st) {
  FlowLayout L = flow_layout(d, sizeof(T) == 4 ? RADTTS_PREC_FP32 : RADTTS_PREC_BF16, 0);
  const int nc = d.n_ch, k = d.ksize, rows = pv.rows_alloc;
  RowMeta meta{pv.pos(), pv.rem()};
  T* x = reinterpret_cast<T*>(xbuf);
  GemmDesc g{};
  g.rows_alloc = rows;
  g.plan = pv.hdr();
  g.nseg = k;
  for (int t = 0; t < k; ++t) g.seg[t] = Seg{x + (size_t)layer * rows * nc, nc, (t - k / 2) << layer, 0, nc};
  g.w = base + L.w_in[layer]; g.ldw = k * nc; g.N = nc;
  EpiBiasAct<T, ACT_SOFTPLUS> e{x + (size_t)(layer + 1) * rows * nc, nc, 0,
                                reinterpret_cast<const float*>(base + L.b_in[layer]), meta, ACT_SOFTPLUS, d.partial_padding,
                                layer, k, 1};
  return run_gemm<T>(g, e, st);
}
extern "C" int radtts_wn_layer_forward(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                       int Tmax, void* x, int layer, int precision, void* stream) {
  RB_TRY(check_dims(dims));
  if (!prepared || !plan || !x || layer < 0 || layer >= dims->n_layers) return RADTTS_ERR_INVALID_ARG;
  PlanView pv = make_plan_view(plan, B, Tmax);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(prepared);
  if (precision == RADTTS_PREC_FP32) return wn_layer_impl<float>(*dims, base, pv, x, layer, (cudaStream_t)stream);
  if (precision == RADTTS_PREC_BF16) return wn_layer_impl<__nv_bfloat16>(*dims, base, pv, x, layer, (cudaStream_t)stream);
  return RADTTS_ERR_INVALID_ARG;
}

extern "C" int radtts_flowstep_forward(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                       int Tmax, const radtts_flow_buffers* buf, int precision, void* stream) {
  return flowstep_entry(dims, prepared, plan, B, Tmax, buf, precision, 0, stream);
}
extern "C" int radtts_flowstep_inverse(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                       int Tmax, const radtts_flow_buffers* buf, int precision, void* stream) {
  return flowstep_entry(dims, prepared, plan, B, Tmax, buf, precision, 1, stream);
}
