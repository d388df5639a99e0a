// LU-parameterised 1x1-conv weights of a whole flow stack in a handful of launches
// (Invertible1x1ConvLUS, reference common.py:407-428):
//     U = triu(upper, 1) + diag(upper_diag),  L = tril(lower, -1) + diag(lower_diag),  W = P (L U),
//     log_det_W = sum log|upper_diag|
// and what autograd derives for it, given G = dLoss/dW and g = dLoss/dlog_det_W:
//     A = P^T G,   g_lower = tril(A U^T, -1),   M = L^T A,   g_upper = triu(M, 1),
//     g_upper_diag = diag(M) + g / upper_diag.
// The reference runs ~12 tiny kernels per flow forward (triu, tril, diag, add, two mm, abs, log, sum ...) and ~30
// backward; here every product of ALL flows is one batched launch of a small masked fp32 matmul: 2 launches forward,
// 2 backward.  C <= 256 (160 ... 154 in the shipped configs); 2 C^3 = 8 MFLOP per product -- latency, not throughput.
#include "common.cuh"

namespace rb {

constexpr int kLusMaxBatch = 32;

enum LusMask : int { kMaskNone = 0, kMaskUnitLower = 1, kMaskUpperDiag = 2 };
enum LusOut : int { kOutFull = 0, kOutStrictLower = 1, kOutStrictUpperPlusDiag = 2 };

struct LusMm {
  const float* a; const float* b; float* c;
  const float* diag_a; const float* diag_b;   // diagonal source for a masked operand
  float* out_diag; const float* g_scalar; const float* ud;   // kOutStrictUpperPlusDiag: out_diag[i] = C[i][i] + *g / ud[i]
  int n, lda, ldb, ldc;
  int trans_a, trans_b, mask_a, mask_b, out_mode;
};
struct LusBatch { LusMm p[kLusMaxBatch]; };

__device__ __forceinline__ float lus_elem(const float* m, int ld, int i, int j, int mask, const float* diag) {
  // (i, j) = coordinates in the STORED matrix
  if (mask == kMaskUnitLower) return i > j ? m[(size_t)i * ld + j] : (i == j ? diag[i] : 0.f);
  if (mask == kMaskUpperDiag) return i < j ? m[(size_t)i * ld + j] : (i == j ? diag[i] : 0.f);
  return m[(size_t)i * ld + j];
}

// C = op(A) op(B), 32x32 output tile per CTA, 256 threads x (2x2), K in steps of 32 through smem
__global__ void __launch_bounds__(256) lus_mm_kernel(const __grid_constant__ LusBatch batch) {
  const LusMm& q = batch.p[blockIdx.z];
  const int n = q.n;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  if (i0 >= n || j0 >= n) return;
  __shared__ float As[32][33], Bs[32][33];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < n; k0 += 32) {
    for (int e = threadIdx.x; e < 1024; e += 256) {
      const int r = e >> 5, c = e & 31;
      {  // op(A)[i0 + r][k0 + c]
        const int i = i0 + r, k = k0 + c;
        float v = 0.f;
        if (i < n && k < n) v = q.trans_a ? lus_elem(q.a, q.lda, k, i, q.mask_a, q.diag_a) : lus_elem(q.a, q.lda, i, k, q.mask_a, q.diag_a);
        As[r][c] = v;
      }
      {  // op(B)[k0 + r][j0 + c]
        const int k = k0 + r, j = j0 + c;
        float v = 0.f;
        if (k < n && j < n) v = q.trans_b ? lus_elem(q.b, q.ldb, j, k, q.mask_b, q.diag_b) : lus_elem(q.b, q.ldb, k, j, q.mask_b, q.diag_b);
        Bs[r][c] = v;
      }
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[ty][k], a1 = As[ty + 16][k], b0 = Bs[k][tx], b1 = Bs[k][tx + 16];
      acc[0][0] = fmaf(a0, b0, acc[0][0]); acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]); acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int i = i0 + ty + 16 * a, j = j0 + tx + 16 * b;
      if (i >= n || j >= n) continue;
      float v = acc[a][b];
      if (q.out_mode == kOutStrictLower) v = i > j ? v : 0.f;
      if (q.out_mode == kOutStrictUpperPlusDiag) {
        if (i == j) q.out_diag[i] = v + (q.g_scalar ? *q.g_scalar : 0.f) / q.ud[i];
        v = i < j ? v : 0.f;
      }
      q.c[(size_t)i * q.ldc + j] = v;
    }
}

struct LusLogdet { const float* ud[kLusMaxBatch]; float* out[kLusMaxBatch]; int n[kLusMaxBatch]; };

__global__ void __launch_bounds__(256) lus_logdet_kernel(const __grid_constant__ LusLogdet p) {
  const float* ud = p.ud[blockIdx.x];
  const int n = p.n[blockIdx.x];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += logf(fabsf(ud[i]));
  __shared__ float red[8];
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    *p.out[blockIdx.x] = t;
  }
}

static int launch_mm(const LusBatch& b, int count, int nmax, cudaStream_t s) {
  dim3 grid(ceil_div(nmax, 32), ceil_div(nmax, 32), count);
  lus_mm_kernel<<<grid, 256, 0, s>>>(b);
  return after_launch();
}

}  // namespace rb

using namespace rb;

extern "C" int radtts_lus_compose(int n_flows, const int* n_host, const float* const* lower_host,
                                  const float* const* upper_host, const float* const* upper_diag_host,
                                  const float* const* lower_diag_host, const float* const* p_host,
                                  float* const* tmp_host, float* const* w_host, float* const* log_det_host, void* stream) {
  if (n_flows <= 0 || n_flows > kLusMaxBatch || !n_host || !lower_host || !upper_host || !upper_diag_host ||
      !lower_diag_host || !p_host || !tmp_host || !w_host || !log_det_host)
    return RADTTS_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  LusBatch b1{}, b2{};
  LusLogdet ld{};
  int nmax = 0;
  for (int k = 0; k < n_flows; ++k) {
    const int n = n_host[k];
    if (n <= 0 || n > 256) return RADTTS_ERR_UNSUPPORTED;
    nmax = n > nmax ? n : nmax;
    LusMm& a = b1.p[k];   // tmp = L U
    a.a = lower_host[k]; a.b = upper_host[k]; a.c = tmp_host[k]; a.diag_a = lower_diag_host[k]; a.diag_b = upper_diag_host[k];
    a.n = n; a.lda = a.ldb = a.ldc = n; a.mask_a = kMaskUnitLower; a.mask_b = kMaskUpperDiag; a.out_mode = kOutFull;
    LusMm& c = b2.p[k];   // W = P tmp
    c.a = p_host[k]; c.b = tmp_host[k]; c.c = w_host[k]; c.n = n; c.lda = c.ldb = c.ldc = n; c.out_mode = kOutFull;
    ld.ud[k] = upper_diag_host[k]; ld.out[k] = log_det_host[k]; ld.n[k] = n;
  }
  RB_TRY(launch_mm(b1, n_flows, nmax, s));
  RB_TRY(launch_mm(b2, n_flows, nmax, s));
  lus_logdet_kernel<<<n_flows, 256, 0, s>>>(ld);
  return after_launch();
}

extern "C" int radtts_lus_backward(int n_flows, const int* n_host, const float* const* lower_host,
                                   const float* const* upper_host, const float* const* upper_diag_host,
                                   const float* const* lower_diag_host, const float* const* p_host,
                                   const float* const* g_w_host, const int* g_w_ld_host,
                                   const float* const* g_log_det_host, float* const* tmp_host, float* const* g_lower_host,
                                   float* const* g_upper_host, float* const* g_upper_diag_host, void* stream) {
  if (n_flows <= 0 || 2 * n_flows > kLusMaxBatch || !n_host || !lower_host || !upper_host || !upper_diag_host ||
      !lower_diag_host || !p_host || !g_w_host || !g_w_ld_host || !g_log_det_host || !tmp_host || !g_lower_host ||
      !g_upper_host || !g_upper_diag_host)
    return RADTTS_ERR_INVALID_ARG;
  cudaStream_t s = (cudaStream_t)stream;
  LusBatch b1{}, b2{};
  int nmax = 0;
  for (int k = 0; k < n_flows; ++k) {
    const int n = n_host[k];
    if (n <= 0 || n > 256) return RADTTS_ERR_UNSUPPORTED;
    nmax = n > nmax ? n : nmax;
    LusMm& a = b1.p[k];   // A = P^T G
    a.a = p_host[k]; a.trans_a = 1; a.b = g_w_host[k]; a.ldb = g_w_ld_host[k]; a.c = tmp_host[k];
    a.n = n; a.lda = a.ldc = n; a.out_mode = kOutFull;
    LusMm& l = b2.p[2 * k];       // g_lower = tril(A U^T, -1)
    l.a = tmp_host[k]; l.b = upper_host[k]; l.trans_b = 1; l.mask_b = kMaskUpperDiag; l.diag_b = upper_diag_host[k];
    l.c = g_lower_host[k]; l.n = n; l.lda = l.ldb = l.ldc = n; l.out_mode = kOutStrictLower;
    LusMm& u = b2.p[2 * k + 1];   // M = L^T A; g_upper = triu(M, 1); g_upper_diag = diag(M) + g / upper_diag
    u.a = lower_host[k]; u.trans_a = 1; u.mask_a = kMaskUnitLower; u.diag_a = lower_diag_host[k]; u.b = tmp_host[k];
    u.c = g_upper_host[k]; u.n = n; u.lda = u.ldb = u.ldc = n; u.out_mode = kOutStrictUpperPlusDiag;
    u.out_diag = g_upper_diag_host[k]; u.g_scalar = g_log_det_host[k]; u.ud = upper_diag_host[k];
  }
  RB_TRY(launch_mm(b1, n_flows, nmax, s));
  return launch_mm(b2, 2 * n_flows, nmax, s);
}
