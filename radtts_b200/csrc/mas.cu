// Kernel 1: batch-parallel Monotonic Alignment Search fused with backtrack and hard-attention scatter.
//
// Replaces RADTTS.binarize_attention (reference radtts.py:320-334) + alignment.mas_width1 (reference
// alignment.py:31-59).  One CTA per utterance:
//   warp 1   : TMA producer -- streams the utterance's (out_len x T2) cost rows HBM -> smem ring with 1-D
//              bulk async copies (cp.async.bulk + mbarrier), many KB in flight per SM;
//   warp 0   : dynamic-programming warp -- each lane owns NC consecutive text columns in registers; the
//              j-1 neighbour of a lane's first column comes from one __shfl_up per row (off the critical
//              path); the `diagonal` decision of every cell is kept as ONE BIT (warp ballot) in smem;
//   warps 2-7: zero-fill the utterance's slab of the dense hard map while the DP runs;
//   then lane 0 backtracks through the bit lattice and all threads scatter the ones, frame->token
//   indices and durations.
// Arithmetic is exactly the reference's: one fp32 add per cell, `>=` tie-break toward the diagonal,
// row 0 restricted to column 0, and the unconditional opt[0,0] = 1 (alignment.py:59).
#include <math_constants.h>

#include "common.cuh"
#include "ptx.cuh"

namespace rb {

constexpr int kMasThreads = 256;
constexpr int kMasMaxStages = 16;
constexpr int kMasHeaderBytes = 2 * kMasMaxStages * 8;

struct MasPlan {
  int nc;               // columns per lane (odd -> conflict-free smem reads)
  int rows_per_chunk;   // rows per bulk copy
  int stages;           // ring depth
  uint32_t stage_bytes;
  uint32_t path_off, dur_off, bits_off, ring_off, smem_bytes;
  int bits_in_smem;
  size_t bits_ws_bytes;  // global fallback for the bit lattice
};

static int make_plan(int B, int T1, int T2, MasPlan* p) {
  int nc = (T2 + 31) / 32;
  if (nc % 2 == 0) nc += 1;
  if (nc > 17 || T1 > 65535) return RADTTS_ERR_UNSUPPORTED;
  p->nc = nc;
  int r = 12288 / (T2 * 4);
  if (r < 1) r = 1;
  if (r > 64) r = 64;
  p->rows_per_chunk = r;
  p->stage_bytes = (uint32_t)round_up(r * T2 * 4 + 32, 128);
  p->path_off = kMasHeaderBytes;
  p->dur_off = p->path_off + (uint32_t)round_up(T1 * 2, 16);
  p->bits_off = p->dur_off + (uint32_t)round_up(T2 * 4, 16);
  uint32_t bits_bytes = (uint32_t)round_up(T1 * nc * 4, 128);
  uint32_t fixed = (uint32_t)round_up((int)p->bits_off, 128);
  p->bits_off = fixed;
  if (fixed + bits_bytes + 3 * p->stage_bytes <= (uint32_t)kSmemBudget) {
    p->bits_in_smem = 1;
    p->ring_off = fixed + bits_bytes;
    p->bits_ws_bytes = 0;
  } else {
    p->bits_in_smem = 0;
    p->ring_off = fixed;
    p->bits_ws_bytes = (size_t)B * T1 * nc * 4;
  }
  int st = (int)((kSmemBudget - p->ring_off) / p->stage_bytes);
  if (st > kMasMaxStages) st = kMasMaxStages;
  if (st < 2) return RADTTS_ERR_UNSUPPORTED;
  p->stages = st;
  p->smem_bytes = p->ring_off + (uint32_t)st * p->stage_bytes;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Optional pre-pass for probability input: y = logf(x).  (HBM-bound, all SMs.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mas_log_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t n4 = n / 4;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float4* y4 = reinterpret_cast<float4*>(y);
  for (size_t k = i; k < n4; k += stride) {
    float4 v = x4[k];
    v.x = logf(v.x); v.y = logf(v.y); v.z = logf(v.z); v.w = logf(v.w);
    y4[k] = v;
  }
  for (size_t k = n4 * 4 + i; k < n; k += stride) y[k] = logf(x[k]);
}

__device__ __forceinline__ void zero_fill(float* p, size_t n, int tid, int nthreads) {
  // head to 16-byte alignment, float4 body, scalar tail
  size_t head = ((16 - ((uintptr_t)p & 15)) & 15) / 4;
  if (head > n) head = n;
  for (size_t k = tid; k < head; k += nthreads) p[k] = 0.f;
  float4* p4 = reinterpret_cast<float4*>(p + head);
  size_t n4 = (n - head) / 4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t k = tid; k < n4; k += nthreads) p4[k] = z;
  for (size_t k = head + n4 * 4 + tid; k < n; k += nthreads) p[k] = 0.f;
}

template <int NC>
__global__ void __launch_bounds__(kMasThreads, 1)
mas_kernel(const float* __restrict__ logp, const int64_t* __restrict__ in_lens, const int64_t* __restrict__ out_lens,
           int T1, int T2, unsigned long long total_bytes, float* __restrict__ hard, int32_t* __restrict__ f2t,
           int32_t* __restrict__ dur, uint32_t* __restrict__ bits_ws, MasPlan plan) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMasMaxStages;
  uint16_t* path = reinterpret_cast<uint16_t*>(smem + plan.path_off);
  int* dur_s = reinterpret_cast<int*>(smem + plan.dur_off);
  uint8_t* ring = smem + plan.ring_off;

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  long long ol = out_lens[b], il = in_lens[b];
  const int olen = (int)(ol < 0 ? 0 : (ol > T1 ? T1 : ol));
  const int ilen = (int)(il < 0 ? 0 : (il > T2 ? T2 : il));
  uint32_t* bits = plan.bits_in_smem ? reinterpret_cast<uint32_t*>(smem + plan.bits_off)
                                     : bits_ws + (size_t)b * T1 * NC;
  const int R = plan.rows_per_chunk;
  const int stages = plan.stages;
  const bool active = (olen > 0 && ilen > 0);
  const int nchunks = active ? (olen + R - 1) / R : 0;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_fence_init();
  }
  for (int j = tid; j < T2; j += kMasThreads) dur_s[j] = 0;
  __syncthreads();

  const size_t slab = (size_t)b * T1 * T2;

  if (warp == 1) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      const unsigned long long total16 = total_bytes & ~15ull;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % stages;
        const int k = c / stages;
        if (k > 0) mbar_wait(&empty[s], (uint32_t)((k & 1) ^ 1));
        const int r0 = c * R;
        const int r1 = min(r0 + R, olen);
        const unsigned long long sb = (unsigned long long)(slab + (size_t)r0 * T2) * 4ull;
        const unsigned long long eb = (unsigned long long)(slab + (size_t)r1 * T2) * 4ull;
        const unsigned long long s16 = sb & ~15ull;
        unsigned long long e16 = (eb + 15ull) & ~15ull;
        if (e16 > total16) e16 = total16;
        uint8_t* dst = ring + (size_t)s * plan.stage_bytes;
        const uint32_t nbytes = e16 > s16 ? (uint32_t)(e16 - s16) : 0u;
        // bytes past the last 16-byte boundary of the tensor (at most 12) are fetched by hand
        const unsigned long long t0 = (s16 + nbytes > sb) ? (s16 + nbytes) : sb;
        for (unsigned long long x = t0; x < eb; x += 4)
          *reinterpret_cast<float*>(dst + (x - s16)) = *reinterpret_cast<const float*>(
              reinterpret_cast<const uint8_t*>(logp) + x);
        mbar_arrive_expect_tx(&full[s], nbytes);
        if (nbytes) bulk_g2s(dst, reinterpret_cast<const uint8_t*>(logp) + s16, nbytes, &full[s]);
      }
    }
  } else if (warp == 0) {
    // ------------------------------ DP warp ------------------------------
    if (active) {
      float v[NC];
      int off[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) off[c] = min(lane * NC + c, T2 - 1);
      const float nanv = __int_as_float(0x7fc00000);
      int row = 0;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % stages;
        mbar_wait(&full[s], (uint32_t)((c / stages) & 1));
        const int r0 = c * R;
        const int r1 = min(r0 + R, olen);
        const unsigned long long sb = (unsigned long long)(slab + (size_t)r0 * T2) * 4ull;
        const float* st = reinterpret_cast<const float*>(ring + (size_t)s * plan.stage_bytes + (sb & 15ull));
        float a[NC];
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) a[cc] = st[off[cc]];
        for (; row < r1; ++row) {
          float an[NC];
          const float* nx = st + (size_t)(row + 1 - r0) * T2;
          if (row + 1 < r1) {
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) an[cc] = nx[off[cc]];
          } else {
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) an[cc] = 0.f;
          }
          if (row == 0) {
#pragma unroll
            for (int cc = 0; cc < NC; ++cc) v[cc] = (lane * NC + cc == 0) ? a[cc] : -CUDART_INF_F;
          } else {
            float left = __shfl_up_sync(0xffffffffu, v[NC - 1], 1);
            if (lane == 0) left = nanv;  // column 0 has no diagonal predecessor: NaN >= x is false
            uint32_t mine = 0;
#pragma unroll
            for (int cc = NC - 1; cc >= 0; --cc) {
              const float l = (cc == 0) ? left : v[cc - 1];
              const float u = v[cc];
              const bool diag = (l >= u);
              const float m = diag ? l : u;
              v[cc] = __fadd_rn(a[cc], m);
              const uint32_t w = __ballot_sync(0xffffffffu, diag);
              if (lane == cc) mine = w;
            }
            if (lane < NC) bits[(size_t)row * NC + lane] = mine;
          }
#pragma unroll
          for (int cc = 0; cc < NC; ++cc) a[cc] = an[cc];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
      }
      __syncwarp();
      // ------------------------------ backtrack ------------------------------
      if (lane == 0) {
        int j = ilen - 1;
        for (int i = olen - 1; i >= 1; --i) {
          path[i] = (uint16_t)j;
          const uint32_t w = bits[(size_t)i * NC + (j % NC)];
          j -= (int)((w >> (j / NC)) & 1u);
        }
        path[0] = (uint16_t)j;
      }
    }
  } else {
    // ------------------------------ zero fill ------------------------------
    zero_fill(hard + slab, (size_t)T1 * T2, tid - 64, kMasThreads - 64);
  }
  __syncthreads();

  // ------------------------------ scatter ------------------------------
  if (active) {
    for (int i = tid; i < olen; i += kMasThreads) {
      const int j = path[i];
      hard[slab + (size_t)i * T2 + j] = 1.0f;
      if (f2t) f2t[(size_t)b * T1 + i] = j;
      atomicAdd(&dur_s[j], 1);
    }
    if (tid == 0) {
      hard[slab] = 1.0f;  // alignment.py:59
    }
  }
  if (f2t)
    for (int i = olen + tid; i < T1; i += kMasThreads) f2t[(size_t)b * T1 + i] = -1;
  __syncthreads();
  if (dur) {
    if (active && tid == 0 && path[0] != 0) dur_s[0] += 1;
    __syncthreads();
    for (int j = tid; j < T2; j += kMasThreads) dur[(size_t)b * T2 + j] = dur_s[j];
  }
}

template <int NC>
static int launch_mas(const float* logp, const int64_t* in_lens, const int64_t* out_lens, int B, int T1, int T2,
                      float* hard, int32_t* f2t, int32_t* dur, uint32_t* bits_ws, const MasPlan& plan,
                      cudaStream_t stream) {
  static int configured_smem = 0;
  if ((int)plan.smem_bytes > configured_smem) {
    RB_CUDA(cudaFuncSetAttribute(mas_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured_smem = kSmemBudget;
  }
  const unsigned long long total_bytes = (unsigned long long)B * T1 * T2 * 4ull;
  mas_kernel<NC><<<B, kMasThreads, plan.smem_bytes, stream>>>(logp, in_lens, out_lens, T1, T2, total_bytes, hard,
                                                             f2t, dur, bits_ws, plan);
  return after_launch();
}

}  // namespace rb

using namespace rb;

extern "C" size_t radtts_mas_workspace_bytes(int B, int T1, int T2, int is_prob) {
  if (B <= 0 || T1 <= 0 || T2 <= 0) return 0;
  MasPlan p;
  if (make_plan(B, T1, T2, &p)) return 0;
  size_t n = 0;
  if (is_prob) n += round_up((size_t)B * T1 * T2 * 4, (size_t)256);
  n += round_up(p.bits_ws_bytes, (size_t)256);
  return n;
}

extern "C" int radtts_mas_forward(const float* attn, int is_prob, const int64_t* in_lens, const int64_t* out_lens,
                                  int B, int T1, int T2, float* attn_hard, int32_t* frame_to_token,
                                  int32_t* durations, void* ws, size_t ws_bytes, void* stream_) {
  if (B < 0 || T1 < 0 || T2 < 0) return RADTTS_ERR_INVALID_ARG;
  if (B == 0 || T1 == 0 || T2 == 0) return 0;
  if (!attn || !in_lens || !out_lens || !attn_hard) return RADTTS_ERR_INVALID_ARG;
  if (((uintptr_t)attn & 15) != 0) return RADTTS_ERR_INVALID_ARG;
  cudaStream_t stream = (cudaStream_t)stream_;
  MasPlan plan;
  RB_TRY(make_plan(B, T1, T2, &plan));
  if (ws_bytes < radtts_mas_workspace_bytes(B, T1, T2, is_prob)) return RADTTS_ERR_WORKSPACE;
  if (ws_bytes > 0 && (!ws || ((uintptr_t)ws & 15) != 0)) return RADTTS_ERR_INVALID_ARG;
  uint8_t* wsp = reinterpret_cast<uint8_t*>(ws);
  const float* logp = attn;
  if (is_prob) {
    const size_t n = (size_t)B * T1 * T2;
    float* lbuf = reinterpret_cast<float*>(wsp);
    wsp += round_up(n * 4, (size_t)256);
    int blocks = (int)((n / 4 + 255) / 256);
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (blocks < 1) blocks = 1;
    mas_log_kernel<<<blocks, 256, 0, stream>>>(attn, lbuf, n);
    RB_TRY(after_launch());
    logp = lbuf;
  }
  uint32_t* bits_ws = plan.bits_in_smem ? nullptr : reinterpret_cast<uint32_t*>(wsp);
  switch (plan.nc) {
#define RB_MAS_CASE(N) \
  case N: return launch_mas<N>(logp, in_lens, out_lens, B, T1, T2, attn_hard, frame_to_token, durations, bits_ws, plan, stream);
    RB_MAS_CASE(1) RB_MAS_CASE(3) RB_MAS_CASE(5) RB_MAS_CASE(7) RB_MAS_CASE(9) RB_MAS_CASE(11) RB_MAS_CASE(13)
    RB_MAS_CASE(15) RB_MAS_CASE(17)
#undef RB_MAS_CASE
    default: return RADTTS_ERR_UNSUPPORTED;
  }
}
