// Kernel 3: fused ConvAttention core -- pairwise L2 distance, log-softmax over text, + log prior, masked softmax.
// Replaces reference common.py:907-923: the (B, 80, T1, T2) broadcast temporary (1.2 GB at 32x800x150) is never
// materialised; HBM traffic is the algorithmic 12 B per (b, t1, t2) cell (read prior, write attn_logprob, write attn).
//
// Forward: one CTA = 32 mel frames of one utterance; the utterance's projected keys live in smem channel-major
// (ks[c][t2], conflict-free for lanes that own consecutive t2), a warp owns a frame: lanes hold the frame's T2
// distances in registers, two warp reductions give the log-softmax and the masked softmax.
// Backward: pass 1 recomputes the distances, forms g_d per frame (Appendix C of SURVEY.md), reduces g_q in-warp
// and spills g_d (B, T1, T2) once; pass 2 walks text-major over g_d to reduce g_k.
#include <math_constants.h>

#include "common.cuh"

namespace rb {

constexpr int kAttRows = 32;     // frames per CTA
constexpr int kAttThreads = 256; // 8 warps x 4 frames
constexpr int kAttMaxJ = 18;     // T2 <= 576

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// stage keys (C x T2, T2 padded to T2p) and the CTA's query tile (C x 32) into smem
__device__ __forceinline__ void stage_qk(const float* __restrict__ q, const float* __restrict__ k, int C, int T1, int T2,
                                         int T2p, int t1_0, float* ks, float* qs) {
  for (int i = threadIdx.x; i < C * T2p; i += kAttThreads) {
    const int c = i / T2p, t2 = i - c * T2p;
    ks[i] = t2 < T2 ? k[(size_t)c * T2 + t2] : 0.f;
  }
  for (int i = threadIdx.x; i < C * kAttRows; i += kAttThreads) {
    const int c = i / kAttRows, r = i - c * kAttRows;
    qs[i] = (t1_0 + r) < T1 ? q[(size_t)c * T1 + t1_0 + r] : 0.f;
  }
}

// Forward.  A CTA stages the utterance's projected keys ONCE (ks[c][t2] channel-major + a table of |k|^2) and then walks
// over 32-frame tiles of that utterance (grid.x CTAs per utterance share its tiles round-robin), so the 100 KB key stage is
// amortised over many tiles.  Inside a tile a warp owns FOUR consecutive frames: every key value fetched from smem feeds
// four FMAs (the one-frame version was bound by shared-memory loads: one LDS per FMA), and the squared distance is
// expanded, |q - k|^2 = |q|^2 + |k|^2 - 2 q.k -- one FMA per (frame, key, channel) instead of a subtract and an FMA.
// The prior row of each frame is fetched into registers BEFORE the channel loop, which hides its DRAM latency.
// Frames per warp: with four, a channel step is 11 shared-memory wavefronts for 44 FMA instructions per warp; eight (12
// wavefronts for 88 FMAs, prior rows fetched in the epilogue because the accumulators fill the register budget) was
// measured SLOWER at 64 x 2000 x 300 (0.62 vs 0.43 ms): one 8-warp CTA per SM cannot hide the epilogue's latencies.  What
// the four-frame kernel spends outside the FMAs is the epilogue -- two exponentials and a logarithm per cell -- which
// uses the MUFU forms (ex2 / lg2.approx, relative error 2^-21: far inside the 1e-3 parity bound of this kernel).
constexpr int kFwdThreads = 256;
template <int NJ>
struct FwdCfg {
  static constexpr int kFr = 4;                                // frames per warp (8 measured slower: see below)
  static constexpr int kRows = (kFwdThreads / 32) * kFr;       // frames per tile
  static constexpr bool kPrefetchPrior = true;
};

template <int NJ>
__global__ void __launch_bounds__(kFwdThreads) convattn_fwd_kernel(const float* __restrict__ q_enc, const float* __restrict__ k_enc,
                                                                   const float* __restrict__ prior, const int64_t* __restrict__ key_lens,
                                                                   int C, int T1, int T2, float temp, float* __restrict__ attn,
                                                                   float* __restrict__ logprob, float* __restrict__ lse_out) {
  extern __shared__ __align__(16) float sm[];
  constexpr int kFwdFr = FwdCfg<NJ>::kFr, kFwdRows = FwdCfg<NJ>::kRows;
  constexpr bool kPrefetch = FwdCfg<NJ>::kPrefetchPrior;
  const int T2p = NJ * 32;
  float* ks = sm;                                   // [C][T2p]
  float* qs = sm + (size_t)C * T2p;                 // [C][kFwdRows]
  float* kn = qs + (size_t)C * kFwdRows;            // [T2p]
  const int b = blockIdx.y;
  const float* q = q_enc + (size_t)b * C * T1;
  {
    const float* k = k_enc + (size_t)b * C * T2;
    for (int i = threadIdx.x; i < C * T2p; i += kFwdThreads) {
      const int c = i / T2p, t2 = i - c * T2p;
      ks[i] = t2 < T2 ? k[(size_t)c * T2 + t2] : 0.f;
    }
  }
  __syncthreads();
  for (int t2 = threadIdx.x; t2 < T2p; t2 += kFwdThreads) {
    float s = 0.f;
    for (int c = 0; c < C; ++c) { const float v = ks[(size_t)c * T2p + t2]; s = fmaf(v, v, s); }
    kn[t2] = s;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int klen = key_lens ? (int)min((long long)key_lens[b], (long long)T2) : T2;
  const int r0 = warp * kFwdFr;
  const int n_tiles = (T1 + kFwdRows - 1) / kFwdRows;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int t1_0 = tile * kFwdRows;
    __syncthreads();                                // previous tile's readers of qs are done (and kn is visible)
    for (int i = threadIdx.x; i < C * kFwdRows; i += kFwdThreads) {
      const int c = i / kFwdRows, r = i - c * kFwdRows;
      qs[i] = (t1_0 + r) < T1 ? q[(size_t)c * T1 + t1_0 + r] : 0.f;
    }
    __syncthreads();
    if (t1_0 + r0 >= T1) continue;
    // prior rows of this warp's frames, in flight underneath the channel loop
    float pr[kPrefetch ? kFwdFr : 1][NJ];
    if (prior && kPrefetch) {
#pragma unroll
      for (int f = 0; f < kFwdFr; ++f) {
        const int t1 = t1_0 + r0 + f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int t2 = lane + 32 * j;
          pr[f][j] = (t1 < T1 && t2 < T2) ? __ldg(prior + ((size_t)b * T1 + t1) * T2 + t2) : 0.f;
        }
      }
    }
    float dot[kFwdFr][NJ];
    float qn[kFwdFr];
#pragma unroll
    for (int f = 0; f < kFwdFr; ++f) {
      qn[f] = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) dot[f][j] = 0.f;
    }
    for (int c = 0; c < C; ++c) {
      float qv[kFwdFr];
#pragma unroll
      for (int f = 0; f < kFwdFr; f += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(qs + (size_t)c * kFwdRows + r0 + f);
        qv[f] = q4.x; qv[f + 1] = q4.y; qv[f + 2] = q4.z; qv[f + 3] = q4.w;
      }
      const float* kr = ks + (size_t)c * T2p + lane;
#pragma unroll
      for (int f = 0; f < kFwdFr; ++f) qn[f] = fmaf(qv[f], qv[f], qn[f]);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const float kv = kr[32 * j];
#pragma unroll
        for (int f = 0; f < kFwdFr; ++f) dot[f][j] = fmaf(qv[f], kv, dot[f][j]);
      }
    }
#pragma unroll
    for (int f = 0; f < kFwdFr; ++f) {
      const int t1 = t1_0 + r0 + f;
      if (t1 >= T1) break;
      const size_t row_off = ((size_t)b * T1 + t1) * T2;
      float d[NJ];
      float m = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        d[j] = -temp * fmaxf(fmaf(-2.f, dot[f][j], qn[f] + kn[lane + 32 * j]), 0.f);
        if (lane + 32 * j < T2) m = fmaxf(m, d[j]);
      }
      if (prior) {
        if (!kPrefetch) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int t2 = lane + 32 * j;
            pr[0][j] = t2 < T2 ? __ldg(prior + row_off + t2) : 0.f;
          }
        }
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (lane + 32 * j < T2) s += __expf(d[j] - m);
        s = warp_sum(s);
        const float lse = m + __logf(s);
        if (lane == 0 && lse_out) lse_out[(size_t)b * T1 + t1] = lse;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          d[j] = (lane + 32 * j < T2) ? (d[j] - lse) + __logf(pr[kPrefetch ? f : 0][j] + 1e-8f) : -CUDART_INF_F;
      } else {
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (lane + 32 * j >= T2) d[j] = -CUDART_INF_F;
      }
      float m2 = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int t2 = lane + 32 * j;
        if (t2 < T2) logprob[row_off + t2] = d[j];
        if (t2 < klen) m2 = fmaxf(m2, d[j]);
      }
      m2 = warp_max(m2);
      float s2 = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        d[j] = (lane + 32 * j < klen) ? __expf(d[j] - m2) : 0.f;
        s2 += d[j];
      }
      s2 = warp_sum(s2);
      const float inv = 1.f / s2;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int t2 = lane + 32 * j;
        if (t2 < T2) attn[row_off + t2] = d[j] * inv;
      }
    }
  }
}

// backward pass 1: per frame g_d (spilled to gd) and g_q
template <int NJ>
__global__ void __launch_bounds__(kAttThreads) convattn_bwd_rows_kernel(
    const float* __restrict__ q_enc, const float* __restrict__ k_enc, const float* __restrict__ lse_in,
    const float* __restrict__ attn, const float* __restrict__ g_attn, const float* __restrict__ g_logprob, int has_prior,
    int C, int T1, int T2, float temp, float* __restrict__ gd, float* __restrict__ g_q) {
  extern __shared__ float sm[];
  const int T2p = NJ * 32;
  float* ks = sm;
  float* qs = sm + (size_t)C * T2p;
  const int b = blockIdx.y, t1_0 = blockIdx.x * kAttRows;
  stage_qk(q_enc + (size_t)b * C * T1, k_enc + (size_t)b * C * T2, C, T1, T2, T2p, t1_0, ks, qs);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rr = warp; rr < kAttRows; rr += kAttThreads / 32) {
    const int t1 = t1_0 + rr;
    if (t1 >= T1) break;
    const size_t row_off = ((size_t)b * T1 + t1) * T2;
    float ga[NJ];
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int t2 = lane + 32 * j;
      ga[j] = (t2 < T2 && g_attn) ? attn[row_off + t2] * g_attn[row_off + t2] : 0.f;  // y * g_soft
      dot += ga[j];
    }
    dot = warp_sum(dot);
    float sum_ga = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int t2 = lane + 32 * j;
      float v = 0.f;
      if (t2 < T2) {
        const float y = attn[row_off + t2];
        v = ga[j] - y * dot + (g_logprob ? g_logprob[row_off + t2] : 0.f);
      }
      ga[j] = v;
      sum_ga += v;
    }
    float S = 0.f;
    if (has_prior) {
      sum_ga = warp_sum(sum_ga);
      // recompute d to get softmax(d) = exp(d - lse)
      float d[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) d[j] = 0.f;
      for (int c = 0; c < C; ++c) {
        const float qc = qs[c * kAttRows + rr];
        const float* kr = ks + (size_t)c * T2p + lane;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const float df = qc - kr[32 * j];
          d[j] = fmaf(df, df, d[j]);
        }
      }
      const float lse = lse_in[(size_t)b * T1 + t1];
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (lane + 32 * j < T2) ga[j] -= expf(-temp * d[j] - lse) * sum_ga;
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int t2 = lane + 32 * j;
      if (t2 < T2) gd[row_off + t2] = ga[j];
      else ga[j] = 0.f;
      S += ga[j];
    }
    S = warp_sum(S);
    if (g_q) {
      // g_q[c] = -2 temp (q_c S - sum_t2 g_d k[c][t2])
      for (int c = 0; c < C; ++c) {
        const float* kr = ks + (size_t)c * T2p + lane;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc = fmaf(ga[j], kr[32 * j], acc);
        acc = warp_sum(acc);
        if (lane == 0) g_q[((size_t)b * C + c) * T1 + t1] = -2.f * temp * (qs[c * kAttRows + rr] * S - acc);
      }
    }
  }
}

// backward pass 2: g_k[c][t2] = 2 temp (sum_t1 g_d[t1][t2] q[c][t1] - k[c][t2] sum_t1 g_d[t1][t2])
__global__ void __launch_bounds__(256) convattn_bwd_cols_kernel(const float* __restrict__ q_enc, const float* __restrict__ k_enc,
                                                                const float* __restrict__ gd, int C, int T1, int T2, float temp,
                                                                float* __restrict__ g_k) {
  __shared__ float gds[32][33];   // [t1][t2]
  __shared__ float qsm[128][33];  // [c][t1]
  const int b = blockIdx.y, t2_0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // tx: t2 in tile; ty: channel group
  const int cpg = (C + 7) / 8;                              // channels per group (<= 16)
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  float colsum = 0.f;
  const float* q = q_enc + (size_t)b * C * T1;
  for (int t1_0 = 0; t1_0 < T1; t1_0 += 32) {
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
      const int r = i >> 5, cc = i & 31;
      const int t1 = t1_0 + r, t2 = t2_0 + cc;
      gds[r][cc] = (t1 < T1 && t2 < T2) ? gd[((size_t)b * T1 + t1) * T2 + t2] : 0.f;
    }
    for (int i = threadIdx.x; i < C * 32; i += 256) {
      const int c = i >> 5, r = i & 31;
      qsm[c][r] = (t1_0 + r) < T1 ? q[(size_t)c * T1 + t1_0 + r] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      const float g = gds[r][tx];
      colsum += g;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int c = ty * cpg + i;
        if (i < cpg && c < C) acc[i] = fmaf(g, qsm[c][r], acc[i]);
      }
    }
    __syncthreads();
  }
  const int t2 = t2_0 + tx;
  if (t2 < T2) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = ty * cpg + i;
      if (i < cpg && c < C) {
        const size_t o = ((size_t)b * C + c) * T2 + t2;
        g_k[o] = 2.f * temp * (acc[i] - k_enc[o] * colsum);
      }
    }
  }
}

template <int NJ>
static int launch_fwd(const float* q, const float* k, const float* prior, const int64_t* key_lens, int B, int C, int T1,
                      int T2, float temp, float* attn, float* logprob, float* lse, cudaStream_t st) {
  constexpr int kFwdRows = FwdCfg<NJ>::kRows;
  const size_t smem = ((size_t)C * NJ * 32 + (size_t)C * kFwdRows + (size_t)NJ * 32) * sizeof(float);
  if (smem > (size_t)kSmemBudget) return RADTTS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    RB_CUDA(cudaFuncSetAttribute(convattn_fwd_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  // ~2 CTAs per SM in total; every CTA of an utterance stages its keys once and walks its share of the frame tiles
  // sized to ONE resident wave (1 or 2 CTAs per SM, whatever the key stage leaves room for)
  const int n_tiles = ceil_div(T1, kFwdRows);
  int per_sm = (int)((size_t)233472 / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 2048 / kFwdThreads ? 2048 / kFwdThreads : per_sm);
  int per_utt = per_sm * kNumSMs / B;
  per_utt = per_utt < 1 ? 1 : (per_utt > n_tiles ? n_tiles : per_utt);
  dim3 grid(per_utt, B);
  convattn_fwd_kernel<NJ><<<grid, kFwdThreads, smem, st>>>(q, k, prior, key_lens, C, T1, T2, temp, attn, logprob, lse);
  return after_launch();
}
template <int NJ>
static int launch_bwd_rows(const float* q, const float* k, const float* lse, const float* attn, const float* g_attn,
                           const float* g_lp, int has_prior, int B, int C, int T1, int T2, float temp, float* gd,
                           float* g_q, cudaStream_t st) {
  const size_t smem = ((size_t)C * NJ * 32 + (size_t)C * kAttRows) * sizeof(float);
  if (smem > (size_t)kSmemBudget) return RADTTS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    RB_CUDA(cudaFuncSetAttribute(convattn_bwd_rows_kernel<NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  dim3 grid(ceil_div(T1, kAttRows), B);
  convattn_bwd_rows_kernel<NJ><<<grid, kAttThreads, smem, st>>>(q, k, lse, attn, g_attn, g_lp, has_prior, C, T1, T2, temp,
                                                               gd, g_q);
  return after_launch();
}

}  // namespace rb

using namespace rb;

extern "C" int radtts_convattn_forward(const float* q_enc, const float* k_enc, const float* prior,
                                       const int64_t* key_lens, int B, int C, int T1, int T2, float temp, float* attn,
                                       float* attn_logprob, float* lse, void* stream) {
  if (!q_enc || !k_enc || !attn || !attn_logprob || B <= 0 || C <= 0 || T1 <= 0 || T2 <= 0) return RADTTS_ERR_INVALID_ARG;
  if (prior && !lse) return RADTTS_ERR_INVALID_ARG;
  const int nj = ceil_div(T2, 32);
  if (nj > kAttMaxJ || C > 128) return RADTTS_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
#define RB_CALL_FWD(N) launch_fwd<N>(q_enc, k_enc, prior, key_lens, B, C, T1, T2, temp, attn, attn_logprob, lse, st)
  switch (nj) {
    case 1: return RB_CALL_FWD(1); case 2: return RB_CALL_FWD(2); case 3: return RB_CALL_FWD(3); case 4: return RB_CALL_FWD(4);
    case 5: return RB_CALL_FWD(5); case 6: return RB_CALL_FWD(6); case 7: case 8: return RB_CALL_FWD(8);
    case 9: case 10: return RB_CALL_FWD(10); case 11: case 12: return RB_CALL_FWD(12); case 13: case 14: return RB_CALL_FWD(14);
    default: return RB_CALL_FWD(18);
  }
#undef RB_CALL_FWD
}

extern "C" int radtts_convattn_backward(const float* q_enc, const float* k_enc, const float* lse, const float* attn,
                                        const float* g_attn, const float* g_logprob, int has_prior, int B, int C, int T1,
                                        int T2, float temp, float* gd_ws, float* g_q, float* g_k, void* stream) {
  if (!q_enc || !k_enc || !attn || !gd_ws || !g_k || B <= 0 || C <= 0 || T1 <= 0 || T2 <= 0) return RADTTS_ERR_INVALID_ARG;
  if (has_prior && !lse) return RADTTS_ERR_INVALID_ARG;
  const int nj = ceil_div(T2, 32);
  if (nj > kAttMaxJ || C > 128) return RADTTS_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
#define RB_CALL_BWD(N) launch_bwd_rows<N>(q_enc, k_enc, lse, attn, g_attn, g_logprob, has_prior, B, C, T1, T2, temp, gd_ws, g_q, st)
  switch (nj) {
    case 1: rc = RB_CALL_BWD(1); break; case 2: rc = RB_CALL_BWD(2); break; case 3: rc = RB_CALL_BWD(3); break;
    case 4: rc = RB_CALL_BWD(4); break; case 5: rc = RB_CALL_BWD(5); break; case 6: rc = RB_CALL_BWD(6); break;
    case 7: case 8: rc = RB_CALL_BWD(8); break; case 9: case 10: rc = RB_CALL_BWD(10); break;
    case 11: case 12: rc = RB_CALL_BWD(12); break; case 13: case 14: rc = RB_CALL_BWD(14); break;
    default: rc = RB_CALL_BWD(18); break;
  }
#undef RB_CALL_BWD
  RB_TRY(rc);
  dim3 grid(ceil_div(T2, 32), B);
  convattn_bwd_cols_kernel<<<grid, 256, 0, st>>>(q_enc, k_enc, gd_ws, C, T1, T2, temp, g_k);
  return after_launch();
}
