// Attention CTC loss (reference loss.py:111-135): for every utterance the alignment logits
// attn_logprob[b, 0, :T_b, :K_b] are extended with a constant blank logit, log-softmax'ed over the K_b + 1 classes
// and scored with CTC against the target 1, 2, ..., K_b.  The reference does this in a Python loop over the batch
// (pad + log_softmax + nn.CTCLoss per utterance, two host syncs each); here one CTA per utterance runs the fused
// log-softmax, the alpha recursion, the beta recursion and the gradient w.r.t. the logits in a single launch:
//     dL_b / d a[t][k] = softmax_t(k) - posterior_t(label k)
// (the classic CTC-through-softmax gradient), already scaled by 1 / (K_b * B) = the reference's reduction
// (nn.CTCLoss 'mean' on a batch of one divides by the target length, loss.py:133-134 averages over the batch).
// The DP lattice (T_b x (2 K_b + 1) states) has the same shape as the MAS lattice of kernel 1.
#include <math_constants.h>

#include "common.cuh"

namespace rb {

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m));
}
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + logf(expf(a - m) + expf(b - m) + expf(c - m));
}

// logits (B,1,T1,T2); alpha_ws (B, T1, S_max) with S_max = 2*T2+1; grad (B,1,T1,T2) out; losses (B) out
__global__ void __launch_bounds__(1024) attn_ctc_kernel(const float* __restrict__ logits, const int64_t* __restrict__ in_lens,
                                                        const int64_t* __restrict__ out_lens, int B, int T1, int T2,
                                                        float blank, float* __restrict__ alpha_ws,
                                                        float* __restrict__ grad, float* __restrict__ losses) {
  extern __shared__ float sm[];
  const int b = blockIdx.x;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int K = (int)min((long long)max((long long)in_lens[b], 0ll), (long long)T2);
  const int Tb = (int)min((long long)max((long long)out_lens[b], 0ll), (long long)T1);
  const int S = 2 * K + 1, Smax = 2 * T2 + 1;
  float* lse_s = sm;                      // [T1]
  float* buf = sm + T1;                   // [2][Smax + 2]  (one -inf guard cell on each side)
  const float* a = logits + (size_t)b * T1 * T2;
  float* g = grad + (size_t)b * T1 * T2;
  float* al = alpha_ws + (size_t)b * T1 * Smax;
  const float NEG = -CUDART_INF_F;

  // zero the gradient slab (frames >= Tb, tokens >= K and the zero_infinity case rely on it)
  for (size_t i = tid; i < (size_t)T1 * T2; i += nthr) g[i] = 0.f;
  if (K == 0 || Tb == 0) {
    if (tid == 0) losses[b] = 0.f;
    return;
  }
  // per-frame log-sum-exp over {blank, a[t][0..K)}: one warp per frame
  {
    const int warp = tid >> 5, lane = tid & 31, nw = nthr >> 5;
    for (int t = warp; t < Tb; t += nw) {
      float m = blank;
      for (int k = lane; k < K; k += 32) m = fmaxf(m, a[(size_t)t * T2 + k]);
#pragma unroll
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float s = lane == 0 ? expf(blank - m) : 0.f;
      for (int k = lane; k < K; k += 32) s += expf(a[(size_t)t * T2 + k] - m);
#pragma unroll
      for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) lse_s[t] = m + logf(s);
    }
  }
  for (int i = tid; i < 2 * (Smax + 2); i += nthr) buf[i] = NEG;
  __syncthreads();

  const int s = tid;                       // state owned by this thread
  const bool live = s < S;
  const bool is_label = (s & 1) != 0;
  const int k = (s - 1) >> 1;              // token index of a label state
  auto lp_of = [&](int t) -> float {
    const float x = is_label ? a[(size_t)t * T2 + k] : blank;
    return x - lse_s[t];
  };
  float* b0 = buf + 1;                     // buf[p][s] = b0[p * (Smax + 2) + s], valid for s in [-1, Smax]
  const int ld = Smax + 2;

  // ---------------- alpha ----------------
  float lp_next = live ? lp_of(0) : NEG;
  float alpha = NEG;
  for (int t = 0; t < Tb; ++t) {
    const float lp = lp_next;
    if (live && t + 1 < Tb) lp_next = lp_of(t + 1);
    if (t == 0) {
      alpha = (live && s <= 1) ? lp : NEG;
    } else if (live) {
      const float* prev = b0 + ((t - 1) & 1) * ld;
      const float p2 = (is_label && s >= 3) ? prev[s - 2] : NEG;
      alpha = lp + lse3(prev[s], prev[s - 1], p2);
    }
    if (live) {
      b0[(t & 1) * ld + s] = alpha;
      al[(size_t)t * Smax + s] = alpha;
    }
    __syncthreads();
  }
  const float* last = b0 + ((Tb - 1) & 1) * ld;
  const float ll = lse2(last[S - 1], S >= 2 ? last[S - 2] : NEG);   // log p(target | input)
  const bool finite = ll > NEG;                                     // zero_infinity=True
  if (tid == 0) losses[b] = finite ? -ll : 0.f;
  __syncthreads();
  if (!finite) return;

  // ---------------- beta + gradient ----------------
  for (int i = tid; i < 2 * ld; i += nthr) buf[i] = NEG;
  __syncthreads();
  const float scale = 1.f / ((float)K * (float)B);
  float lp_cur = live ? lp_of(Tb - 1) : NEG;
  float beta = NEG;
  for (int t = Tb - 1; t >= 0; --t) {
    const float lp = lp_cur;
    if (live && t > 0) lp_cur = lp_of(t - 1);
    if (t == Tb - 1) {
      beta = (live && s >= S - 2) ? lp : NEG;
    } else if (live) {
      const float* nxt = b0 + ((t + 1) & 1) * ld;
      const float n2 = (is_label && s + 2 < S) ? nxt[s + 2] : NEG;
      beta = lp + lse3(nxt[s], nxt[s + 1], n2);
    }
    if (live) {
      b0[(t & 1) * ld + s] = beta;
      if (is_label) {
        const float post = expf(al[(size_t)t * Smax + s] + beta - lp - ll);   // posterior of this label at frame t
        g[(size_t)t * T2 + k] = (expf(lp) - post) * scale;
      }
    }
    __syncthreads();
  }
}

}  // namespace rb

using namespace rb;

extern "C" size_t radtts_attn_ctc_workspace_bytes(int B, int T1, int T2) {
  if (B <= 0 || T1 <= 0 || T2 <= 0) return 0;
  return (size_t)B * T1 * (2 * T2 + 1) * sizeof(float);
}

extern "C" int radtts_attn_ctc(const float* attn_logprob, const int64_t* in_lens, const int64_t* out_lens, int B, int T1,
                               int T2, float blank_logprob, float* losses, float* grad, void* ws, size_t ws_bytes,
                               void* stream) {
  if (!attn_logprob || !in_lens || !out_lens || !losses || !grad || !ws || B <= 0 || T1 <= 0 || T2 <= 0)
    return RADTTS_ERR_INVALID_ARG;
  if (2 * T2 + 1 > 1024) return RADTTS_ERR_UNSUPPORTED;
  if (ws_bytes < radtts_attn_ctc_workspace_bytes(B, T1, T2)) return RADTTS_ERR_WORKSPACE;
  const int threads = round_up(2 * T2 + 1, 32);
  const size_t smem = ((size_t)T1 + 2 * (size_t)(2 * T2 + 3)) * sizeof(float);
  if (smem > (size_t)kSmemBudget) return RADTTS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    RB_CUDA(cudaFuncSetAttribute(attn_ctc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  attn_ctc_kernel<<<B, threads, smem, (cudaStream_t)stream>>>(attn_logprob, in_lens, out_lens, B, T1, T2, blank_logprob,
                                                             reinterpret_cast<float*>(ws), grad, losses);
  return after_launch();
}
