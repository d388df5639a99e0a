// Persistent bidirectional LSTM recurrence ("next" row f-2 of SURVEY section 8: the context BiLSTM of
// RADTTS.preprocess_context, reference radtts.py:262-302, and the text-encoder BiLSTM, common.py:359-371).
//
// cuDNN runs a variable-length (packed) LSTM as ~10 tiny kernels per time step: at T' = 400 that is >4000 launches
// and ~20 ms per train step -- more than the whole decoder flow stack.  Here the time loop lives inside ONE
// cooperative kernel per pass: the input projection W_ih x_t (+ biases) for all t is a plain GEMM done beforehand;
// each CTA owns a slice of U hidden units of one direction, keeps its rows of W_hh resident in shared memory for
// the whole sequence, and the CTAs of a direction exchange h_t through L2 with one grid barrier per step.
// Variable lengths follow packed-sequence semantics: state and output are zero for t >= len[b], so the reverse
// direction starts each utterance at its own last frame.
//
// fp32 throughout (state, weights, accumulation).  Gate order i, f, g, o as in torch.nn.LSTM.
#include "common.cuh"
#include <cuda_bf16.h>

#include "ptx.cuh"
#include "lstm_cluster.cuh"

namespace rb {

constexpr int kLstmThreads = 256;   // 32 batch lanes x 8 slices
constexpr int kLstmMaxB = 32;
constexpr int kLstmLd = 32;          // exchange buffers are stored transposed ([k][32 batch lanes]) in GLOBAL memory, so one
                                    // 1-D bulk async copy (TMA) per step lands them in smem in the conflict-free layout

struct LstmDims {
  int T, B, H, U, G;   // U units per CTA, G CTAs per direction (G * U >= H)
};

__device__ __forceinline__ float ldcg(const float* p) { return __ldcg(p); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// Split barrier among the CTAs of one direction: arrive right after the exchange stores, do the remaining (bulky,
// nobody-waits-for-them) stores of the step, then wait.  Keeps those stores off the barrier's critical path.
__device__ __forceinline__ void dir_barrier_arrive(unsigned* counter) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
  }
}
__device__ __forceinline__ void dir_barrier_wait(unsigned* counter, unsigned nblocks, unsigned& phase) {
  if (threadIdx.x == 0) {
    const unsigned target = (phase + 1) * nblocks;
    unsigned spins = 0;
    while (true) {
      unsigned v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (v >= target) break;
      if (++spins > (1u << 28)) __trap();
    }
    __threadfence();
  }
  ++phase;
  __syncthreads();
}

// ----------------------------------------------------------------------------------------------------------
// forward:  gates = gx[t] + W_hh h_{t-1};  c = f c + i g;  h = o tanh(c)
//   gx     [2][T][B][4H]   x-projection + b_ih + b_hh per direction
//   whh    [2][4H][H]
//   h_all  [T][B][2H]      output (direction d in columns [dH, (d+1)H)), zeros for t >= len
//   gates_save [2][T][B][4H] (i, f, g, o after the non-linearities), c_save [2][T][B][H]   (training only)
//   hbuf   [2][2][B][H]    ping-pong exchange buffer, zero-initialised;  counters [2] zero-initialised
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLstmThreads, 1)
lstm_fwd_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const int* __restrict__ lens, LstmDims d,
                float* __restrict__ h_all, float* __restrict__ gates_save, float* __restrict__ c_save,
                float* __restrict__ hbuf, unsigned* __restrict__ counters) {
  extern __shared__ __align__(128) float sm[];
  const int T = d.T, B = d.B, H = d.H, U = d.U;
  const int dir = blockIdx.x / d.G, cta = blockIdx.x % d.G;
  const int u0 = cta * U;
  const int R = 4 * U;                       // gate rows owned by this CTA
  float* ws = sm;                            // [H][R]   (k-major, rows contiguous)
  float* hs = ws + (size_t)H * R;            // [H][32]  h_{t-1} transposed
  float* red = hs + (size_t)H * kLstmLd;     // [16 k-slices][R][32]
  __shared__ __align__(8) uint64_t bar[8];   // one per warp (= 2 k-slices): a warp starts as soon as ITS part of h landed
  const int tid = threadIdx.x;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    mbar_fence_init();
  }
  const int b = tid & 31;                    // batch lane of the (unit, batch) pair this thread finalises
  // GEMM mapping: 8 x 8 register tiles (gate rows x batch), 16 tiles cover 32 x 32, 16 k-slices of ~H/16.
  // smem -> register traffic per FMA drops 4x vs one-row-of-batch per thread, which was the limiter.
  const int tile = tid & 15, kslice = tid >> 4;
  const int rt = (tile >> 2) * 8, bt = (tile & 3) * 8;
  const int warp = tid >> 5;
  const float* W = whh + (size_t)dir * 4 * H * H;
  for (int i = tid; i < H * R; i += kLstmThreads) {
    const int k = i / R, r = i % R;          // r = gate * U + ul
    const int gate = r / U, ul = r % U;
    const int u = u0 + ul;
    ws[i] = u < H ? W[(size_t)(gate * H + u) * H + k] : 0.f;
  }
  // state owned by thread (ul = slice' ..): after the reduction thread `tid` finalises unit ul = tid / 32, batch b
  float c_state = 0.f;
  const int ul_own = tid >> 5;               // U <= 8 so that U * 32 <= 256 threads own (unit, batch) pairs
  const int len_b = b < B ? lens[b] : 0;
  unsigned phase = 0;
  const int kchunk = (H + 15) / 16;          // k-slice length; a warp owns slices 2*warp, 2*warp + 1
  __syncthreads();

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const float* hprev = hbuf + ((size_t)(dir * 2 + (s & 1)) * kLstmMaxB) * H;   // [H][32]
    float* hnext = hbuf + ((size_t)(dir * 2 + ((s + 1) & 1)) * kLstmMaxB) * H;
    // stage h_{t-1}: one bulk async copy of the whole [H][32] exchange block (written by the other CTAs of this
    // direction before the barrier; the proxy fence orders the async-proxy read after those generic-proxy writes)
    if (tid == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      for (int sl = 0; sl < 8; ++sl) {
        const int ka = min(2 * sl * kchunk, H), kb = min(ka + 2 * kchunk, H);
        if (kb > ka) {
          const uint32_t bytes = (uint32_t)((kb - ka) * kLstmMaxB * sizeof(float));
          mbar_arrive_expect_tx(&bar[sl], bytes);
          bulk_g2s(hs + (size_t)ka * kLstmLd, hprev + (size_t)ka * kLstmMaxB, bytes, &bar[sl]);
        } else {
          mbar_arrive(&bar[sl]);
        }
      }
    }
    // x-projection of this step for the (unit, batch) this thread finalises: issued before the wait
    float gxv[4] = {0.f, 0.f, 0.f, 0.f};
    if (ul_own < U && u0 + ul_own < H && b < B) {
#pragma unroll
      for (int gate = 0; gate < 4; ++gate)
        gxv[gate] = gx[(((size_t)dir * T + t) * B + b) * 4 * H + gate * H + u0 + ul_own];
    }
    mbar_wait(&bar[warp], (uint32_t)(s & 1));
    // partial products: thread (tile, kslice) accumulates an 8 x 8 block over its k-slice
    {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      const int k0 = min(kslice * kchunk, H), k1 = min(k0 + kchunk, H);
      if (rt < R) {
        for (int k = k0; k < k1; ++k) {
          const float4 w0 = *reinterpret_cast<const float4*>(ws + (size_t)k * R + rt);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + (size_t)k * R + rt + 4);
          const float4 h0 = *reinterpret_cast<const float4*>(hs + (size_t)k * kLstmLd + bt);
          const float4 h1 = *reinterpret_cast<const float4*>(hs + (size_t)k * kLstmLd + bt + 4);
          const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
          const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(w[i], h[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (rt + i >= R) break;
          float* rr = red + ((size_t)kslice * R + rt + i) * kLstmMaxB + bt;
          *reinterpret_cast<float4*>(rr) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
          *reinterpret_cast<float4*>(rr + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
        }
      }
    }
    __syncthreads();
    // finalise (unit ul_own, batch b)
    bool wrote = false;
    float hval = 0.f, ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, cn = 0.f;
    if (ul_own < U) {
      const int u = u0 + ul_own;
      if (u < H && b < B) {
        float g4[4];
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
          float v = gxv[gate];
          const int r = gate * U + ul_own;
#pragma unroll
          for (int sl = 0; sl < 16; ++sl) v += red[((size_t)sl * R + r) * kLstmMaxB + b];
          g4[gate] = v;
        }
        const bool on = t < len_b;
        ig = sigmoidf_(g4[0]); fg = sigmoidf_(g4[1]); gg = tanhf(g4[2]); og = sigmoidf_(g4[3]);
        cn = on ? fg * c_state + ig * gg : 0.f;
        hval = on ? og * tanhf(cn) : 0.f;
        c_state = cn;
        hnext[(size_t)u * kLstmMaxB + b] = hval;           // the only store other CTAs wait for
        asm volatile("fence.proxy.async;" ::: "memory");   // ... and they read it with bulk async copies
        wrote = true;
      }
    }
    dir_barrier_arrive(counters + dir);
    if (wrote) {
      const int u = u0 + ul_own;
      if (gates_save) {
        float* gs = gates_save + (((size_t)dir * T + t) * B + b) * 4 * H;
        gs[u] = ig; gs[H + u] = fg; gs[2 * H + u] = gg; gs[3 * H + u] = og;
        c_save[(((size_t)dir * T + t) * B + b) * H + u] = cn;
      }
      h_all[((size_t)t * B + b) * 2 * H + dir * H + u] = hval;
    }
    dir_barrier_wait(counters + dir, d.G, phase);
  }
}

// ----------------------------------------------------------------------------------------------------------
// backward:  walks time in the opposite order of the forward pass of that direction.
//   dh_all [T][B][2H] upstream gradient;  dgates_all [2][T][B][4H] out (pre-activation gate gradients)
//   whh [2][4H][H];  dgbuf [2][2][B][4H] ping-pong exchange of dgates_t (zero-initialised)
// ----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLstmThreads, 1)
lstm_bwd_kernel(const float* __restrict__ dh_all, const float* __restrict__ whh, const int* __restrict__ lens,
                const float* __restrict__ gates_save, const float* __restrict__ c_save, LstmDims d,
                float* __restrict__ dgates_all, float* __restrict__ dgbuf, unsigned* __restrict__ counters) {
  extern __shared__ __align__(128) float sm[];
  const int T = d.T, B = d.B, H = d.H, U = d.U;
  const int dir = blockIdx.x / d.G, cta = blockIdx.x % d.G;
  const int u0 = cta * U;
  const int J = 4 * H;
  float* wt = sm;                              // [J][U]   W_hh^T slice: wt[j][ul] = W[j][u0 + ul]
  const int JC = H / 2;                        // 8 chunks of dgates_{next} per step (H is even)
  const int NCHUNK = 8;
  float* dgs = wt + (size_t)J * 8;             // [2][JC][32] double-buffered chunks of dgates_{next} (transposed)
  float* red = dgs + (size_t)2 * JC * kLstmLd; // [32 j-slices][8 units][32]
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  unsigned nfill = 0;                          // chunk copies issued so far (buffer = nfill & 1, parity = (nfill >> 1) & 1)
  const int b = tid & 31;
  // GEMM mapping: 4 x 8 register tiles (units x batch), 8 tiles cover 8 x 32, 32 j-slices per chunk
  const int tile = tid & 7, jslice = tid >> 3;
  const int ut = (tile >> 2) * 4, bt = (tile & 3) * 8;
  const float* W = whh + (size_t)dir * 4 * H * H;
  for (int i = tid; i < J * 8; i += kLstmThreads) {
    const int j = i / 8, ul = i % 8;
    const int u = u0 + ul;
    wt[i] = (ul < U && u < H) ? W[(size_t)j * H + u] : 0.f;
  }
  const int ul_own = tid >> 5;
  const int len_b = b < B ? lens[b] : 0;
  float dc_state = 0.f;                        // dL/dc carried to the previous step, for (ul_own, b)
  unsigned phase = 0;
  __syncthreads();

  for (int s = 0; s < T; ++s) {
    // forward of this direction visited t_fwd(s') = dir ? T-1-s' : s'; backward visits them in reverse
    const int sf = T - 1 - s;
    const int t = dir == 0 ? sf : T - 1 - sf;
    const int t_prev = dir == 0 ? t - 1 : t + 1;   // the step whose c feeds this one (c_{t-1} in forward order)
    const float* dgnext = dgbuf + ((size_t)(dir * 2 + (s & 1)) * kLstmMaxB) * J;   // [J][32]
    float* dgcur = dgbuf + ((size_t)(dir * 2 + ((s + 1) & 1)) * kLstmMaxB) * J;
    const uint32_t chunk_bytes = (uint32_t)(JC * kLstmMaxB * sizeof(float));
    auto issue = [&](int chunk) {
      const unsigned buf = nfill & 1u;
      mbar_arrive_expect_tx(&bar[buf], chunk_bytes);
      bulk_g2s(dgs + (size_t)buf * JC * kLstmLd, dgnext + (size_t)chunk * JC * kLstmMaxB, chunk_bytes, &bar[buf]);
    };
    if (tid == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      issue(0);
    }
    // operands of the element-wise part, fetched while the copies are in flight
    const bool mine = ul_own < U && u0 + ul_own < H && b < B;
    const bool on = mine && t < len_b;
    float dh = 0.f, ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, cn = 0.f, cp = 0.f;
    if (mine) dh = dh_all[((size_t)t * B + b) * 2 * H + dir * H + u0 + ul_own];
    if (on) {
      const int u = u0 + ul_own;
      const float* gs = gates_save + (((size_t)dir * T + t) * B + b) * 4 * H;
      ig = gs[u]; fg = gs[H + u]; gg = gs[2 * H + u]; og = gs[3 * H + u];
      cn = c_save[(((size_t)dir * T + t) * B + b) * H + u];
      const bool has_prev = (t_prev >= 0 && t_prev < T) && (t_prev < len_b);
      cp = has_prev ? c_save[(((size_t)dir * T + t_prev) * B + b) * H + u] : 0.f;
    }
    // dh_rec[b][u] = sum_j dgates_next[b][j] W[j][u]
    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int chunk = 0; chunk < NCHUNK; ++chunk) {
      const unsigned buf = nfill & 1u, par = (nfill >> 1) & 1u;
      ++nfill;
      if (tid == 0 && chunk + 1 < NCHUNK) issue(chunk + 1);   // the other buffer was released by the sync of chunk - 1
      mbar_wait(&bar[buf], par);
      const float* dg_s = dgs + (size_t)buf * JC * kLstmLd;
      const int j0 = chunk * JC;
      const int jc = (JC + 31) / 32;
      const int ja = min(jslice * jc, JC), jb = min(ja + jc, JC);
      for (int jj = ja; jj < jb; ++jj) {
        const float4 w4 = *reinterpret_cast<const float4*>(wt + (size_t)(j0 + jj) * 8 + ut);
        const float4 g0 = *reinterpret_cast<const float4*>(dg_s + (size_t)jj * kLstmLd + bt);
        const float4 g1 = *reinterpret_cast<const float4*>(dg_s + (size_t)jj * kLstmLd + bt + 4);
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(w[i], g[j], acc[i][j]);
      }
      __syncthreads();   // everyone is done with this buffer before it is refilled two chunks later
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float* rr = red + ((size_t)jslice * 8 + ut + i) * kLstmMaxB + bt;
      *reinterpret_cast<float4*>(rr) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(rr + 4) = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
    }
    __syncthreads();
    float d4[4] = {0.f, 0.f, 0.f, 0.f};
    if (mine) {
      const int u = u0 + ul_own;
#pragma unroll
      for (int sl = 0; sl < 32; ++sl) dh += red[((size_t)sl * 8 + ul_own) * kLstmMaxB + b];
      float di = 0.f, df = 0.f, dg = 0.f, dout = 0.f;
      if (on) {
        const float tc = tanhf(cn);
        const float dc = dh * og * (1.f - tc * tc) + dc_state;
        dout = dh * tc * og * (1.f - og);
        di = dc * gg * ig * (1.f - ig);
        df = dc * cp * fg * (1.f - fg);
        dg = dc * ig * (1.f - gg * gg);
        dc_state = dc * fg;
      } else {
        dc_state = 0.f;
      }
      dgcur[(size_t)u * kLstmMaxB + b] = di;
      dgcur[(size_t)(H + u) * kLstmMaxB + b] = df;
      dgcur[(size_t)(2 * H + u) * kLstmMaxB + b] = dg;
      dgcur[(size_t)(3 * H + u) * kLstmMaxB + b] = dout;
      asm volatile("fence.proxy.async;" ::: "memory");
      d4[0] = di; d4[1] = df; d4[2] = dg; d4[3] = dout;
    }
    dir_barrier_arrive(counters + dir);
    if (mine) {
      const int u = u0 + ul_own;
      float* go = dgates_all + (((size_t)dir * T + t) * B + b) * J;
      go[u] = d4[0]; go[H + u] = d4[1]; go[2 * H + u] = d4[2]; go[3 * H + u] = d4[3];
    }
    dir_barrier_wait(counters + dir, d.G, phase);
  }
}

// ==========================================================================================================
// bf16 tensor-core variants (precision = RADTTS_PREC_BF16, i.e. under autocast -- where cuDNN would run the whole LSTM
// in half precision).  Only the recurrent product uses bf16 operands (W_hh and the exchanged h_{t-1} / dgates_{t+1});
// accumulation, gates, cell state, outputs and saved tensors stay fp32.
//   * the per-step GEMM of a CTA is tiny (32 gate rows x 32 batch x H): mma.sync.m16n8k16 with the operands laid out
//     in shared memory in FRAGMENT ORDER, so a thread fetches its A fragment with one LDS.128 and its B fragment with
//     one LDS.64, conflict-free; the SIMT version spent ~3 us per step on FMAs + a 16-way smem reduction;
//   * the exchange buffers live in global memory in that same fragment order, in bf16: half the bytes per step
//     (fwd 32 KB, bwd 128 KB per CTA) and one bulk async copy lands them ready to use.
// U = 8 units per CTA (one m16n8 C fragment row <-> one unit), H % 8 == 0 (K is zero-padded to a multiple of 16).
// ==========================================================================================================
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits (TF32 keeps 10).  Used by the kSplit = 2 variants,
// which replace the fp32 FMA recurrence wherever cuDNN itself would be allowed TF32 (torch.backends.cudnn.allow_tf32).
__device__ __forceinline__ void split_bf16(float x, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16(x));
  lo = x - hi;
}
__device__ __forceinline__ void pack_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  float ah, al, bh, bl;
  split_bf16(a, ah, al);
  split_bf16(b, bh, bl);
  hi = pack_bf16(ah, bh);
  lo = pack_bf16(al, bl);
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint4& a, const uint2& b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y));
}

constexpr int kLstmMmaFwdThreads = 128;   // 4 warps = 4 batch n-tiles of 8; each warp: 2 m-tiles x full K

// hbuf (bf16, fragment order): [2 dir][2 ping-pong][KS = H/16][4 n-tiles][32 lanes] uint2
template <int kSplit>
__global__ void __launch_bounds__(kLstmMmaFwdThreads, 1)
lstm_fwd_mma_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const int* __restrict__ lens, LstmDims d,
                    float* __restrict__ h_all, float* __restrict__ gates_save, float* __restrict__ c_save,
                    uint2* __restrict__ hbuf, unsigned* __restrict__ counters) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int T = d.T, B = d.B, H = d.H;
  const int KS = (H + 15) / 16;              // K = H padded to 16 with zero weights / zero state
  const int dir = blockIdx.x / d.G, cta = blockIdx.x % d.G;
  const int u0 = cta * 8;
  // kSplit == 2: every operand exists twice (hi part, then lo part, same layout): acc += Wh hh + Wh hl + Wl hh
  uint4* ws = reinterpret_cast<uint4*>(smraw);                       // [kSplit][2 m-tiles][KS][32]
  uint2* hs = reinterpret_cast<uint2*>(smraw + (size_t)kSplit * 2 * KS * 32 * sizeof(uint4));   // [kSplit][KS][4][32]
  const size_t ws_part = (size_t)2 * KS * 32, hs_part = (size_t)KS * 4 * 32;
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  // W slice in A-fragment order.  m-tile 0: rows 0-7 gate i, 8-15 gate f; m-tile 1: gate g, gate o (unit = row & 7)
  const float* W = whh + (size_t)dir * 4 * H * H;
  for (int i = tid; i < 2 * KS * 32; i += kLstmMmaFwdThreads) {
    const int ln = i & 31, ks = (i >> 5) % KS, mt = i / (32 * KS);
    const int gg = ln >> 2, qq = ln & 3;
    const int k = ks * 16 + 2 * qq;
    const float* r0 = W + (size_t)((2 * mt) * H + u0 + gg) * H;       // row gg   (gate 2 mt)
    const float* r1 = W + (size_t)((2 * mt + 1) * H + u0 + gg) * H;   // row gg+8 (gate 2 mt + 1)
    auto at = [&](const float* r, int kk) { return kk < H ? r[kk] : 0.f; };
    uint4 hi, lo;
    pack_split(at(r0, k), at(r0, k + 1), hi.x, lo.x);
    pack_split(at(r1, k), at(r1, k + 1), hi.y, lo.y);
    pack_split(at(r0, k + 8), at(r0, k + 9), hi.z, lo.z);
    pack_split(at(r1, k + 8), at(r1, k + 9), hi.w, lo.w);
    ws[i] = hi;
    if (kSplit == 2) ws[ws_part + i] = lo;
  }
  const int u = u0 + g;                       // the unit this thread finalises, for batch columns bcol, bcol + 1
  const int bcol = warp * 8 + 2 * q;
  int len_b[2];
  len_b[0] = bcol < B ? lens[bcol] : 0;
  len_b[1] = bcol + 1 < B ? lens[bcol + 1] : 0;
  float c_state[2] = {0.f, 0.f};
  unsigned phase = 0;
  const size_t hbytes = (size_t)kSplit * KS * 4 * 32 * sizeof(uint2);   // hi block, then lo block
  const int KS0 = KS / 2;
  // where (k = u, n = bcol + e) sits in the fragment-ordered exchange buffer (32-bit word index; u even lanes store)
  const int kk = u & 15;
  size_t xw[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int lane_dst = (2 * q + e) * 4 + ((kk & 7) >> 1);
    xw[e] = (((size_t)(u >> 4) * 4 + warp) * 32 + lane_dst) * 2 + (kk >> 3);
  }
  __syncthreads();

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const uint2* hprev = hbuf + (size_t)(dir * 2 + (s & 1)) * (hbytes / sizeof(uint2));
    uint32_t* hnext = reinterpret_cast<uint32_t*>(hbuf + (size_t)(dir * 2 + ((s + 1) & 1)) * (hbytes / sizeof(uint2)));
    if (tid == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      // two copies per operand part: k-steps [0, KS0) and [KS0, KS); bar[0] completes when every first half is in
      const uint32_t b0 = (uint32_t)((size_t)KS0 * 4 * 32 * sizeof(uint2));
      const uint32_t b1 = (uint32_t)(hs_part * sizeof(uint2)) - b0;
      mbar_arrive_expect_tx(&bar[0], b0 * kSplit);
      mbar_arrive_expect_tx(&bar[1], b1 * kSplit);
#pragma unroll
      for (int part = 0; part < kSplit; ++part) {
        bulk_g2s(hs + part * hs_part, hprev + part * hs_part, b0, &bar[0]);
        bulk_g2s(hs + part * hs_part + (size_t)KS0 * 4 * 32, hprev + part * hs_part + (size_t)KS0 * 4 * 32, b1, &bar[1]);
      }
    }
    float gxv[4][2];
#pragma unroll
    for (int gate = 0; gate < 4; ++gate)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        gxv[gate][e] = (bcol + e < B) ? gx[(((size_t)dir * T + t) * B + bcol + e) * 4 * H + gate * H + u] : 0.f;
    float acc[2][2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[i][j][r] = 0.f;
    auto mma_step = [&](int ks) {
      const uint2 bh = hs[((size_t)ks * 4 + warp) * 32 + lane];
      const uint4 a0 = ws[(size_t)ks * 32 + lane], a1 = ws[((size_t)KS + ks) * 32 + lane];
      mma_bf16_16816(acc[0][ks & 1], a0, bh);
      mma_bf16_16816(acc[1][ks & 1], a1, bh);
      if (kSplit == 2) {
        const uint2 bl = hs[hs_part + ((size_t)ks * 4 + warp) * 32 + lane];
        mma_bf16_16816(acc[0][ks & 1], a0, bl);
        mma_bf16_16816(acc[1][ks & 1], a1, bl);
        mma_bf16_16816(acc[0][ks & 1], ws[ws_part + (size_t)ks * 32 + lane], bh);
        mma_bf16_16816(acc[1][ks & 1], ws[ws_part + ((size_t)KS + ks) * 32 + lane], bh);
      }
    };
    mbar_wait(&bar[0], (uint32_t)(s & 1));
#pragma unroll 4
    for (int ks = 0; ks < KS0; ++ks) mma_step(ks);
    mbar_wait(&bar[1], (uint32_t)(s & 1));
#pragma unroll 4
    for (int ks = KS0; ks < KS; ++ks) mma_step(ks);
    // C fragment: [0],[1] = (row g, cols 2q, 2q+1), [2],[3] = (row g + 8, cols 2q, 2q+1)
    float hval[2], ig[2], fg[2], gg[2], og[2], cn[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float pi = acc[0][0][e] + acc[0][1][e] + gxv[0][e];
      const float pf = acc[0][0][2 + e] + acc[0][1][2 + e] + gxv[1][e];
      const float pg = acc[1][0][e] + acc[1][1][e] + gxv[2][e];
      const float po = acc[1][0][2 + e] + acc[1][1][2 + e] + gxv[3][e];
      const bool on = t < len_b[e];
      ig[e] = sigmoidf_(pi); fg[e] = sigmoidf_(pf); gg[e] = tanhf(pg); og[e] = sigmoidf_(po);
      cn[e] = on ? fg[e] * c_state[e] + ig[e] * gg[e] : 0.f;
      hval[e] = on ? og[e] * tanhf(cn[e]) : 0.f;
      c_state[e] = cn[e];
      // exchange: units u (even g) and u + 1 (lane + 4) share a 32-bit word
      const float partner = __shfl_down_sync(0xffffffffu, hval[e], 4);
      if ((g & 1) == 0) {
        uint32_t hi, lo;
        pack_split(hval[e], partner, hi, lo);
        hnext[xw[e]] = hi;
        if (kSplit == 2) hnext[hs_part * 2 + xw[e]] = lo;   // the lo block follows the hi block (hs_part uint2 = 2 words each)
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    dir_barrier_arrive(counters + dir);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int b = bcol + e;
      if (b < B) {
        if (gates_save) {
          float* gs = gates_save + (((size_t)dir * T + t) * B + b) * 4 * H;
          gs[u] = ig[e]; gs[H + u] = fg[e]; gs[2 * H + u] = gg[e]; gs[3 * H + u] = og[e];
          c_save[(((size_t)dir * T + t) * B + b) * H + u] = cn[e];
        }
        h_all[((size_t)t * B + b) * 2 * H + dir * H + u] = hval[e];
      }
    }
    dir_barrier_wait(counters + dir, d.G, phase);
  }
}

constexpr int kLstmMmaBwdThreads = 256;   // 8 warps split K = 4H; each warp: 2 batch m-tiles x 8 units

// dgbuf (bf16, A-fragment order): [2 dir][2 ping-pong][KS = 4H/16][2 m-tiles][32 lanes] uint4
template <int kSplit>
__global__ void __launch_bounds__(kLstmMmaBwdThreads, 1)
lstm_bwd_mma_kernel(const float* __restrict__ dh_all, const float* __restrict__ whh, const int* __restrict__ lens,
                    const float* __restrict__ gates_save, const float* __restrict__ c_save, LstmDims d,
                    float* __restrict__ dgates_all, uint4* __restrict__ dgbuf, unsigned* __restrict__ counters) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int T = d.T, B = d.B, H = d.H;
  const int J = 4 * H, KS = J / 16, KSW = (KS + 7) / 8;   // H % 8 == 0 -> J % 16 == 0; warp w owns k-steps [w KSW, ..)
  const int dir = blockIdx.x / d.G, cta = blockIdx.x % d.G;
  const int u0 = cta * 8;
  // kSplit == 2: hi parts, then lo parts (same layout): acc += Ah Bh + Ah Bl + Al Bh
  uint2* wt = reinterpret_cast<uint2*>(smraw);                                                     // [kSplit][KS][32]  B fragments
  uint4* dgs = reinterpret_cast<uint4*>(smraw + (size_t)kSplit * KS * 32 * sizeof(uint2));          // [kSplit][KS][2][32] A fragments
  float* red = reinterpret_cast<float*>(smraw + (size_t)kSplit * (KS * 32 * sizeof(uint2) + (size_t)KS * 2 * 32 * sizeof(uint4)));
  const size_t wt_part = (size_t)KS * 32, dg_part = (size_t)KS * 2 * 32;
  __shared__ __align__(8) uint64_t bar[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  if (tid == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    mbar_fence_init();
  }
  const float* W = whh + (size_t)dir * 4 * H * H;
  for (int i = tid; i < KS * 32; i += kLstmMmaBwdThreads) {
    const int ln = i & 31, ks = i >> 5;
    const int gg = ln >> 2, qq = ln & 3;
    const float* c = W + (size_t)(ks * 16 + 2 * qq) * H + u0 + gg;     // B[k = j][n = unit gg]
    uint2 hi, lo;
    pack_split(c[0], c[H], hi.x, lo.x);
    pack_split(c[(size_t)8 * H], c[(size_t)9 * H], hi.y, lo.y);
    wt[i] = hi;
    if (kSplit == 2) wt[wt_part + i] = lo;
  }
  // element-wise owner: unit ul = tid & 7, batch b = tid >> 3
  const int ul = tid & 7, b = tid >> 3;
  const int u = u0 + ul;
  const bool mine = b < B;
  const int len_b = mine ? lens[b] : 0;
  float dc_state = 0.f;
  unsigned phase = 0;
  const size_t dwords = (size_t)kSplit * KS * 2 * 32;   // uint4 per exchange buffer (hi block, then lo block)
  // where (m = b, k = gate * H + u) sits in the exchange buffer (32-bit word index, even ul lanes store)
  size_t xw[4];
  {
    const int mt = b >> 4, r = b & 15;
#pragma unroll
    for (int gate = 0; gate < 4; ++gate) {
      const int j = gate * H + u, kk = j & 15;
      const int lane_dst = (r & 7) * 4 + ((kk & 7) >> 1);
      const int reg = (r >> 3) + 2 * (kk >> 3);
      xw[gate] = (((size_t)(j >> 4) * 2 + mt) * 32 + lane_dst) * 4 + reg;
    }
  }
  __syncthreads();

  for (int s = 0; s < T; ++s) {
    const int sf = T - 1 - s;
    const int t = dir == 0 ? sf : T - 1 - sf;
    const int t_prev = dir == 0 ? t - 1 : t + 1;
    const uint4* dgnext = dgbuf + (size_t)(dir * 2 + (s & 1)) * dwords;
    uint32_t* dgcur = reinterpret_cast<uint32_t*>(dgbuf + (size_t)(dir * 2 + ((s + 1) & 1)) * dwords);
    if (tid == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");
      for (int w = 0; w < 8; ++w) {
        const int ka = min(w * KSW, KS), kb = min(ka + KSW, KS);
        if (kb > ka) {
          const uint32_t cb = (uint32_t)((size_t)(kb - ka) * 2 * 32 * sizeof(uint4));
          mbar_arrive_expect_tx(&bar[w], cb * kSplit);
#pragma unroll
          for (int part = 0; part < kSplit; ++part)
            bulk_g2s(dgs + part * dg_part + (size_t)ka * 2 * 32, dgnext + part * dg_part + (size_t)ka * 2 * 32, cb, &bar[w]);
        } else {
          mbar_arrive(&bar[w]);
        }
      }
    }
    const bool on = mine && t < len_b;
    float dh = 0.f, ig = 0.f, fg = 0.f, gg = 0.f, og = 0.f, cn = 0.f, cp = 0.f;
    if (mine) dh = dh_all[((size_t)t * B + b) * 2 * H + dir * H + u];
    if (on) {
      const float* gs = gates_save + (((size_t)dir * T + t) * B + b) * 4 * H;
      ig = gs[u]; fg = gs[H + u]; gg = gs[2 * H + u]; og = gs[3 * H + u];
      cn = c_save[(((size_t)dir * T + t) * B + b) * H + u];
      const bool has_prev = (t_prev >= 0 && t_prev < T) && (t_prev < len_b);
      cp = has_prev ? c_save[(((size_t)dir * T + t_prev) * B + b) * H + u] : 0.f;
    }
    // partial dh_rec[m = batch][n = unit] over this warp's K range
    float acc[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[i][r] = 0.f;
    mbar_wait(&bar[warp], (uint32_t)(s & 1));
#pragma unroll 4
    for (int ks = min(warp * KSW, KS), ke = min(ks + KSW, KS); ks < ke; ++ks) {
      const uint2 bh = wt[(size_t)ks * 32 + lane];
      const uint4 a0 = dgs[((size_t)ks * 2) * 32 + lane], a1 = dgs[((size_t)ks * 2 + 1) * 32 + lane];
      mma_bf16_16816(acc[0], a0, bh);
      mma_bf16_16816(acc[1], a1, bh);
      if (kSplit == 2) {
        const uint2 bl = wt[wt_part + (size_t)ks * 32 + lane];
        mma_bf16_16816(acc[0], a0, bl);
        mma_bf16_16816(acc[1], a1, bl);
        mma_bf16_16816(acc[0], dgs[dg_part + ((size_t)ks * 2) * 32 + lane], bh);
        mma_bf16_16816(acc[1], dgs[dg_part + ((size_t)ks * 2 + 1) * 32 + lane], bh);
      }
    }
    // red[warp][batch][unit]
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      float* rr = red + ((size_t)warp * 32 + mt * 16 + g) * 8 + 2 * q;
      *reinterpret_cast<float2*>(rr) = make_float2(acc[mt][0], acc[mt][1]);
      *reinterpret_cast<float2*>(rr + 8 * 8) = make_float2(acc[mt][2], acc[mt][3]);
    }
    __syncthreads();
    float d4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int w = 0; w < 8; ++w) dh += red[((size_t)w * 32 + b) * 8 + ul];
    if (on) {
      const float tc = tanhf(cn);
      const float dc = dh * og * (1.f - tc * tc) + dc_state;
      d4[3] = dh * tc * og * (1.f - og);
      d4[0] = dc * gg * ig * (1.f - ig);
      d4[1] = dc * cp * fg * (1.f - fg);
      d4[2] = dc * ig * (1.f - gg * gg);
      dc_state = dc * fg;
    } else {
      dc_state = 0.f;
    }
#pragma unroll
    for (int gate = 0; gate < 4; ++gate) {
      const float partner = __shfl_down_sync(0xffffffffu, d4[gate], 1);   // unit ul + 1 of the same batch row
      if ((ul & 1) == 0) {
        uint32_t hi, lo;
        pack_split(d4[gate], partner, hi, lo);
        dgcur[xw[gate]] = hi;
        if (kSplit == 2) dgcur[dg_part * 4 + xw[gate]] = lo;   // lo block: dg_part uint4 = 4 words each further on
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    dir_barrier_arrive(counters + dir);
    if (mine) {
      float* go = dgates_all + (((size_t)dir * T + t) * B + b) * J;
      go[u] = d4[0]; go[H + u] = d4[1]; go[2 * H + u] = d4[2]; go[3 * H + u] = d4[3];
    }
    dir_barrier_wait(counters + dir, d.G, phase);
  }
}

static int lstm_plan(int H, LstmDims* d) {
  // U <= 8 units per CTA, at most 74 CTAs per direction (both directions co-resident on 148 SMs)
  int U = (H + 73) / 74;
  if (U < 1) U = 1;
  if (U > 8) return RADTTS_ERR_UNSUPPORTED;
  d->U = U;
  d->G = (H + U - 1) / U;
  return 0;
}

}  // namespace rb

using namespace rb;

extern "C" size_t radtts_lstm_workspace_bytes(int B, int H) {
  // hbuf [2][2][Hp][32] + dgbuf [2][2][4 Hp][32] floats + counters (exchange buffers are padded to 32 batch lanes and to a
  // multiple of 16 units; the bf16 / split-bf16 fragment-ordered exchanges of the tensor-core kernels fit in the same space)
  (void)B;
  const size_t Hp = (size_t)round_up(H, 16);
  return ((size_t)4 * 32 * Hp + (size_t)16 * 32 * Hp) * sizeof(float) + 256;
}

// precision: RADTTS_PREC_FP32 (exact FMA recurrence), RADTTS_PREC_BF16 (single bf16 operands), RADTTS_PREC_BF16X2 (hi + lo
// bf16 operands, 16 mantissa bits: for fp32 LSTMs wherever cuDNN would be allowed TF32)
static int lstm_split(int precision) { return precision == RADTTS_PREC_BF16X2 ? 2 : 1; }
static bool lstm_mma_ok(int H, int precision) {
  return (precision == RADTTS_PREC_BF16 || precision == RADTTS_PREC_BF16X2) && H % 8 == 0 && H >= 32 && H / 8 <= 74;
}

extern "C" int radtts_lstm_forward(const float* gx, const float* whh, const int* lens, int T, int B, int H,
                                   float* h_all, float* gates_save, float* c_save, void* ws, size_t ws_bytes,
                                   int precision, void* stream) {
  if (!gx || !whh || !lens || !h_all || !ws || T <= 0 || B <= 0 || H <= 0) return RADTTS_ERR_INVALID_ARG;
  if (precision != RADTTS_PREC_FP32 && precision != RADTTS_PREC_BF16 && precision != RADTTS_PREC_BF16X2)
    return RADTTS_ERR_INVALID_ARG;
  if (B > kLstmMaxB) return RADTTS_ERR_UNSUPPORTED;
  if (ws_bytes < radtts_lstm_workspace_bytes(B, H)) return RADTTS_ERR_WORKSPACE;
  LstmDims d{T, B, H, 0, 0};
  RB_TRY(lstm_plan(H, &d));
  cudaStream_t st = (cudaStream_t)stream;
  float* hbuf = reinterpret_cast<float*>(ws);
  unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(ws) + (size_t)20 * 32 * round_up(H, 16) * sizeof(float));
  RB_CUDA(cudaMemsetAsync(hbuf, 0, (size_t)4 * 32 * round_up(H, 16) * sizeof(float), st));
  RB_CUDA(cudaMemsetAsync(counters, 0, 64, st));
  if (lstm_mma_ok(H, precision)) {
    // first choice: one thread-block cluster per (direction, 16-utterance slice), exchange through DSMEM (lstm_cluster.cuh)
    if (lstm_cluster_forward(gx, whh, lens, T, B, H, h_all, gates_save, c_save, lstm_split(precision), st) == 0) return 0;
    d.U = 8;
    d.G = H / 8;
    const int KS = (H + 15) / 16;
    const int split = lstm_split(precision);
    const size_t smem = (size_t)split * ((size_t)2 * KS * 32 * sizeof(uint4) + (size_t)KS * 4 * 32 * sizeof(uint2));
    if (smem <= (size_t)kSmemBudget) {
      void* fn = split == 2 ? (void*)lstm_fwd_mma_kernel<2> : (void*)lstm_fwd_mma_kernel<1>;
      static size_t configured_mma[3] = {0, 0, 0};
      if (smem > configured_mma[split]) {
        RB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_mma[split] = smem;
      }
      uint2* hb = reinterpret_cast<uint2*>(hbuf);
      void* args[] = {(void*)&gx, (void*)&whh, (void*)&lens, (void*)&d, (void*)&h_all, (void*)&gates_save,
                      (void*)&c_save, (void*)&hb, (void*)&counters};
      RB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(2 * d.G), dim3(kLstmMmaFwdThreads), args, smem, st));
      return after_launch();
    }
    RB_TRY(lstm_plan(H, &d));   // does not fit: fall through to the FMA kernel
  }
  const int R = 4 * d.U;
  const size_t smem = ((size_t)H * R + (size_t)H * kLstmLd + (size_t)16 * R * kLstmMaxB) * sizeof(float);
  if (smem > (size_t)kSmemBudget) return RADTTS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    RB_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  void* args[] = {(void*)&gx, (void*)&whh, (void*)&lens, (void*)&d, (void*)&h_all, (void*)&gates_save, (void*)&c_save,
                  (void*)&hbuf, (void*)&counters};
  RB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_fwd_kernel, dim3(2 * d.G), dim3(kLstmThreads), args, smem, st));
  return after_launch();
}

extern "C" int radtts_lstm_backward(const float* dh_all, const float* whh, const int* lens, const float* gates_save,
                                    const float* c_save, int T, int B, int H, float* dgates_all, void* ws,
                                    size_t ws_bytes, int precision, void* stream) {
  if (!dh_all || !whh || !lens || !gates_save || !c_save || !dgates_all || !ws || T <= 0 || B <= 0 || H <= 0)
    return RADTTS_ERR_INVALID_ARG;
  if (precision != RADTTS_PREC_FP32 && precision != RADTTS_PREC_BF16 && precision != RADTTS_PREC_BF16X2)
    return RADTTS_ERR_INVALID_ARG;
  if (B > kLstmMaxB) return RADTTS_ERR_UNSUPPORTED;
  if (ws_bytes < radtts_lstm_workspace_bytes(B, H)) return RADTTS_ERR_WORKSPACE;
  LstmDims d{T, B, H, 0, 0};
  RB_TRY(lstm_plan(H, &d));
  cudaStream_t st = (cudaStream_t)stream;
  float* dgbuf = reinterpret_cast<float*>(ws) + (size_t)4 * 32 * round_up(H, 16);
  unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(ws) + (size_t)20 * 32 * round_up(H, 16) * sizeof(float));
  RB_CUDA(cudaMemsetAsync(dgbuf, 0, (size_t)16 * 32 * round_up(H, 16) * sizeof(float), st));
  RB_CUDA(cudaMemsetAsync(counters, 0, 64, st));
  if (lstm_mma_ok(H, precision)) {
    if (lstm_cluster_backward(dh_all, whh, lens, gates_save, c_save, T, B, H, dgates_all, lstm_split(precision), st) == 0)
      return 0;
    d.U = 8;
    d.G = H / 8;
    const int KS = 4 * H / 16;
    const int split = lstm_split(precision);
    const size_t smem = (size_t)split * ((size_t)KS * 32 * sizeof(uint2) + (size_t)KS * 2 * 32 * sizeof(uint4)) +
                        (size_t)8 * 32 * 8 * sizeof(float);
    if (smem <= (size_t)kSmemBudget) {
      void* fn = split == 2 ? (void*)lstm_bwd_mma_kernel<2> : (void*)lstm_bwd_mma_kernel<1>;
      static size_t configured_mma[3] = {0, 0, 0};
      if (smem > configured_mma[split]) {
        RB_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured_mma[split] = smem;
      }
      uint4* db = reinterpret_cast<uint4*>(dgbuf);
      void* args[] = {(void*)&dh_all, (void*)&whh, (void*)&lens, (void*)&gates_save, (void*)&c_save, (void*)&d,
                      (void*)&dgates_all, (void*)&db, (void*)&counters};
      RB_CUDA(cudaLaunchCooperativeKernel(fn, dim3(2 * d.G), dim3(kLstmMmaBwdThreads), args, smem, st));
      return after_launch();
    }
    RB_TRY(lstm_plan(H, &d));   // the split operands do not fit (H > ~400): fall through to the FMA kernel
  }
  if (H % 2) return RADTTS_ERR_UNSUPPORTED;
  const size_t smem = ((size_t)4 * H * 8 + (size_t)2 * (H / 2) * kLstmLd + (size_t)32 * 8 * kLstmMaxB) * sizeof(float);
  if (smem > (size_t)kSmemBudget) return RADTTS_ERR_UNSUPPORTED;
  static size_t configured = 0;
  if (smem > configured) {
    RB_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  void* args[] = {(void*)&dh_all, (void*)&whh, (void*)&lens, (void*)&gates_save, (void*)&c_save, (void*)&d,
                  (void*)&dgates_all, (void*)&dgbuf, (void*)&counters};
  RB_CUDA(cudaLaunchCooperativeKernel((void*)lstm_bwd_kernel, dim3(2 * d.G), dim3(kLstmThreads), args, smem, st));
  return after_launch();
}

extern "C" int radtts_lstm_debug_timeline(unsigned long long* out16_host) {
  if (!out16_host) return RADTTS_ERR_INVALID_ARG;
  RB_CUDA(cudaDeviceSynchronize());
  RB_CUDA(cudaMemcpyFromSymbol(out16_host, g_cl_dbg, 16 * sizeof(unsigned long long)));
  return 0;
}
