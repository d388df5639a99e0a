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
#include "ptx.cuh"

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
  // hbuf [2][2][H][32] + dgbuf [2][2][4H][32] + counters (exchange buffers are padded to 32 batch lanes)
  (void)B;
  return ((size_t)4 * 32 * H + (size_t)16 * 32 * H) * sizeof(float) + 256;
}

extern "C" int radtts_lstm_forward(const float* gx, const float* whh, const int* lens, int T, int B, int H,
                                   float* h_all, float* gates_save, float* c_save, void* ws, size_t ws_bytes,
                                   void* stream) {
  if (!gx || !whh || !lens || !h_all || !ws || T <= 0 || B <= 0 || H <= 0) return RADTTS_ERR_INVALID_ARG;
  if (B > kLstmMaxB) return RADTTS_ERR_UNSUPPORTED;
  if (ws_bytes < radtts_lstm_workspace_bytes(B, H)) return RADTTS_ERR_WORKSPACE;
  LstmDims d{T, B, H, 0, 0};
  RB_TRY(lstm_plan(H, &d));
  cudaStream_t st = (cudaStream_t)stream;
  float* hbuf = reinterpret_cast<float*>(ws);
  unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(ws) + (size_t)20 * 32 * H * sizeof(float));
  RB_CUDA(cudaMemsetAsync(hbuf, 0, (size_t)4 * 32 * H * sizeof(float), st));
  RB_CUDA(cudaMemsetAsync(counters, 0, 64, st));
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
                                    size_t ws_bytes, void* stream) {
  if (!dh_all || !whh || !lens || !gates_save || !c_save || !dgates_all || !ws || T <= 0 || B <= 0 || H <= 0)
    return RADTTS_ERR_INVALID_ARG;
  if (B > kLstmMaxB) return RADTTS_ERR_UNSUPPORTED;
  if (ws_bytes < radtts_lstm_workspace_bytes(B, H)) return RADTTS_ERR_WORKSPACE;
  LstmDims d{T, B, H, 0, 0};
  RB_TRY(lstm_plan(H, &d));
  cudaStream_t st = (cudaStream_t)stream;
  float* dgbuf = reinterpret_cast<float*>(ws) + (size_t)4 * 32 * H;
  unsigned* counters = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(ws) + (size_t)20 * 32 * H * sizeof(float));
  RB_CUDA(cudaMemsetAsync(dgbuf, 0, (size_t)16 * 32 * H * sizeof(float), st));
  RB_CUDA(cudaMemsetAsync(counters, 0, 64, st));
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
