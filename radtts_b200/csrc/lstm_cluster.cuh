// Cluster / DSMEM variants of the persistent BiLSTM recurrence (bf16 and split-bf16 tensor-core precisions).
//
// The cooperative kernels in lstm.cu exchange h_{t-1} (forward) / dgates_{t+1} (backward) through L2: per time step a
// grid barrier on a global counter (fence + atomic + poll) and a 33 KB bulk copy back into shared memory -- ~4 us of the
// ~5 us a step takes; the MMAs themselves need ~0.6 us.  Here ONE thread-block cluster (<= 16 CTAs on one GPC) runs a
// (direction, 16-utterance batch slice): every CTA owns U = 8 * UG hidden units with its rows of W_hh resident in shared
// memory in MMA fragment order, and the per-step exchange never leaves the GPC:
//   forward : a CTA's new h slice (U x 16 bf16) goes straight into the B-fragment buffer of every CTA of the cluster with
//             one smem -> remote-smem bulk copy per peer, completing on the receiver's mbarrier (no grid barrier, no
//             fence / atomic / poll, no L2 round trip);
//   backward: W_hh^T dgates is computed from each CTA's OWN gate rows (no exchange before the MMAs) and the partial
//             sums over its peers' units are reduce-scattered the same way (U x 16 block per peer).
// Two ping-pong receive buffers are enough: a CTA cannot send step s + 1 data before it has received every peer's step s
// data, and a peer sends step s data only after it is done reading step s - 1.
// mma.sync.m16n8k16 (the per-step GEMM is 4U x 16 x H per CTA: far below a tcgen05 tile); fp32 accumulate, fp32 state,
// gates and outputs.  kSplit = 2: hi + lo bf16 operands (three MMAs, 16 mantissa bits) for fp32 LSTMs.
#pragma once
#include <cuda_bf16.h>
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

namespace rb {

// phase cycle sums of cluster rank 0 / batch slice 0 / direction 0 of the last cluster launch (thread 0):
// forward  [0] wait for h, [1] MMAs, [2] gates + stage, [3] barrier + send, [4] output stores, [5] steps
// backward [8] wait + reduce, [9] element-wise + dgates, [10] MMAs + stage, [11] send, [13] steps
// Compiled in only with -DRB_LSTM_TIMELINE=1: the per-step global read-modify-writes of thread 0 delay its warp, and through
// the CTA barrier and the cluster exchange every step of every CTA (radtts_lstm_debug_timeline then returns zeros).
#ifndef RB_LSTM_TIMELINE
#define RB_LSTM_TIMELINE 0
#endif
__device__ unsigned long long g_cl_dbg[16];

struct ClDims {
  int T, B, H;
  int UG;    // 8-unit groups per CTA
  int G;     // CTAs per cluster
  int NG;    // H / 8 unit groups in total
  int KS;    // ceil(H / 16) k-steps of the forward product
};

__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cl_map(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// smem (this CTA) -> smem (CTA of the cluster) bulk copy, completing `bytes` on the RECEIVER's mbarrier
__device__ __forceinline__ void cl_bulk_s2s(uint32_t dst_cluster_addr, const void* src_local, uint32_t bytes,
                                            uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster_addr),
               "r"(smem_u32(src_local)), "r"(bytes), "r"(bar_cluster_addr)
               : "memory");
}

__device__ __forceinline__ float cl_sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ uint32_t cl_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void cl_split(float x, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16(x));
  lo = x - hi;
}
__device__ __forceinline__ void cl_pack_split(float a, float b, uint32_t& hi, uint32_t& lo) {
  float ah, al, bh, bl;
  cl_split(a, ah, al);
  cl_split(b, bh, bl);
  hi = cl_pack(ah, bh);
  lo = cl_pack(al, bl);
}
__device__ __forceinline__ void cl_mma(float (&c)[4], const uint4& a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b0), "r"(b1));
}

// ----------------------------------------------------------------------------------------------------------
// forward.  grid (G, ceil(B / 16), 2 directions), cluster (G, 1, 1), 32 * UG threads (warp = 8-unit group).
// smem: ws  [kSplit][UG][2 m-tiles][KS][32] uint4      W_hh rows of this CTA in A-fragment order
//            (m-tile 0: rows 0-7 gate i, 8-15 gate f; m-tile 1: gates g, o; unit = row & 7)
//       hs  [2][kSplit][NGp = 2 KS groups][2 n-tiles][32] uint32   h_{t-1} of ALL units, B-fragment order, one 256-byte
//            block per 8-unit group (k-step ks = group / 2, register b0 / b1 = group & 1): a CTA's slice is contiguous
//       st  [2][kSplit][UG][2][32] uint32               this CTA's new slice, the source of the bulk copies
// ----------------------------------------------------------------------------------------------------------
template <int kSplit>
__global__ void __launch_bounds__(256, 1)
lstm_fwd_cluster_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const int* __restrict__ lens, ClDims d,
                        float* __restrict__ h_all, float* __restrict__ gates_save, float* __restrict__ c_save) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int T = d.T, B = d.B, H = d.H, UG = d.UG, KS = d.KS;
  const int rank = (int)cl_rank();
  const int dir = blockIdx.z, b0 = blockIdx.y * 16;
  const int g_first = rank * UG;                                   // first 8-unit group of this CTA
  const int n_own = max(0, min(UG, d.NG - g_first));               // groups it really owns
  const int NGp = 2 * KS;
  const size_t ws_part = (size_t)UG * 2 * KS * 32;                 // uint4 per operand part
  const size_t hs_part = (size_t)NGp * 2 * 32;                     // uint32 per operand part
  const size_t st_part = (size_t)UG * 2 * 32;
  uint4* ws = reinterpret_cast<uint4*>(smraw);
  uint32_t* hs = reinterpret_cast<uint32_t*>(smraw + (size_t)kSplit * ws_part * sizeof(uint4));
  uint32_t* st = hs + (size_t)2 * kSplit * hs_part;
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, ug = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  // W slice
  const float* W = whh + (size_t)dir * 4 * H * H;
  for (int i = tid; i < UG * 2 * KS * 32; i += blockDim.x) {
    const int ln = i & 31, ks = (i >> 5) % KS, mt = (i / (32 * KS)) & 1, wg = i / (64 * KS);
    const int gg = ln >> 2, qq = ln & 3;
    const int u = (g_first + wg) * 8 + gg;
    const int k = ks * 16 + 2 * qq;
    const bool real = wg < n_own;
    const float* r0 = W + (size_t)((2 * mt) * H + u) * H;
    const float* r1 = W + (size_t)((2 * mt + 1) * H + u) * H;
    auto at = [&](const float* r, int kk) { return (real && kk < H) ? r[kk] : 0.f; };
    uint4 hi, lo;
    cl_pack_split(at(r0, k), at(r0, k + 1), hi.x, lo.x);
    cl_pack_split(at(r1, k), at(r1, k + 1), hi.y, lo.y);
    cl_pack_split(at(r0, k + 8), at(r0, k + 9), hi.z, lo.z);
    cl_pack_split(at(r1, k + 8), at(r1, k + 9), hi.w, lo.w);
    ws[i] = hi;
    if (kSplit == 2) ws[ws_part + i] = lo;
  }
  for (int i = tid; i < (int)(2 * kSplit * hs_part); i += blockDim.x) hs[i] = 0u;   // h_{-1} = 0 (and the K padding)
  const bool active = ug < n_own;
  const int u = (g_first + ug) * 8 + g;              // unit this thread finalises (4 batch columns: nt x e)
  int len_b[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int b = b0 + nt * 8 + 2 * q + e;
      len_b[nt][e] = (active && b < B) ? lens[b] : 0;
    }
  float c_state[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  const uint32_t slice_bytes = (uint32_t)(n_own * 256);
  const uint32_t total_tx = (uint32_t)(d.NG * 256 * kSplit);
  __syncthreads();
  cl_sync();                                          // every CTA's barriers and zero state exist before anyone sends

  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const int cur = s & 1, nxt = cur ^ 1;
    float gxv[4][2][2];
#pragma unroll
    for (int gate = 0; gate < 4; ++gate)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int b = b0 + nt * 8 + 2 * q + e;
          gxv[gate][nt][e] = (active && b < B) ? __ldg(gx + (((size_t)dir * T + t) * B + b) * 4 * H + gate * H + u) : 0.f;
        }
    const bool dbg = RB_LSTM_TIMELINE && (tid == 0 && rank == 0 && blockIdx.y == 0 && blockIdx.z == 0);
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (dbg) c0 = clock64();
    if (s > 0) mbar_wait(&bar[cur], (uint32_t)(((s - 1) >> 1) & 1));
    if (dbg) c1 = clock64();
    float acc[2][2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[a][b][c][r] = 0.f;
    if (active) {
      const uint32_t* hb = hs + (size_t)cur * kSplit * hs_part;
      const uint4* wa = ws + (size_t)ug * 2 * KS * 32;
#pragma unroll 2
      for (int ks = 0; ks < KS; ++ks) {
        const uint4 a0 = wa[(size_t)ks * 32 + lane], a1 = wa[((size_t)KS + ks) * 32 + lane];
        const uint32_t* hk = hb + (size_t)ks * 128 + lane;     // [half][nt][lane]
        const uint32_t b00 = hk[0], b01 = hk[64], b10 = hk[32], b11 = hk[96];   // (nt0: b0, b1), (nt1: b0, b1)
        cl_mma(acc[0][0][ks & 1], a0, b00, b01);
        cl_mma(acc[1][0][ks & 1], a1, b00, b01);
        cl_mma(acc[0][1][ks & 1], a0, b10, b11);
        cl_mma(acc[1][1][ks & 1], a1, b10, b11);
        if (kSplit == 2) {
          const uint32_t* hl = hk + hs_part;
          const uint32_t l00 = hl[0], l01 = hl[64], l10 = hl[32], l11 = hl[96];
          cl_mma(acc[0][0][ks & 1], a0, l00, l01);
          cl_mma(acc[1][0][ks & 1], a1, l00, l01);
          cl_mma(acc[0][1][ks & 1], a0, l10, l11);
          cl_mma(acc[1][1][ks & 1], a1, l10, l11);
          const uint4 w0 = wa[ws_part + (size_t)ks * 32 + lane], w1 = wa[ws_part + ((size_t)KS + ks) * 32 + lane];
          cl_mma(acc[0][0][ks & 1], w0, b00, b01);
          cl_mma(acc[1][0][ks & 1], w1, b00, b01);
          cl_mma(acc[0][1][ks & 1], w0, b10, b11);
          cl_mma(acc[1][1][ks & 1], w1, b10, b11);
        }
      }
    }
    if (dbg) c2 = clock64();
    // C fragment: [0],[1] = (row g, cols 2q, 2q+1), [2],[3] = (row g + 8, ...): m-tile 0 rows = gates i | f, m-tile 1 = g | o
    float hval[2][2], ig[2][2], fg[2][2], gg[2][2], og[2][2], cn[2][2];
    uint32_t* sdst = st + (size_t)cur * kSplit * st_part;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pi = acc[0][nt][0][e] + acc[0][nt][1][e] + gxv[0][nt][e];
        const float pf = acc[0][nt][0][2 + e] + acc[0][nt][1][2 + e] + gxv[1][nt][e];
        const float pg = acc[1][nt][0][e] + acc[1][nt][1][e] + gxv[2][nt][e];
        const float po = acc[1][nt][0][2 + e] + acc[1][nt][1][2 + e] + gxv[3][nt][e];
        const bool on = t < len_b[nt][e];
        ig[nt][e] = cl_sigmoid(pi); fg[nt][e] = cl_sigmoid(pf); gg[nt][e] = tanhf(pg); og[nt][e] = cl_sigmoid(po);
        cn[nt][e] = on ? fg[nt][e] * c_state[nt][e] + ig[nt][e] * gg[nt][e] : 0.f;
        hval[nt][e] = on ? og[nt][e] * tanhf(cn[nt][e]) : 0.f;
        c_state[nt][e] = cn[nt][e];
        // units (g even, g + 1) of one batch column share a 32-bit word of the B fragment: lane (col & 7) * 4 + g / 2
        const float partner = __shfl_down_sync(0xffffffffu, hval[nt][e], 4);
        if (active && (g & 1) == 0) {
          uint32_t hi, lo;
          cl_pack_split(hval[nt][e], partner, hi, lo);
          const int w = (ug * 2 + nt) * 32 + (2 * q + e) * 4 + (g >> 1);
          sdst[w] = hi;
          if (kSplit == 2) sdst[st_part + w] = lo;
        }
      }
    if (dbg) c3 = clock64();
    fence_proxy_async();
    __syncthreads();
    if (s + 1 < T && tid < 32) {
      if (lane == 0) mbar_arrive_expect_tx(&bar[nxt], total_tx);     // armed before my own slice leaves: no peer can
      __syncwarp();                                                  // send step s + 1 data without having received it
      if (lane < d.G && n_own > 0) {
        const uint32_t rbar = cl_map(smem_u32(&bar[nxt]), (uint32_t)lane);
#pragma unroll
        for (int part = 0; part < kSplit; ++part) {
          uint32_t* dst_local = hs + ((size_t)nxt * kSplit + part) * hs_part + (size_t)g_first * 64;
          cl_bulk_s2s(cl_map(smem_u32(dst_local), (uint32_t)lane), sdst + part * st_part, slice_bytes, rbar);
        }
      }
    }
    if (active) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int b = b0 + nt * 8 + 2 * q + e;
          if (b < B) {
            if (gates_save) {
              float* gs = gates_save + (((size_t)dir * T + t) * B + b) * 4 * H;
              gs[u] = ig[nt][e]; gs[H + u] = fg[nt][e]; gs[2 * H + u] = gg[nt][e]; gs[3 * H + u] = og[nt][e];
              c_save[(((size_t)dir * T + t) * B + b) * H + u] = cn[nt][e];
            }
            h_all[((size_t)t * B + b) * 2 * H + dir * H + u] = hval[nt][e];
          }
        }
    }
    if (dbg) {
      c4 = clock64();
      if (s == 0) for (int i = 0; i < 6; ++i) g_cl_dbg[i] = 0;
      g_cl_dbg[0] += c1 - c0; g_cl_dbg[1] += c2 - c1; g_cl_dbg[2] += c3 - c2; g_cl_dbg[4] += c4 - c3; g_cl_dbg[5] += 1;
    }
  }
  cl_sync();     // nobody leaves while a peer's copies into / out of its shared memory may still be in flight
}

// ----------------------------------------------------------------------------------------------------------
// forward, register-resident weights.  The smem version above is bound by shared-memory bandwidth: the CTA's whole W_hh
// slice (169 KB at H = 520) crosses the LDS port every time step (measured 3200 of 7600 cycles per step).  Here every
// 8-unit group is run by TWO warps, each holding the A fragments of one half of K in REGISTERS for the whole sequence
// (2 m-tiles x KSH k-steps x 4 registers), so only the h_{t-1} fragments are read from shared memory; the two halves
// swap partial sums through smem and each finalises one 8-utterance n-tile (two (unit, utterance) pairs per thread
// instead of four).  Gate non-linearities use the MUFU-based exp (abs. error ~1e-6, far inside the bf16 / split-bf16
// contract).  KSH = k-steps per half is a compile-time constant so that the fragment arrays stay in registers.
// smem: hs [2][kSplit][NGp = 4 KSH groups][2][32] uint32, st [2][kSplit][UG][2][32] uint32, red [UG][2][32][8] float.
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float cl_sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float cl_tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

template <int kSplit, int KSH>
__global__ void __launch_bounds__(KSH > 8 ? 320 : 128, 1)   // UG <= 5 (H <= 544) resp. UG <= 2 (H <= 256): 64 threads per group
lstm_fwd_cluster_reg_kernel(const float* __restrict__ gx, const float* __restrict__ whh, const int* __restrict__ lens, ClDims d,
                            float* __restrict__ h_all, float* __restrict__ gates_save, float* __restrict__ c_save) {
  extern __shared__ __align__(128) uint8_t smraw[];
  const int T = d.T, B = d.B, H = d.H, UG = d.UG;
  const int rank = (int)cl_rank();
  const int dir = blockIdx.z, b0 = blockIdx.y * 16;
  const int g_first = rank * UG;
  const int n_own = max(0, min(UG, d.NG - g_first));
  constexpr int NGp = 4 * KSH;                                     // groups covered by 2 KSH k-steps (>= NG, zero padded)
  constexpr size_t hs_part = (size_t)NGp * 2 * 32;
  const size_t st_part = (size_t)UG * 2 * 32;
  uint32_t* hs = reinterpret_cast<uint32_t*>(smraw);
  uint32_t* st = hs + (size_t)2 * kSplit * hs_part;
  float* red = reinterpret_cast<float*>(st + (size_t)2 * kSplit * st_part);      // [UG][2 dest halves][32][8]
  // Global traffic of a step goes through shared memory (measured before: 8 scalar loads + 12 scalar stores per thread
  // and step, 4-byte pieces of 6 arrays with 8 KB strides -- issuing them took 3.4 us of the 5.1 us step at H = 520):
  //   gxs [2][16 batch rows][4 gates][U] (+4 floats per row: bank spread)  next step's input-projection slice, fetched
  //        one step ahead with 16-byte cp.async, U * 4-byte runs;
  //   outs [16][6][U] (+4)   i, f, g, o, c, h of this step, written out by all threads as 16-byte stores of whole runs.
  const int U = 8 * UG;
  const int gx_ld = 4 * U + 4, out_ld = 6 * U + 4;
  float* gxs = red + (size_t)UG * 2 * 32 * 8;
  float* outs = gxs + (size_t)2 * 16 * gx_ld;
  const int U_own = 8 * n_own, u0 = g_first * 8;
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ug = warp >> 1, half = warp & 1;
  const int g = lane >> 2, q = lane & 3;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  const bool active = ug < n_own;
  const int u = (g_first + ug) * 8 + g;
  // A fragments of this warp's K half, straight from global memory into registers
  uint4 areg[kSplit][2][KSH];
  {
    const float* W = whh + (size_t)dir * 4 * H * H;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int ks = 0; ks < KSH; ++ks) {
        const int k = (half * KSH + ks) * 16 + 2 * q;
        const float* r0 = W + (size_t)((2 * mt) * H + u) * H;
        const float* r1 = W + (size_t)((2 * mt + 1) * H + u) * H;
        auto at = [&](const float* r, int kk) { return (active && kk < H) ? __ldg(r + kk) : 0.f; };
        uint4 hi, lo;
        cl_pack_split(at(r0, k), at(r0, k + 1), hi.x, lo.x);
        cl_pack_split(at(r1, k), at(r1, k + 1), hi.y, lo.y);
        cl_pack_split(at(r0, k + 8), at(r0, k + 9), hi.z, lo.z);
        cl_pack_split(at(r1, k + 8), at(r1, k + 9), hi.w, lo.w);
        areg[0][mt][ks] = hi;
        if (kSplit == 2) areg[kSplit - 1][mt][ks] = lo;
      }
  }
  for (int i = tid; i < (int)(2 * kSplit * hs_part); i += blockDim.x) hs[i] = 0u;
  // this thread finalises n-tile `half`: utterances b0 + half * 8 + 2q + e
  int len_b[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int b = b0 + half * 8 + 2 * q + e;
    len_b[e] = (active && b < B) ? lens[b] : 0;
  }
  float c_state[2] = {0.f, 0.f};
  const uint32_t slice_bytes = (uint32_t)(n_own * 256);
  const uint32_t total_tx = (uint32_t)(d.NG * 256 * kSplit);
  __syncthreads();
  cl_sync();

  // gx slice of time step tt -> gxs[buf]: (row r, gate, 4 units) per 16-byte copy; rows beyond B are never read.
  // The (row, gate / array, vector) decomposition of a thread's copies is the same every step: worked out once, here
  // (run-time divisions inside the time loop cost more than the copies).
  const int v_per_run = U_own / 4, n_vec = 16 * 4 * v_per_run;
  constexpr int kMaxIn = 3, kMaxOut = 4;          // 16 * 4 * 10 / 320 = 2 resp. 16 * 6 * 10 / 320 = 3 at H = 520
  int in_src[kMaxIn], in_dst[kMaxIn];             // element offsets: global (without the time-step term), gxs (buffer 0)
#pragma unroll
  for (int j = 0; j < kMaxIn; ++j) {
    const int i = tid + j * (int)blockDim.x;
    in_src[j] = -1; in_dst[j] = 0;
    if (i < n_vec) {
      const int v = i % v_per_run, gate = (i / v_per_run) & 3, r = i / (4 * v_per_run);
      if (b0 + r < B) {
        in_src[j] = ((b0 + r) * 4 + gate) * H + u0 + 4 * v;
        in_dst[j] = r * gx_ld + gate * U + 4 * v;
      }
    }
  }
  const int n_arr = gates_save ? 6 : 1;
  int out_src[kMaxOut], out_dst[kMaxOut], out_k[kMaxOut];
#pragma unroll
  for (int j = 0; j < kMaxOut; ++j) {
    const int i = tid + j * (int)blockDim.x;
    out_src[j] = -1; out_dst[j] = 0; out_k[j] = 0;
    if (i < 16 * n_arr * v_per_run) {
      const int v = i % v_per_run, k0 = (i / v_per_run) % n_arr, r = i / (n_arr * v_per_run);
      const int k = gates_save ? k0 : 5, bb = b0 + r;
      if (bb < B) {
        out_src[j] = r * out_ld + k * U + 4 * v;
        out_k[j] = k;
        out_dst[j] = k < 4 ? (bb * 4 + k) * H + u0 + 4 * v : (k == 4 ? bb * H + u0 + 4 * v : (bb * 2 + dir) * H + u0 + 4 * v);
      }
    }
  }
  // (64 UG threads: 128 UG input and 192 UG output vectors per step always fit in kMaxIn / kMaxOut rounds)
  auto prefetch_gx = [&](int tt, int buf) {
    const float* base = gx + ((size_t)dir * T + tt) * B * 4 * H;
    const uint32_t sbase = smem_u32(gxs + (size_t)buf * 16 * gx_ld);
#pragma unroll
    for (int j = 0; j < kMaxIn; ++j)
      if (in_src[j] >= 0)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sbase + 4u * (uint32_t)in_dst[j]), "l"(base + in_src[j]) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch_gx(dir == 0 ? 0 : T - 1, 0);
#if RB_LSTM_TIMELINE
  long long dacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c_prev = 0;
#endif
  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const int cur = s & 1, nxt = cur ^ 1;
    long long c_top = 0;
    if (RB_LSTM_TIMELINE) c_top = clock64();
    // next step's slice goes into the other buffer (its readers finished before the barriers of step s - 1); this step's
    // slice (committed one step ago) must have landed before the partial-sum barrier below publishes it to every thread
    if (s + 1 < T) prefetch_gx(dir == 0 ? s + 1 : T - 2 - s, nxt);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    const bool dbg = RB_LSTM_TIMELINE && (tid == 0 && rank == 0 && blockIdx.y == 0 && blockIdx.z == 0);
    long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
    if (dbg) c0 = clock64();
    if (s > 0) mbar_wait(&bar[cur], (uint32_t)(((s - 1) >> 1) & 1));
    if (dbg) c1 = clock64();
    float acc[2][2][2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[a][b][c][r] = 0.f;
    {
      const uint32_t* hb = hs + (size_t)cur * kSplit * hs_part + (size_t)half * KSH * 128 + lane;
#pragma unroll
      for (int ks = 0; ks < KSH; ++ks) {
        const uint32_t* hk = hb + (size_t)ks * 128;     // [half-of-k-step][nt][lane]
        const uint32_t b00 = hk[0], b01 = hk[64], b10 = hk[32], b11 = hk[96];
        cl_mma(acc[0][0][ks & 1], areg[0][0][ks], b00, b01);
        cl_mma(acc[1][0][ks & 1], areg[0][1][ks], b00, b01);
        cl_mma(acc[0][1][ks & 1], areg[0][0][ks], b10, b11);
        cl_mma(acc[1][1][ks & 1], areg[0][1][ks], b10, b11);
        if (kSplit == 2) {
          const uint32_t* hl = hk + hs_part;
          const uint32_t l00 = hl[0], l01 = hl[64], l10 = hl[32], l11 = hl[96];
          cl_mma(acc[0][0][ks & 1], areg[0][0][ks], l00, l01);
          cl_mma(acc[1][0][ks & 1], areg[0][1][ks], l00, l01);
          cl_mma(acc[0][1][ks & 1], areg[0][0][ks], l10, l11);
          cl_mma(acc[1][1][ks & 1], areg[0][1][ks], l10, l11);
          cl_mma(acc[0][0][ks & 1], areg[kSplit - 1][0][ks], b00, b01);
          cl_mma(acc[1][0][ks & 1], areg[kSplit - 1][1][ks], b00, b01);
          cl_mma(acc[0][1][ks & 1], areg[kSplit - 1][0][ks], b10, b11);
          cl_mma(acc[1][1][ks & 1], areg[kSplit - 1][1][ks], b10, b11);
        }
      }
    }
    if (dbg) c2 = clock64();
    // swap partial sums: this warp keeps n-tile `half`, hands n-tile `1 - half` to its partner warp
    float mine[2][4];
    {
      float* rdst = red + (((size_t)ug * 2 + (1 - half)) * 32 + lane) * 8;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          // selects, not acc[mt][half]: a run-time index would put the accumulators in local memory
          const float s0 = acc[mt][0][0][r] + acc[mt][0][1][r], s1 = acc[mt][1][0][r] + acc[mt][1][1][r];
          rdst[mt * 4 + r] = half ? s0 : s1;
          mine[mt][r] = half ? s1 : s0;
        }
      }
    }
    __syncthreads();
    float gxv[4][2];
#pragma unroll
    for (int gate = 0; gate < 4; ++gate)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int r = half * 8 + 2 * q + e;
        gxv[gate][e] = (active && b0 + r < B) ? gxs[((size_t)cur * 16 + r) * gx_ld + gate * U + ug * 8 + g] : 0.f;
      }
    {
      const float* rsrc = red + (((size_t)ug * 2 + half) * 32 + lane) * 8;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int r = 0; r < 4; ++r) mine[mt][r] += rsrc[mt * 4 + r];
    }
    float hval[2], ig[2], fg[2], gg[2], og[2], cn[2];
    uint32_t* sdst = st + (size_t)cur * kSplit * st_part;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float pi = mine[0][e] + gxv[0][e];
      const float pf = mine[0][2 + e] + gxv[1][e];
      const float pg = mine[1][e] + gxv[2][e];
      const float po = mine[1][2 + e] + gxv[3][e];
      const bool on = t < len_b[e];
      ig[e] = cl_sigmoid_fast(pi); fg[e] = cl_sigmoid_fast(pf); gg[e] = cl_tanh_fast(pg); og[e] = cl_sigmoid_fast(po);
      cn[e] = on ? fg[e] * c_state[e] + ig[e] * gg[e] : 0.f;
      hval[e] = on ? og[e] * cl_tanh_fast(cn[e]) : 0.f;
      c_state[e] = cn[e];
      if (active) {
        float* o = outs + (size_t)(half * 8 + 2 * q + e) * out_ld + ug * 8 + g;
        o[0] = ig[e]; o[U] = fg[e]; o[2 * U] = gg[e]; o[3 * U] = og[e]; o[4 * U] = cn[e]; o[5 * U] = hval[e];
      }
      const float partner = __shfl_down_sync(0xffffffffu, hval[e], 4);
      if (active && (g & 1) == 0) {
        uint32_t hi, lo;
        cl_pack_split(hval[e], partner, hi, lo);
        const int w = (ug * 2 + half) * 32 + (2 * q + e) * 4 + (g >> 1);
        sdst[w] = hi;
        if (kSplit == 2) sdst[st_part + w] = lo;
      }
    }
    if (dbg) c3 = clock64();
    fence_proxy_async();
    __syncthreads();
    if (s + 1 < T && tid < 32) {
      if (lane == 0) mbar_arrive_expect_tx(&bar[nxt], total_tx);
      __syncwarp();
      if (lane < d.G && n_own > 0) {
        const uint32_t rbar = cl_map(smem_u32(&bar[nxt]), (uint32_t)lane);
#pragma unroll
        for (int part = 0; part < kSplit; ++part) {
          uint32_t* dst_local = hs + ((size_t)nxt * kSplit + part) * hs_part + (size_t)g_first * 64;
          cl_bulk_s2s(cl_map(smem_u32(dst_local), (uint32_t)lane), sdst + part * st_part, slice_bytes, rbar);
        }
      }
    }
    {
      // outs was completed before the barrier above: whole U-float runs as 16-byte stores, all threads
      float* gbase = gates_save ? gates_save + ((size_t)dir * T + t) * B * 4 * H : nullptr;
      float* cbase = c_save ? c_save + ((size_t)dir * T + t) * B * H : nullptr;
      float* hbase = h_all + (size_t)t * B * 2 * H;
#pragma unroll
      for (int j = 0; j < kMaxOut; ++j)
        if (out_src[j] >= 0) {
          const float4 val = *reinterpret_cast<const float4*>(outs + out_src[j]);
          float* dst = (out_k[j] < 4 ? gbase : (out_k[j] == 4 ? cbase : hbase)) + out_dst[j];
          *reinterpret_cast<float4*>(dst) = val;
        }
    }
#if RB_LSTM_TIMELINE
    if (dbg) {
      c4 = clock64();
      dacc[0] += c1 - c0; dacc[1] += c2 - c1; dacc[2] += c3 - c2; dacc[4] += c4 - c3; dacc[5] += 1;
      dacc[3] += c0 - c_top;                      // prefetch issue
      if (s > 0) dacc[6] += c_top - c_prev;       // loop back edge
      c_prev = c4;
    }
#endif
  }
#if RB_LSTM_TIMELINE
  if (tid == 0 && rank == 0 && blockIdx.y == 0 && blockIdx.z == 0)
    for (int i = 0; i < 8; ++i) g_cl_dbg[i] = (unsigned long long)dacc[i];
#endif
  cl_sync();
}

// ----------------------------------------------------------------------------------------------------------
// backward.  Same grid / cluster; 256 threads.  Per step, for the CTA's own units u and the slice's 16 utterances:
//   dh = dh_out[t] + sum over peers of their partial (W_hh^T dgates_{t+1})[u]      (received during the previous step)
//   dgates_t = f(dh, dc, saved gates)  -> dgates_all (global, fp32) and, as bf16, the B operand of this step's product
//   partial[m, n] = sum_{k in own gate rows} W_hh[k, m] dgates_t[k, n]   for ALL units m  -> reduce-scattered to the owners
// smem: wt  [kSplit][MT = KS][KB = UG / 2 .. 4 U / 16][32] uint4   W_hh^T columns of this CTA's gate rows, A-fragment order
//            (A[m][k]: m = unit of the product, k = gate * U + local unit)
//       dgs [kSplit][KB][2 n-tiles][32] uint2                      own dgates, B-fragment order
//       stg [MT * 16][16] PT                                       partial sums (all units), unit-major: a peer's block is contiguous
//       rcv [2][G][U * 16] PT                                      partial sums received from every CTA of the cluster
// PT = float when kSplit == 2 (keeps the 16-mantissa-bit contract), bf16 otherwise.
// ----------------------------------------------------------------------------------------------------------
template <int kSplit>
struct ClPartial { using type = __nv_bfloat16; };
template <>
struct ClPartial<2> { using type = float; };

constexpr int kClBwdThreads = 256;
constexpr int kClMaxPairs = 4;     // (unit, batch) pairs finalised per thread: 8 UG * 16 / 256
constexpr int kClMaxMtPerWarp = 5; // m-tiles of the partial product per warp: ceil(H / 16) <= 40

template <int kSplit>
__global__ void __launch_bounds__(kClBwdThreads, 1)
lstm_bwd_cluster_kernel(const float* __restrict__ dh_all, const float* __restrict__ whh, const int* __restrict__ lens,
                        const float* __restrict__ gates_save, const float* __restrict__ c_save, ClDims d,
                        float* __restrict__ dgates_all) {
  using PT = typename ClPartial<kSplit>::type;
  extern __shared__ __align__(128) uint8_t smraw[];
  const int T = d.T, B = d.B, H = d.H, UG = d.UG, MT = d.KS, G = d.G;
  const int U = 8 * UG, KB = U / 4;                    // K = 4 U own gate rows = KB k-steps
  const int J = 4 * H;
  const int rank = (int)cl_rank();
  const int dir = blockIdx.z, b0 = blockIdx.y * 16;
  const int g_first = rank * UG;
  const int n_own = max(0, min(UG, d.NG - g_first));
  const int U_own = 8 * n_own, u0 = g_first * 8;
  const size_t wt_part = (size_t)MT * KB * 32, dg_part = (size_t)KB * 2 * 32;
  uint4* wt = reinterpret_cast<uint4*>(smraw);
  uint2* dgs = reinterpret_cast<uint2*>(smraw + (size_t)kSplit * wt_part * sizeof(uint4));
  PT* stg = reinterpret_cast<PT*>(reinterpret_cast<uint8_t*>(dgs) + (size_t)kSplit * dg_part * sizeof(uint2));
  PT* rcv = stg + (size_t)MT * 16 * 16;
  const size_t rcv_buf = (size_t)G * U * 16;
  __shared__ __align__(8) uint64_t bar[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  const float* W = whh + (size_t)dir * 4 * H * H;
  // A[m][k] = W[gate(k) * H + u0 + ul(k)][m],  k = gate * U + ul
  for (int i = tid; i < MT * KB * 32; i += kClBwdThreads) {
    const int ln = i & 31, kb = (i >> 5) % KB, mt = i / (32 * KB);
    const int gg = ln >> 2, qq = ln & 3;
    auto at = [&](int m, int k) {
      const int gate = k / U, ul = k % U;
      return (m < H && ul < U_own) ? W[(size_t)(gate * H + u0 + ul) * H + m] : 0.f;
    };
    const int m0 = mt * 16 + gg, k0 = kb * 16 + 2 * qq;
    uint4 hi, lo;
    cl_pack_split(at(m0, k0), at(m0, k0 + 1), hi.x, lo.x);
    cl_pack_split(at(m0 + 8, k0), at(m0 + 8, k0 + 1), hi.y, lo.y);
    cl_pack_split(at(m0, k0 + 8), at(m0, k0 + 9), hi.z, lo.z);
    cl_pack_split(at(m0 + 8, k0 + 8), at(m0 + 8, k0 + 9), hi.w, lo.w);
    wt[i] = hi;
    if (kSplit == 2) wt[wt_part + i] = lo;
  }
  for (int i = tid; i < (int)(kSplit * dg_part * 2); i += kClBwdThreads) reinterpret_cast<uint32_t*>(dgs)[i] = 0u;
  // element-wise ownership: pair p = tid + j * 256 -> (n = p / U_own, ul = p % U_own): consecutive threads, consecutive units
  const int n_pairs = U_own * 16;
  int pn[kClMaxPairs], pul[kClMaxPairs], plen[kClMaxPairs];
  float dc_state[kClMaxPairs];
#pragma unroll
  for (int j = 0; j < kClMaxPairs; ++j) {
    const int p = tid + j * kClBwdThreads;
    const bool ok = p < n_pairs;
    pn[j] = ok ? p / U_own : -1;
    pul[j] = ok ? p % U_own : 0;
    const int b = b0 + pn[j];
    plen[j] = (ok && b < B) ? lens[b] : 0;
    if (ok && b >= B) pn[j] = -1;
    dc_state[j] = 0.f;
  }
  const uint32_t blk_bytes_own = (uint32_t)(U_own * 16 * sizeof(PT));         // what every peer sends me
  const uint32_t total_tx = blk_bytes_own * (uint32_t)G;
  __syncthreads();
  cl_sync();

  for (int s = 0; s < T; ++s) {
    const int sf = T - 1 - s;
    const int t = dir == 0 ? sf : T - 1 - sf;
    const int t_prev = dir == 0 ? t - 1 : t + 1;
    const int cur = s & 1, nxt = cur ^ 1;
    float dh[kClMaxPairs], ig[kClMaxPairs], fg[kClMaxPairs], gg[kClMaxPairs], og[kClMaxPairs], cn[kClMaxPairs], cp[kClMaxPairs];
#pragma unroll
    for (int j = 0; j < kClMaxPairs; ++j) {
      dh[j] = ig[j] = fg[j] = gg[j] = og[j] = cn[j] = cp[j] = 0.f;
      if (pn[j] >= 0) {
        const int b = b0 + pn[j], u = u0 + pul[j];
        dh[j] = dh_all[((size_t)t * B + b) * 2 * H + dir * H + u];
        if (t < plen[j]) {
          const float* gs = gates_save + (((size_t)dir * T + t) * B + b) * 4 * H;
          ig[j] = gs[u]; fg[j] = gs[H + u]; gg[j] = gs[2 * H + u]; og[j] = gs[3 * H + u];
          cn[j] = c_save[(((size_t)dir * T + t) * B + b) * H + u];
          const bool has_prev = (t_prev >= 0 && t_prev < T) && (t_prev < plen[j]);
          cp[j] = has_prev ? c_save[(((size_t)dir * T + t_prev) * B + b) * H + u] : 0.f;
        }
      }
    }
    if (s > 0) {
      mbar_wait(&bar[cur], (uint32_t)(((s - 1) >> 1) & 1));
      const PT* rb = rcv + (size_t)cur * rcv_buf;
#pragma unroll
      for (int j = 0; j < kClMaxPairs; ++j)
        if (pn[j] >= 0) {
          float acc = 0.f;
          for (int src = 0; src < G; ++src) acc += (float)rb[((size_t)src * U + pul[j]) * 16 + pn[j]];
          dh[j] += acc;
        }
    }
#pragma unroll
    for (int j = 0; j < kClMaxPairs; ++j) {
      float d4[4] = {0.f, 0.f, 0.f, 0.f};
      if (pn[j] >= 0 && t < plen[j]) {
        const float tc = cl_tanh_fast(cn[j]);
        const float dc = dh[j] * og[j] * (1.f - tc * tc) + dc_state[j];
        d4[3] = dh[j] * tc * og[j] * (1.f - og[j]);
        d4[0] = dc * gg[j] * ig[j] * (1.f - ig[j]);
        d4[1] = dc * cp[j] * fg[j] * (1.f - fg[j]);
        d4[2] = dc * ig[j] * (1.f - gg[j] * gg[j]);
        dc_state[j] = dc * fg[j];
      } else {
        dc_state[j] = 0.f;
      }
      if (pn[j] >= 0) {
        const int b = b0 + pn[j], u = u0 + pul[j];
        float* go = dgates_all + (((size_t)dir * T + t) * B + b) * J;
        go[u] = d4[0]; go[H + u] = d4[1]; go[2 * H + u] = d4[2]; go[3 * H + u] = d4[3];
        // B operand (k = gate * U + ul, n): element index inside the fragment-ordered buffer
#pragma unroll
        for (int gate = 0; gate < 4; ++gate) {
          const int k = gate * U + pul[j], n = pn[j];
          const int kb = k >> 4, kk = k & 15;
          const size_t word = ((((size_t)kb * 2 + (n >> 3)) * 32 + (n & 7) * 4 + ((kk & 7) >> 1)) * 2 + (kk >> 3));
          float hi, lo;
          cl_split(d4[gate], hi, lo);
          reinterpret_cast<__nv_bfloat16*>(dgs)[word * 2 + (kk & 1)] = __float2bfloat16(hi);
          if (kSplit == 2) reinterpret_cast<__nv_bfloat16*>(dgs + dg_part)[word * 2 + (kk & 1)] = __float2bfloat16(lo);
        }
      }
    }
    __syncthreads();
    if (s + 1 < T) {
      // partial[m][n] for all units m: warp w owns m-tiles w, w + 8, ... (accumulators in registers), so the small B
      // operand is fetched once per k-step and every A fragment exactly once per time step
      float acc[kClMaxMtPerWarp][2][4];
#pragma unroll
      for (int i = 0; i < kClMaxMtPerWarp; ++i)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[i][nt][r] = 0.f;
      for (int kb = 0; kb < KB; ++kb) {
        const uint2 bh0 = dgs[((size_t)kb * 2) * 32 + lane], bh1 = dgs[((size_t)kb * 2 + 1) * 32 + lane];
        uint2 bl0 = bh0, bl1 = bh1;
        if (kSplit == 2) { bl0 = dgs[dg_part + ((size_t)kb * 2) * 32 + lane]; bl1 = dgs[dg_part + ((size_t)kb * 2 + 1) * 32 + lane]; }
#pragma unroll
        for (int i = 0; i < kClMaxMtPerWarp; ++i) {
          const int mt = warp + i * (kClBwdThreads / 32);
          if (mt < MT) {
            const uint4 a = wt[((size_t)mt * KB + kb) * 32 + lane];
            cl_mma(acc[i][0], a, bh0.x, bh0.y);
            cl_mma(acc[i][1], a, bh1.x, bh1.y);
            if (kSplit == 2) {
              cl_mma(acc[i][0], a, bl0.x, bl0.y);
              cl_mma(acc[i][1], a, bl1.x, bl1.y);
              const uint4 al = wt[wt_part + ((size_t)mt * KB + kb) * 32 + lane];
              cl_mma(acc[i][0], al, bh0.x, bh0.y);
              cl_mma(acc[i][1], al, bh1.x, bh1.y);
            }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kClMaxMtPerWarp; ++i) {
        const int mt = warp + i * (kClBwdThreads / 32);
        if (mt < MT) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            PT* r0 = stg + ((size_t)(mt * 16 + g) * 16 + nt * 8 + 2 * q);
            PT* r1 = stg + ((size_t)(mt * 16 + g + 8) * 16 + nt * 8 + 2 * q);
            r0[0] = (PT)acc[i][nt][0]; r0[1] = (PT)acc[i][nt][1];
            r1[0] = (PT)acc[i][nt][2]; r1[1] = (PT)acc[i][nt][3];
          }
        }
      }
      fence_proxy_async();
      __syncthreads();
      if (tid < 32) {
        if (lane == 0) mbar_arrive_expect_tx(&bar[nxt], total_tx);
        __syncwarp();
        if (lane < G) {
          // peer `lane` owns units [lane * U, lane * U + U_peer): its block of my partial sums is contiguous in stg
          const int peer_groups = max(0, min(UG, d.NG - lane * UG));
          const uint32_t bytes = (uint32_t)(peer_groups * 8 * 16 * sizeof(PT));
          if (bytes) {
            PT* dst_local = rcv + (size_t)nxt * rcv_buf + (size_t)rank * U * 16;
            cl_bulk_s2s(cl_map(smem_u32(dst_local), (uint32_t)lane), stg + (size_t)lane * U * 16, bytes,
                        cl_map(smem_u32(&bar[nxt]), (uint32_t)lane));
          }
        }
      }
    }
  }
  cl_sync();
}

// ----------------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------------
inline bool cl_plan(int T, int B, int H, ClDims* d) {
  if (H % 8 || H < 32) return false;
  d->T = T; d->B = B; d->H = H;
  d->NG = H / 8;
  d->UG = ceil_div(d->NG, 16);
  if (d->UG > 8 || ceil_div(H, 16) > kClMaxMtPerWarp * (kClBwdThreads / 32)) return false;
  d->G = ceil_div(d->NG, d->UG);
  d->KS = ceil_div(H, 16);
  return true;
}
inline size_t cl_fwd_smem(const ClDims& d, int split) {
  return (size_t)split * ((size_t)d.UG * 2 * d.KS * 32 * 16 + (size_t)2 * (2 * d.KS) * 2 * 32 * 4 + (size_t)2 * d.UG * 2 * 32 * 4);
}
inline size_t cl_bwd_smem(const ClDims& d, int split) {
  const size_t U = 8 * d.UG, KB = U / 4, pt = split == 2 ? 4 : 2;
  return (size_t)split * ((size_t)d.KS * KB * 32 * 16 + KB * 2 * 32 * 8) + (size_t)d.KS * 16 * 16 * pt +
         (size_t)2 * d.G * U * 16 * pt;
}
// RADTTS_LSTM_CLUSTER: 0 = cooperative (L2-exchange) kernels, 1 = cluster kernels for both passes, 2 = forward only (the
// default: measured inside the cfg2 train step, 32 x 400 x 520: forward 30.27 -> 29.26 ms per step; the cluster backward
// still re-reads its 169 KB W_hh^T slice from shared memory every step and is no faster than the cooperative one), 3 =
// backward only
inline bool cl_enabled(bool backward) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("RADTTS_LSTM_CLUSTER");
    mode = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 2;
  }
  return mode == 1 || (mode == 2 && !backward) || (mode == 3 && backward);
}

template <typename... Args>
inline int cl_launch(void (*kernel)(Args...), const ClDims& d, int threads, size_t smem, cudaStream_t st, Args... args) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
      cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return RADTTS_ERR_UNSUPPORTED;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(d.G, ceil_div(d.B, 16), 2);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = d.G;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int clusters = 0;
  if (cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg) != cudaSuccess || clusters < 1) {
    cudaGetLastError();
    return RADTTS_ERR_UNSUPPORTED;      // this cluster shape cannot be scheduled on the device: the caller falls back
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args...);
  if (e != cudaSuccess) { cudaGetLastError(); return RADTTS_ERR_UNSUPPORTED; }
  ++g_launches;
  return 0;
}

// returns 0 when the cluster kernel was launched, RADTTS_ERR_UNSUPPORTED when the caller should use the cooperative path
inline int lstm_cluster_forward(const float* gx, const float* whh, const int* lens, int T, int B, int H, float* h_all,
                                float* gates_save, float* c_save, int split, cudaStream_t st) {
  ClDims d;
  if (!cl_enabled(false) || !cl_plan(T, B, H, &d)) return RADTTS_ERR_UNSUPPORTED;
  {
    // register-resident weights when an instantiation covers this H: KSH k-steps per warp half, 2 KSH >= ceil(H / 16)
    auto reg_smem = [&](int ksh) {
      const size_t U = (size_t)8 * d.UG;
      return (size_t)split * ((size_t)2 * (4 * ksh) * 2 * 32 * 4 + (size_t)2 * d.UG * 2 * 32 * 4) + (size_t)d.UG * 2 * 32 * 8 * 4 +
             ((size_t)2 * 16 * (4 * U + 4) + (size_t)16 * (6 * U + 4)) * 4;
    };
    if (split == 1 && d.KS <= 34 && d.KS > 16)
      return cl_launch(lstm_fwd_cluster_reg_kernel<1, 17>, d, 64 * d.UG, reg_smem(17), st, gx, whh, lens, d, h_all, gates_save, c_save);
    if (split == 1 && d.KS <= 16)
      return cl_launch(lstm_fwd_cluster_reg_kernel<1, 8>, d, 64 * d.UG, reg_smem(8), st, gx, whh, lens, d, h_all, gates_save, c_save);
    if (split == 2 && d.KS <= 16)
      return cl_launch(lstm_fwd_cluster_reg_kernel<2, 8>, d, 64 * d.UG, reg_smem(8), st, gx, whh, lens, d, h_all, gates_save, c_save);
  }
  const size_t smem = cl_fwd_smem(d, split);
  if (smem > (size_t)kSmemBudget) return RADTTS_ERR_UNSUPPORTED;
  if (split == 2) return cl_launch(lstm_fwd_cluster_kernel<2>, d, 32 * d.UG, smem, st, gx, whh, lens, d, h_all, gates_save, c_save);
  return cl_launch(lstm_fwd_cluster_kernel<1>, d, 32 * d.UG, smem, st, gx, whh, lens, d, h_all, gates_save, c_save);
}
inline int lstm_cluster_backward(const float* dh_all, const float* whh, const int* lens, const float* gates_save,
                                 const float* c_save, int T, int B, int H, float* dgates_all, int split, cudaStream_t st) {
  ClDims d;
  if (!cl_enabled(true) || !cl_plan(T, B, H, &d)) return RADTTS_ERR_UNSUPPORTED;
  const size_t smem = cl_bwd_smem(d, split);
  if (smem > (size_t)kSmemBudget) return RADTTS_ERR_UNSUPPORTED;
  if (split == 2)
    return cl_launch(lstm_bwd_cluster_kernel<2>, d, kClBwdThreads, smem, st, dh_all, whh, lens, gates_save, c_save, d, dgates_all);
  return cl_launch(lstm_bwd_cluster_kernel<1>, d, kClBwdThreads, smem, st, dh_all, whh, lens, gates_save, c_save, d, dgates_all);
}

}  // namespace rb
