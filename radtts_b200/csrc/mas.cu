// Kernel 1: batch-parallel Monotonic Alignment Search fused with backtrack and hard-attention scatter.
//
// Replaces RADTTS.binarize_attention (reference radtts.py:320-334) + alignment.mas_width1 (reference
// alignment.py:31-59).  One CTA per utterance:
//   producer warp : streams the utterance's (out_len x T2) cost rows HBM -> smem ring with 1-D bulk async copies
//                   (cp.async.bulk + mbarrier), many KB in flight per SM;
//   W DP warps    : warp w owns 32 NC text columns, NC (= 2 for T2 > 32) per lane (lane l: columns l, l + 32 of the
//                   block), running scores in registers.  Row i needs only row i-1 (columns j-1, j): inside a warp one
//                   rotate-by-one shuffle per column set; across warps the last column's score goes through a smem ring
//                   in which the VALUE IS ITS OWN FLAG (signalling-NaN sentinel, see dp_warp), checked once per group of
//                   four rows, so warp w trails warp w-1 by a few rows (skewed wavefront, no CTA-wide barrier, no
//                   fences in the loop).  The diagonal/straight decision of every cell is ONE BIT: a warp ballot per
//                   row yields the 32 decisions of a column set as one word, already in natural column order;
//   fill warps    : zero the utterance's slab of the dense hard map while the DP runs;
//   then warp 0 backtracks 32 rows at a time with the bit windows held in registers (no dependent smem latency per
//   step), and all threads scatter the ones, frame->token indices and durations.
// Arithmetic is exactly the reference's: one fp32 add per cell, `>=` tie-break toward the diagonal, row 0 restricted
// to column 0, and the unconditional opt[0,0] = 1 (alignment.py:59).
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace rb {

constexpr int kMasMaxDpWarps = 9;    // T2 <= 576
constexpr int kMasFillWarps = 4;
constexpr int kMasMaxStages = 32;
constexpr int kMasEdge = 256;        // slots of each inter-warp edge ring (1 KB, 1 KB aligned)
constexpr uint32_t kMasSentinel = 0x7fa00001u;   // a SIGNALLING NaN: no fp32 add can ever produce it (see dp_warp)

struct MasPlan {
  int W;                // DP warps
  int NC;               // column sets per lane (1..4)
  int Wb;               // decision words per row (NC per DP warp)
  int threads;
  int rows_per_chunk;   // rows per bulk copy (= progress-publication granularity)
  int stages;
  uint32_t stage_bytes;
  uint32_t path_off, dur_off, edge_off, bits_off, ring_off, smem_bytes;
  int bits_in_smem;
  size_t bits_ws_bytes;
};

static int make_plan(int B, int T1, int T2, MasPlan* p) {
  // NC column sets per lane (lane l of warp w owns columns 32 NC w + l + 32 k, k < NC) so that at most ~4 DP warps run,
  // one per scheduler: more warps cost ~8 cycles per row each (issue interference), more columns per lane are almost free
  // (independent dependency chains that interleave)
  // measured (64 x 2000 x T2, cycles per row of the last warp): NC=1: 161 (T2=300, 10 warps); NC=2: 159 (5 warps);
  // NC=3: 190 (4 warps) -- two sets per lane it is, three or four only where the warp limit requires them
  const int NC = T2 <= 32 ? 1 : (T2 <= 64 * kMasMaxDpWarps ? 2 : 4);
  const int W = (T2 + 32 * NC - 1) / (32 * NC);
  if (W > kMasMaxDpWarps || T1 > 65535) return RADTTS_ERR_UNSUPPORTED;
  p->W = W;
  p->NC = NC;
  p->Wb = NC * W;                    // 32-bit decision words per lattice row
  p->threads = (W + 1 + kMasFillWarps) * 32;
  int r = 1;
  uint32_t off = 2 * kMasMaxStages * 8 + 256;            // mbarriers + progress counters
  p->path_off = off;            off += (uint32_t)round_up(T1 * 2, 16);
  p->dur_off = off;             off += (uint32_t)round_up(T2 * 4, 16);
  p->edge_off = off;            off += (uint32_t)((W + 2) * kMasEdge * 4);   // ring W = dump ring of the lanes != 31; + 1 KB so
                                                                             // that the kernel can align the rings to 1 KB
  off = (uint32_t)round_up((int)off, 128);
  p->bits_off = off;
  const uint32_t bits_bytes = (uint32_t)round_up(T1 * p->Wb * 4, 128);
  const uint32_t need_stages = 5;                         // DP warps trail each other by a row or two: a few stages of
                                                          // prefetch depth are enough, the rest goes into longer chunks
  auto stage_bytes_for = [&](int rows) { return (uint32_t)round_up(rows * T2 * 4 + 256, 128); };  // + alignment slack
  const uint32_t budget = (uint32_t)kSmemBudget - 4096u;  // head-room for the prefetch-overshoot row
  if (off + bits_bytes + need_stages * stage_bytes_for(4) <= budget) {
    p->bits_in_smem = 1;
    p->ring_off = off + bits_bytes;
    p->bits_ws_bytes = 0;
  } else {
    p->bits_in_smem = 0;
    p->ring_off = off;
    p->bits_ws_bytes = (size_t)B * T1 * p->Wb * 4;
  }
  // rows per chunk (= per-chunk synchronisation amortised over that many rows): as many as 5 stages allow, <= 16
  for (r = 16; r > 1; --r)
    if (p->ring_off + need_stages * stage_bytes_for(r) <= budget) break;
  p->rows_per_chunk = r;
  p->stage_bytes = stage_bytes_for(r);
  const uint32_t row_pad = (uint32_t)round_up(T2 * 4, 16);   // one row of prefetch overshoot behind the last stage
  int st = (int)((kSmemBudget - p->ring_off - row_pad) / p->stage_bytes);
  if (st > kMasMaxStages) st = kMasMaxStages;
  // a DP warp may lead its right neighbour by at most `stages` chunks (ring back-pressure); the edge ring between
  // them is never checked for overrun, so that lead must stay below its kMasEdge slots
  if (W > 1 && (st + 1) * r > kMasEdge - 8) st = (kMasEdge - 8) / r - 1;
  if (st < 2) return RADTTS_ERR_UNSUPPORTED;
  p->stages = st;
  p->smem_bytes = p->ring_off + (uint32_t)st * p->stage_bytes + row_pad;
  return 0;
}

// Optional pre-pass for probability input: y = logf(x).  (HBM-bound, all SMs.)
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
  size_t head = ((16 - ((uintptr_t)p & 15)) & 15) / 4;
  if (head > n) head = n;
  for (size_t k = tid; k < head; k += nthreads) p[k] = 0.f;
  float4* p4 = reinterpret_cast<float4*>(p + head);
  size_t n4 = (n - head) / 4;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t k = tid; k < n4; k += nthreads) p4[k] = z;
  for (size_t k = head + n4 * 4 + tid; k < n; k += nthreads) p[k] = 0.f;
}

__device__ __forceinline__ int ld_acquire_s32(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
  asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void st_shared_pred(uint32_t addr, uint32_t v, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.shared.b32 [%0], %1; }" ::"r"(addr), "r"(v), "r"(pred)
               : "memory");
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u32x2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_u32x2_pred(uint32_t addr, uint32_t a, uint32_t b, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %3, 0; @q st.shared.v2.b32 [%0], {%1, %2}; }" ::"r"(addr), "r"(a), "r"(b),
               "r"(pred)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_pred(uint64_t* bar, bool pred) {
  asm volatile("{ .reg .pred q; .reg .b64 t; setp.ne.s32 q, %1, 0; @q mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"(
                   smem_u32(bar)),
               "r"((int)pred)
               : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds_f32_pred(uint32_t addr, int pred, float otherwise) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q ld.shared.f32 %0, [%1]; }" : "+f"(otherwise) : "r"(addr), "r"(pred)
               : "memory");
  return otherwise;
}
__device__ __forceinline__ void st_global_pred(uint32_t* p, uint32_t v, int pred) {
  asm volatile("{ .reg .pred q; setp.ne.s32 q, %2, 0; @q st.global.b32 [%0], %1; }" ::"l"(p), "r"(v), "r"(pred) : "memory");
}
// producer-side wait with back-off: a tight try_wait spin slows the DP warps of the same SM by ~20 cycles per row
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(256);
    if (++spins > (1u << 22)) __trap();
  }
}

// phase timeline of CTA 0 (globaltimer ns): [0] start, [1] DP warp 0 done, [2] all DP/fill warps done, [3] backtrack
// done, [4] end, [5] fill warps done.  Read back with radtts_mas_debug_timeline(); costs a handful of stores.
__device__ unsigned long long g_mas_timeline[16];
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One DP warp: columns [32 warp, 32 warp + 32) of the utterance, one per lane, R rows per ring stage.
//
// Hand-off between neighbouring warps is PER ROW and the value is its own flag: lane 31 of warp w stores its running score
// of row r into slot r of ring w; lane 0 of warp w+1 needs exactly that number for row r+1, spins while the slot holds
// the sentinel and writes the sentinel back once it has the value.  The sentinel is a signalling NaN: every score is
// the result of an fp32 add, which can only produce quiet NaNs, so no computed value collides with it (row 0 takes
// `a + 0.0f` for that reason).  A 32-bit shared-memory store is atomic: no fence, no separate progress counter, and
// warp w+1 trails warp w by ~1 row instead of one ring stage.  Ring overrun is impossible by construction: a warp can
// lead its neighbour by at most stages * R rows (the cost ring's back-pressure), which make_plan keeps below the ring.
//
// Instruction diet (W >= 10 warps share 4 issue ports): all addresses are running 32-bit shared-window pointers, the
// decision word is stored by all lanes to the same address, the edge value by all lanes to a lane-dependent address
// (lane 31 -> the real ring, the others -> a dump ring), i.e. no predicates and no divergent single-lane branches
// (two of those cost ~120 cycles per row, tools/mas_micro.cu).
template <bool kBitsInSmem, bool kFirstWarp, int NC>
__device__ __forceinline__ void dp_warp(uint64_t* full, uint64_t* empty, float* edge, uint8_t* ring, uint32_t* bits_s,
                                        uint32_t* bits_g, const MasPlan& plan, size_t slab, int T2, int olen, int nchunks,
                                        int warp, int lane, bool probe) {
  const int W = plan.W, Wb = plan.Wb, R = plan.rows_per_chunk, stages = plan.stages;
  // NC column sets per lane: lane l owns columns c_k = 32 NC warp + 32 k + l, so that each warp ballot is already a word
  // of 32 consecutive columns.  Left neighbour of c_k: lane l-1's c_k, and for lane 0 lane 31's c_{k-1} (k > 0) or the
  // upstream warp's last column (k = 0): ONE rotate-by-one shuffle per column set delivers both cases.
  const int col0 = warp * 32 * NC + lane;
  const float nanv = __int_as_float(0x7fc00000);
  const bool is_l0 = lane == 0;
  const int rot = (lane + 31) & 31;
  const uint32_t row_bytes = (uint32_t)T2 * 4u;
  // Single-lane stores are PREDICATED: a divergent branch per store costs ~60 cycles (tools/mas_micro.cu).
  const uint32_t my_ring = smem_u32(edge + (size_t)warp * kMasEdge);   // written by lane 31 only
  const int l0i = lane == 0, l31i = lane == 31;
  const uint32_t up_base = smem_u32(edge + (size_t)(kFirstWarp ? W : warp - 1) * kMasEdge);
  uint32_t pb = smem_u32(bits_s + NC * warp);                  // decision words of (row, warp), smem variant
  uint32_t* pbg = bits_g + NC * warp;                          // ... global variant
  const uint32_t ring_base = smem_u32(ring);
  uint32_t cofs[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) cofs[k] = (uint32_t)min(col0 + 32 * k, T2 - 1) * 4u;
  uint32_t mis = (uint32_t)((slab * 4) & 127);          // offset of the chunk's first byte inside its stage
  const uint32_t mis_step = ((uint32_t)R * row_bytes) & 127u;
  float v[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) v[k] = -CUDART_INF_F;
  int s = 0;
  uint32_t phase = 0;
  long long tw_wait = 0, tw_rows = 0, tqs = clock64();
  long long n_spin = 0;
  const bool probe0 = kFirstWarp && blockIdx.x == 0;
  auto slot = [&](int q) { return up_base + ((uint32_t)q & (kMasEdge - 1)) * 4u; };
  // one lattice row: compare / select / add exactly as the reference, one ballot per column set, decision words stored
  // by lane 0, the last column's score published by lane 31
  auto step = [&](const float (&a)[NC], float e_up, int row, uint32_t pbits, uint32_t* pbitsg) {
    float t[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) t[k] = __shfl_sync(0xffffffffu, v[k], rot);
    uint32_t w[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const float left = is_l0 ? (k == 0 ? e_up : t[k - 1]) : t[k];   // first warp: e_up = NaN, NaN >= x is false
      const bool d = (left >= v[k]);
      v[k] = __fadd_rn(a[k], d ? left : v[k]);
      w[k] = __ballot_sync(0xffffffffu, d);
    }
    if constexpr (kBitsInSmem) {
      if constexpr (NC == 1) st_shared_pred(pbits, w[0], l0i);
      else if constexpr (NC == 2) sts_u32x2_pred(pbits, w[0], w[1], l0i);
      else {
#pragma unroll
        for (int k = 0; k < NC; ++k) st_shared_pred(pbits + 4u * k, w[k], l0i);
      }
    } else if (is_l0) {
#pragma unroll
      for (int k = 0; k < NC; ++k) pbitsg[k] = w[k];
    }
    st_shared_pred(my_ring + ((uint32_t)row & (kMasEdge - 1)) * 4u, __float_as_uint(v[NC - 1]), l31i);
  };
  float eu[4] = {nanv, nanv, nanv, nanv};   // upstream scores for the current group of 4 rows (first warp: NaN)
  for (int c = 0, r0 = 0; c < nchunks; ++c, r0 += R) {
    const int r1 = min(r0 + R, olen);
    mbar_wait(&full[s], phase);
    const long long tq1 = clock64();
    uint32_t pa = ring_base + (uint32_t)s * plan.stage_bytes + mis;   // row pointer (column offsets added per load)
    float a[NC];
#pragma unroll
    for (int k = 0; k < NC; ++k) a[k] = lds_f32(pa + cofs[k]);
    pa += row_bytes;
    int row = r0;
    if (row == 0) {                      // alignment.py:41-42: row 0 may only sit on token 0 (and carries no decisions)
      v[0] = (col0 == 0) ? __fadd_rn(a[0], 0.f) : -CUDART_INF_F;
      st_shared_pred(my_ring, __float_as_uint(v[NC - 1]), l31i);
#pragma unroll
      for (int k = 0; k < NC; ++k) a[k] = lds_f32(pa + cofs[k]);
      pa += row_bytes;
      pb += (uint32_t)Wb * 4u;
      pbg += Wb;
      row = 1;
    }
    // groups of 4 rows: ONE wait (on the newest upstream score the group needs; lane 31 of the upstream warp stores its
    // scores in row order, so the older three are then in place too), 4 hand-backs, and a spin-free body the compiler
    // can unroll and overlap.
    while (row + 4 <= r1) {
      if constexpr (!kFirstWarp) {
        eu[3] = lds_f32(slot(row + 2));
        unsigned spins = 0;              // bounded like every other wait: a protocol bug must trap, not hang
        while (__any_sync(0xffffffffu, __float_as_uint(eu[3]) == kMasSentinel)) {
          __nanosleep(32);               // a tight spin steals issue slots from the DP warps sharing this scheduler
          eu[3] = lds_f32(slot(row + 2));
          ++n_spin;
          if (++spins > (1u << 22)) __trap();
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) eu[k] = lds_f32(slot(row - 1 + k));
#pragma unroll
        for (int k = 0; k < 4; ++k) st_shared_pred(slot(row - 1 + k), kMasSentinel, l0i);   // hand the slots back
      }
      float an[4][NC];
#pragma unroll
      for (int j = 0; j < 4; ++j)        // costs of the next four rows (overshoots the chunk by one row at most)
#pragma unroll
        for (int k = 0; k < NC; ++k) an[j][k] = lds_f32(pa + (uint32_t)j * row_bytes + cofs[k]);
      pa += 4u * row_bytes;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        step(a, eu[j], row + j, pb + (uint32_t)(j * Wb) * 4u, pbg + (size_t)j * Wb);
#pragma unroll
        for (int k = 0; k < NC; ++k) a[k] = an[j][k];
      }
      pb += 4u * (uint32_t)Wb * 4u;
      pbg += 4 * Wb;
      row += 4;
    }
    for (; row < r1; ++row) {            // chunk tail (< 4 rows)
      float n[NC];
#pragma unroll
      for (int k = 0; k < NC; ++k) n[k] = lds_f32(pa + cofs[k]);   // next row: unconditional (padded ring)
      pa += row_bytes;
      float e = nanv;
      if constexpr (!kFirstWarp) {
        const uint32_t s0 = slot(row - 1);
        e = lds_f32(s0);
        unsigned spins = 0;
        while (__any_sync(0xffffffffu, __float_as_uint(e) == kMasSentinel)) {
          __nanosleep(32);
          e = lds_f32(s0);
          if (++spins > (1u << 22)) __trap();
        }
        st_shared_pred(s0, kMasSentinel, l0i);   // consumed: hand the slot back
      }
      step(a, e, row, pb, pbg);
      pb += (uint32_t)Wb * 4u;
      pbg += Wb;
#pragma unroll
      for (int k = 0; k < NC; ++k) a[k] = n[k];
    }
    const long long tq2 = clock64();
    __syncwarp();
    mbar_arrive_pred(&empty[s], is_l0);
    if (++s == stages) { s = 0; phase ^= 1u; }
    mis = (mis + mis_step) & 127u;
    if (probe || probe0) { tw_wait += tq1 - tqs; tw_rows += tq2 - tq1; }
    tqs = clock64();
  }
  if (probe0 && !probe && lane == 0) g_mas_timeline[13] = tw_rows;
  if (probe && lane == 0) {
    g_mas_timeline[14] = n_spin;
    g_mas_timeline[8] = tw_wait; g_mas_timeline[9] = tw_rows; g_mas_timeline[10] = 0;
    g_mas_timeline[11] = nchunks; g_mas_timeline[12] = R;
  }
}

// kBitsInSmem selects at COMPILE time where the bit lattice lives: a pointer chosen at run time between shared and
// global memory compiles to generic ST.E / LD.E, and a generic store in the row loop cost ~300 cycles per DP row
// (every following LDS waits for the generic address to resolve).
template <bool kBitsInSmem>
__global__ void __launch_bounds__((kMasMaxDpWarps + 1 + kMasFillWarps) * 32, 1)
mas_kernel(const float* __restrict__ logp, const int64_t* __restrict__ in_lens, const int64_t* __restrict__ out_lens,
           int T1, int T2, unsigned long long total_bytes, float* __restrict__ hard, int32_t* __restrict__ f2t,
           int32_t* __restrict__ dur, uint32_t* __restrict__ bits_ws, MasPlan plan) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem);
  uint64_t* empty = full + kMasMaxStages;
  uint16_t* path = reinterpret_cast<uint16_t*>(smem + plan.path_off);
  int* dur_s = reinterpret_cast<int*>(smem + plan.dur_off);
  // [W + 1][kMasEdge] rings on a 1 KB boundary of the shared WINDOW (the wrap-around of a slot pointer is a bit mask)
  float* edge = reinterpret_cast<float*>(smem + plan.edge_off + ((1024u - ((smem_u32(smem) + plan.edge_off) & 1023u)) & 1023u));
  uint8_t* ring = smem + plan.ring_off;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int W = plan.W;
  long long ol = out_lens[b], il = in_lens[b];
  const int olen = (int)(ol < 0 ? 0 : (ol > T1 ? T1 : ol));
  const int ilen = (int)(il < 0 ? 0 : (il > T2 ? T2 : il));
  uint32_t* bits_s = reinterpret_cast<uint32_t*>(smem + plan.bits_off);
  uint32_t* bits_g = bits_ws + (kBitsInSmem ? 0 : (size_t)b * T1 * plan.Wb);
  auto put_bits = [&](size_t idx, uint32_t w) {
    if constexpr (kBitsInSmem) bits_s[idx] = w; else bits_g[idx] = w;
  };
  auto get_bits = [&](size_t idx) -> uint32_t {
    if constexpr (kBitsInSmem) return bits_s[idx]; else return bits_g[idx];
  };
  const int R = plan.rows_per_chunk;
  const int stages = plan.stages;
  const bool active = (olen > 0 && ilen > 0);
  const int nchunks = active ? (olen + R - 1) / R : 0;
  // only the warps that hold live columns take part in the DP
  const int Wl = active ? (ilen + 32 * plan.NC - 1) / (32 * plan.NC) : 0;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (uint32_t)(Wl > 0 ? Wl : 1));
    }
    mbar_fence_init();
  }
  for (int j = tid; j < T2; j += nthreads) dur_s[j] = 0;
  for (int i = tid; i < (W + 1) * kMasEdge; i += nthreads) reinterpret_cast<uint32_t*>(edge)[i] = kMasSentinel;
  if (active)
    for (int j = tid; j < plan.Wb; j += nthreads) put_bits(j, 0u);   // row 0 carries no decisions
  __syncthreads();

  const size_t slab = (size_t)b * T1 * T2;
  if (b == 0 && tid == 0) { g_mas_timeline[0] = gtime(); g_mas_timeline[6] = (unsigned long long)clock64(); }

  if (warp == W) {
    // ------------------------------ producer ------------------------------
    if (lane == 0) {
      const unsigned long long total16 = total_bytes & ~15ull;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % stages;
        const int k = c / stages;
        if (k > 0) mbar_wait_relaxed(&empty[s], (uint32_t)((k & 1) ^ 1));
        const int r0 = c * R;
        const int r1 = min(r0 + R, olen);
        const unsigned long long sb = (unsigned long long)(slab + (size_t)r0 * T2) * 4ull;
        const unsigned long long eb = (unsigned long long)(slab + (size_t)r1 * T2) * 4ull;
        // 128-byte aligned source window: a bulk copy whose source is only 16-byte aligned runs at a fraction of the
        // bandwidth (measured ~5 GB/s per SM), so fetch a few bytes more and index into the stage with the offset
        const unsigned long long s16 = sb & ~127ull;
        unsigned long long e16 = (eb + 127ull) & ~127ull;
        if (e16 > total16) e16 = total16;
        uint8_t* dst = ring + (size_t)s * plan.stage_bytes;
        const uint32_t nbytes = e16 > s16 ? (uint32_t)(e16 - s16) : 0u;
        const unsigned long long t0 = (s16 + nbytes > sb) ? (s16 + nbytes) : sb;
        for (unsigned long long x = t0; x < eb; x += 4)   // at most 12 bytes past the last 16-byte boundary
          *reinterpret_cast<float*>(dst + (x - s16)) =
              *reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(logp) + x);
        mbar_arrive_expect_tx(&full[s], nbytes);
        if (nbytes) bulk_g2s(dst, reinterpret_cast<const uint8_t*>(logp) + s16, nbytes, &full[s]);
      }
    }
  } else if (warp < W) {
    // ------------------------------ DP warps (skewed wavefront) ------------------------------
    if (warp < Wl) {
      auto run = [&](auto nc_tag) {
        constexpr int kNC = decltype(nc_tag)::value;
        if (warp == 0)
          dp_warp<kBitsInSmem, true, kNC>(full, empty, edge, ring, bits_s, bits_g, plan, slab, T2, olen, nchunks, warp,
                                          lane, b == 0 && warp == Wl - 1);
        else
          dp_warp<kBitsInSmem, false, kNC>(full, empty, edge, ring, bits_s, bits_g, plan, slab, T2, olen, nchunks, warp,
                                           lane, b == 0 && warp == Wl - 1);
      };
      switch (plan.NC) {
        case 1: run(std::integral_constant<int, 1>{}); break;
        case 2: run(std::integral_constant<int, 2>{}); break;
        case 3: run(std::integral_constant<int, 3>{}); break;
        default: run(std::integral_constant<int, 4>{}); break;
      }
      if (b == 0 && tid == 0) { g_mas_timeline[1] = gtime(); g_mas_timeline[7] = (unsigned long long)clock64(); }
    }
  } else {
    // ------------------------------ zero fill ------------------------------
    zero_fill(hard + slab, (size_t)T1 * T2, tid - (W + 1) * 32, kMasFillWarps * 32);
    if (b == 0 && tid == (W + 1) * 32) g_mas_timeline[5] = gtime();
  }
  __syncthreads();
  if (b == 0 && tid == 0) g_mas_timeline[2] = gtime();

  // ------------------------------ backtrack: 32 rows per round, bit windows in registers ------------------------------
  if (active && warp == 0) {
    int j = ilen - 1;                                  // column at row `top`
    for (int top = olen - 1; top >= 0; top -= 32) {
      const int row = top - lane;                      // lane l looks at row top - l
      uint32_t win = 0;                                // bit 31 <-> column j, bit k <-> column j - 31 + k
      if (row >= 1) {
        const int wj = j >> 5;
        const uint32_t hi = get_bits((size_t)row * plan.Wb + wj);
        const uint32_t lo = wj > 0 ? get_bits((size_t)row * plan.Wb + wj - 1) : 0u;
        const unsigned long long x = ((unsigned long long)hi << 32) | lo;
        win = (uint32_t)(x >> ((j & 31) + 1));
      }
      // all 32 windows into registers FIRST: with the shuffle issued inside the step loop it sat on the dependent
      // chain (shfl 24 + 3 ALU ops = ~40 cycles per row); hoisted, a step is shift / and / sub = ~15 cycles
      uint32_t wk[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) wk[k] = __shfl_sync(0xffffffffu, win, k);
      int pos = 31, mycol = 0;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        mycol = (lane == k) ? j - (31 - pos) : mycol;
        pos -= (int)((wk[k] >> pos) & 1u);             // rows <= 0 have win == 0: no movement
      }
      if (row >= 0) path[row] = (uint16_t)mycol;
      j -= (31 - pos);
    }
  }
  __syncthreads();
  if (b == 0 && tid == 0) g_mas_timeline[3] = gtime();

  // ------------------------------ scatter ------------------------------
  if (active) {
    for (int i = tid; i < olen; i += nthreads) {
      const int jj = path[i];
      hard[slab + (size_t)i * T2 + jj] = 1.0f;
      if (f2t) f2t[(size_t)b * T1 + i] = jj;
      atomicAdd(&dur_s[jj], 1);
    }
    if (tid == 0) hard[slab] = 1.0f;  // alignment.py:59
  }
  if (f2t)
    for (int i = olen + tid; i < T1; i += nthreads) f2t[(size_t)b * T1 + i] = -1;
  __syncthreads();
  if (dur) {
    if (active && tid == 0 && path[0] != 0) dur_s[0] += 1;
    __syncthreads();
    for (int jj = tid; jj < T2; jj += nthreads) dur[(size_t)b * T2 + jj] = dur_s[jj];
  }
  if (b == 0 && tid == 0) g_mas_timeline[4] = gtime();
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
  static bool configured = false;
  if (!configured) {
    RB_CUDA(cudaFuncSetAttribute(mas_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    RB_CUDA(cudaFuncSetAttribute(mas_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget));
    configured = true;
  }
  const unsigned long long total_bytes = (unsigned long long)B * T1 * T2 * 4ull;
  if (plan.bits_in_smem)
    mas_kernel<true><<<B, plan.threads, plan.smem_bytes, stream>>>(logp, in_lens, out_lens, T1, T2, total_bytes, attn_hard,
                                                                  frame_to_token, durations, bits_ws, plan);
  else
    mas_kernel<false><<<B, plan.threads, plan.smem_bytes, stream>>>(logp, in_lens, out_lens, T1, T2, total_bytes, attn_hard,
                                                                   frame_to_token, durations, bits_ws, plan);
  return after_launch();
}

extern "C" int radtts_mas_debug_timeline(unsigned long long* out16_host) {
  if (!out16_host) return RADTTS_ERR_INVALID_ARG;
  RB_CUDA(cudaDeviceSynchronize());
  RB_CUDA(cudaMemcpyFromSymbol(out16_host, g_mas_timeline, 16 * sizeof(unsigned long long)));
  return 0;
}
