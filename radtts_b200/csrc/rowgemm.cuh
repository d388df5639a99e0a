// "Row GEMM": the one contraction shape every conv of the RADTTS hot path reduces to once frames are
// packed channels-last (see frameplan.cuh):
//
//     Y[r, n] = epilogue( sum_seg sum_k  A_seg[r + shift_seg, kcol_seg + k] * W[n, kofs_seg + k] )
//
// A segment is a (buffer, row shift, column window); a dilated k-tap conv is k segments over the same buffer
// with shifts (t - k/2) * dilation, a concatenated input (z0 | context) is two segments, a fused
// dgrad (res_skip^T + in_layer^T) is 1 + 5 segments.  W is [N][K_total] with K contiguous.
//
// Two engines share the epilogue functors below:
//   * rowgemm_simt<Epi>  -- fp32 SIMT, exact-order fp32 FMA accumulation: the fp32 parity path
//   * rowgemm_tc<Epi>    -- bf16 tcgen05/TMEM/TMA (rowgemm_tc.cuh): the performance path
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace rb {

constexpr int kMaxSeg = 8;

struct Seg {
  const void* a;  // activation buffer [rows_alloc][lda]
  int lda;        // elements
  int shift;      // row shift applied to the output row index
  int kcol;       // first column of the window
  int klen;       // window length (multiple of 16 for SIMT, 64 for TC)
};

struct GemmDesc {
  Seg seg[kMaxSeg];
  int nseg;
  const void* w;   // [N][ldw], K contiguous, segments consecutive along K
  int ldw;
  int N;           // output columns (multiple of 8)
  int rows_alloc;  // rows of every A buffer; rows outside [0, rows_alloc) read as zero
  const int* plan; // plan[0] = number of packed rows in use (tiles beyond are skipped)
};

// ------------------------------------------------------------------------------------------------------
// element helpers
// ------------------------------------------------------------------------------------------------------
template <typename T>
struct Act;
template <>
struct Act<float> {
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  template <int W>
  static __device__ __forceinline__ void ldv(const float* p, float (&v)[W]) {
#pragma unroll
    for (int i = 0; i < W; i += 4) {
      float4 t = *reinterpret_cast<const float4*>(p + i);
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  }
  template <int W>
  static __device__ __forceinline__ void stv(float* p, const float (&v)[W]) {
#pragma unroll
    for (int i = 0; i < W; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
};
template <>
struct Act<__nv_bfloat16> {
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  template <int W>
  static __device__ __forceinline__ void ldv(const __nv_bfloat16* p, float (&v)[W]) {
#pragma unroll
    for (int i = 0; i < W; i += 8) {
      uint4 t = *reinterpret_cast<const uint4*>(p + i);
      const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[i + 2 * j] = __uint_as_float(u[j] << 16);
        v[i + 2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
      }
    }
  }
  template <int W>
  static __device__ __forceinline__ void stv(__nv_bfloat16* p, const float (&v)[W]) {
#pragma unroll
    for (int i = 0; i < W; i += 8) {
      uint32_t u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[i + 2 * j], v[i + 2 * j + 1]);
        u[j] = *reinterpret_cast<uint32_t*>(&h);
      }
      *reinterpret_cast<uint4*>(p + i) = make_uint4(u[0], u[1], u[2], u[3]);
    }
  }
};

// torch.nn.Softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplus20(float x) { return x > 20.f ? x : log1pf(expf(x)); }
// d softplus / dx expressed through the OUTPUT y = softplus(x):  sigmoid(x) = 1 - exp(-y)
__device__ __forceinline__ float softplus_grad_from_out(float y) { return 1.f - expf(-y); }
// bf16 path: MUFU-based versions (their ~1e-6 error is far below the bf16 rounding of the stored result)
// branch-free: for x > 20 the correction term is below fp32 resolution, so the threshold needs no select
__device__ __forceinline__ float softplus20_fast(float x) { return fmaxf(x, 0.f) + __logf(1.f + __expf(-fabsf(x))); }
__device__ __forceinline__ float softplus_grad_from_out_fast(float y) { return 1.f - __expf(-y); }
template <typename T>
__device__ __forceinline__ float softplus_t(float x) { return sizeof(T) == 2 ? softplus20_fast(x) : softplus20(x); }
template <typename T>
__device__ __forceinline__ float softplus_grad_t(float y) {
  return sizeof(T) == 2 ? softplus_grad_from_out_fast(y) : softplus_grad_from_out(y);
}

// Row validity + partial-conv renormalisation (reference partialconv1d.py:51-56): for a frame at position
// `pos` of an utterance with `rem` frames after it, a k-tap conv with dilation 2^log2d sees
// cnt = 1 + min(half, pos/d) + min(half, rem/d) valid taps and the output is scaled by k / (cnt + 1e-6).
struct RowMeta {
  const int* pos;  // -1 on gap rows
  const int* rem;
  __device__ __forceinline__ bool valid(int row) const { return pos[row] >= 0; }
  __device__ __forceinline__ float ratio(int row, int log2d, int ksize) const {
    const int half = ksize >> 1;
    const int cnt = 1 + min(half, pos[row] >> log2d) + min(half, rem[row] >> log2d);
    return (float)ksize / ((float)cnt + 1e-6f);
  }
};

// Per-row state an epilogue needs (validity, partial-conv ratio): computed ONCE per row per tile and kept in
// registers -- re-reading pos/rem from global memory for every 16-column chunk put a full memory round trip on
// the epilogue's critical path (ncu: 80 % of epilogue stalls were long-scoreboard on exactly these loads).
struct RowState {
  bool ok;
  float rt;
};

// ------------------------------------------------------------------------------------------------------
// epilogue functors:
//   RowState prep(int row)                                  once per row per tile
//   const float* colvec()                                   per-column vector (bias) the engine stages per tile, or null
//   template<int W> void operator()(int row, int col0, const float (&acc)[W], const RowState& rs, const float (&cv)[W])
//        cv[i] = colvec()[col0 + i], handed over in registers (zeros when colvec() is null)
// `row` < rows in use, col0 % W == 0, col0 + W <= round_up(N, W); functors guard col < N themselves when
// their N is not a multiple of W.
// ------------------------------------------------------------------------------------------------------
enum : int { ACT_NONE = 0, ACT_SOFTPLUS = 1, ACT_RELU = 2 };

// y = act(acc * ratio + bias) on valid rows, 0 on gap rows.            (start / in_layers / res_skip / ReLU convs)
template <typename T, int ACT = -1>   // ACT >= 0: activation fixed at compile time (branch-free inner loop)
struct EpiBiasAct {
  T* out; int ldo; int col_off;
  const float* bias;
  RowMeta meta;
  int act;           // used when ACT < 0
  int partial;       // 1: multiply acc by the partial-conv ratio
  int log2d, ksize;
  int mask_rows;     // 1: zero gap rows; 0: plain conv (ConvAttention projections)
  __device__ __forceinline__ RowState prep(int row) const {
    RowState rs;
    rs.ok = !mask_rows || meta.valid(row);
    rs.rt = (partial && rs.ok) ? meta.ratio(row, log2d, ksize) : 1.f;
    return rs;
  }
  __device__ __forceinline__ const float* colvec() const { return bias; }
  template <int W>
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[W], const RowState& rs,
                                             const float (&cv)[W]) const {
    float y[W];
    const bool ok = rs.ok;
    const float rt = rs.rt;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      float v = acc[i] * rt + cv[i];
      const int a = ACT >= 0 ? ACT : act;
      if (a == ACT_SOFTPLUS) v = softplus_t<T>(v);
      else if (a == ACT_RELU) v = fmaxf(v, 0.f);
      y[i] = ok ? v : 0.f;
    }
    Act<T>::template stv<W>(out + (size_t)row * ldo + col_off + col0, y);
  }
};

// Plain fp32 store (used for 1x1 invertible conv and generic dgrads):  out[row][col] = valid ? acc : 0
struct EpiStoreF32 {
  float* out; int ldo;
  RowMeta meta;
  int mask_rows;
  __device__ __forceinline__ RowState prep(int row) const { return RowState{!mask_rows || meta.valid(row), 1.f}; }
  __device__ __forceinline__ const float* colvec() const { return nullptr; }
  template <int W>
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[W], const RowState& rs,
                                             const float (&)[W]) const {
    float y[W];
    const bool ok = rs.ok;
#pragma unroll
    for (int i = 0; i < W; ++i) y[i] = ok ? acc[i] : 0.f;
    Act<float>::stv<W>(out + (size_t)row * ldo + col0, y);
  }
};

// Invertible 1x1 conv epilogue (forward direction): zmid = W z; copies the columns that are final for this
// flow (exited channels + z0) into zout and drops a (possibly bf16) copy of z0 for the `start` GEMM.
template <typename T>
struct EpiInvConv {
  float* zmid; float* zout; T* z0;  // zout / z0 may be null
  int c_off, h, zld;                // active block starts at c_off; z0 = [c_off, c_off + h)
  RowMeta meta;
  __device__ __forceinline__ RowState prep(int row) const { return RowState{meta.valid(row), 1.f}; }
  __device__ __forceinline__ const float* colvec() const { return nullptr; }
  template <int W>
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[W], const RowState& rs,
                                             const float (&)[W]) const {
    const bool ok = rs.ok;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      const int c = col0 + i;
      const float v = ok ? acc[i] : 0.f;
      zmid[(size_t)row * zld + c] = v;
      if (zout && c < c_off + h) zout[(size_t)row * zld + c] = v;
      if (z0 && c >= c_off && c < c_off + 128) {
        const float zv = (c < c_off + h) ? v : 0.f;
        if (sizeof(T) == 4) reinterpret_cast<float*>(z0)[(size_t)row * 128 + (c - c_off)] = zv;
        else reinterpret_cast<__nv_bfloat16*>(z0)[(size_t)row * 128 + (c - c_off)] = __float2bfloat16(zv);
      }
    }
  }
};

// `end` conv + affine coupling.  Columns are interleaved pairs (2c: raw scale, 2c+1: translation) so that a
// chunk always holds complete channels.  scaling 'tanh': s = tanh(x) + 1 + 1e-6 (reference common.py:782-784).
//   forward : z1' = s * z1 + b, log_s = log s          inverse : z1 = (z1' - b) / s
// kFast (the bf16 tensor-core engine): MUFU forms of tanh / exp / log (relative error ~1e-6 -- the engine's operands are
// bf16), and the scaling mode dispatched ONCE per call.  With the libm forms and the four scaling modes expanded inside the
// fully unrolled chunk loops this epilogue was 23 K SASS instructions: it missed the instruction cache all the way
// through (ncu: `stalled_no_instruction` the top stall) and made the 160-wide `end` GEMM take 84 us where the same GEMM
// with a plain epilogue takes 38 (tools/gemm_micro.cu, last problem).
// The scaling mode S is a template parameter as well (the host picks the instantiation): one variant per kernel keeps the
// unrolled epilogue at a quarter of the code.
template <bool kFast, int S>
struct EpiCouplingT {
  const float* bias;      // interleaved, length 2h (padded with zeros to N)
  const float* zsrc;      // [rows][160]: z1 is read from column c_off + h + c
  float* zdst;            // [rows][160]: written at the same column
  float* log_s;           // [rows][80] (forward only, may be null)
  float* params;          // [rows][160] raw (x, b) pairs saved for backward (may be null)
  int c_off, h, zld;
  int inverse;            // S: 0 tanh, 1 exp, 2 sigmoid, 3 translate (reference common.py:775-787)
  RowMeta meta;
  __device__ __forceinline__ RowState prep(int row) const { return RowState{meta.valid(row), 1.f}; }
  __device__ __forceinline__ const float* colvec() const { return bias; }
  static __device__ __forceinline__ void scale_of(float x, float& s, float& lsv) {
    if (kFast) {
      if (S == 0) { s = ((1.f - __fdividef(2.f, __expf(2.f * x) + 1.f)) + 1.f) + 1e-6f; lsv = __logf(s); }
      else if (S == 1) { s = __expf(x); lsv = x; }
      else if (S == 2) { s = __fdividef(1.f, 1.f + __expf(-(x + 10.f))) + 1e-6f; lsv = __logf(s); }
      else { s = 1.f; lsv = 0.f; }
    } else {
      if (S == 0) { s = (tanhf(x) + 1.f) + 1e-6f; lsv = logf(s); }
      else if (S == 1) { s = expf(x); lsv = x; }
      else if (S == 2) { s = 1.f / (1.f + expf(-(x + 10.f))) + 1e-6f; lsv = logf(s); }
      else { s = 1.f; lsv = 0.f; }
    }
  }
  template <int W>
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[W], const RowState& rs,
                                             const float (&cv)[W]) const {
    const bool ok = rs.ok;
    constexpr int NP = W / 2;                 // channel pairs in this chunk
    const int c0 = col0 >> 1;
    const size_t zi0 = (size_t)row * zld + c_off + h + c0;
    // A thread owns one row and W consecutive columns: its outputs are contiguous runs (W floats of params, W/2 of z
    // and of log_s).  Written as 16-byte vectors where the run is whole and aligned -- as scalars every warp-wide store
    // touched 32 different rows, one sector each, and the 160-wide `end` GEMM spent most of its time in this epilogue.
    if (NP % 4 == 0 && c0 + NP <= h) {
      const bool zvec = ((c_off + h + c0) & 3) == 0;   // the z columns start at c_off + h: aligned for some flows only
      float z1[NP], outv[NP], ls[NP], pv[W];
      if (ok) {
        if (zvec) {
#pragma unroll
          for (int i = 0; i < NP; i += 4) {
            const float4 t = *reinterpret_cast<const float4*>(zsrc + zi0 + i);
            z1[i] = t.x; z1[i + 1] = t.y; z1[i + 2] = t.z; z1[i + 3] = t.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < NP; ++i) z1[i] = zsrc[zi0 + i];
        }
      }
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float x = acc[2 * i] + cv[2 * i];
        const float b = acc[2 * i + 1] + cv[2 * i + 1];
        outv[i] = 0.f; ls[i] = 0.f;
        if (ok) {
          float sc, lsv;
          scale_of(x, sc, lsv);
          if (inverse) outv[i] = kFast ? __fdividef(z1[i] - b, sc) : (z1[i] - b) / sc;
          else { outv[i] = sc * z1[i] + b; ls[i] = lsv; }
        }
        pv[2 * i] = ok ? x : 0.f;
        pv[2 * i + 1] = ok ? b : 0.f;
      }
      if (zvec) {
#pragma unroll
        for (int i = 0; i < NP; i += 4)
          *reinterpret_cast<float4*>(zdst + zi0 + i) = make_float4(outv[i], outv[i + 1], outv[i + 2], outv[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < NP; ++i) zdst[zi0 + i] = outv[i];
      }
      if (log_s) {
#pragma unroll
        for (int i = 0; i < NP; i += 4)
          *reinterpret_cast<float4*>(log_s + (size_t)row * (zld / 2) + c0 + i) = make_float4(ls[i], ls[i + 1], ls[i + 2], ls[i + 3]);
      }
      if (params) {
#pragma unroll
        for (int i = 0; i < W; i += 4)
          *reinterpret_cast<float4*>(params + (size_t)row * zld + col0 + i) = make_float4(pv[i], pv[i + 1], pv[i + 2], pv[i + 3]);
      }
      return;
    }
#pragma unroll
    for (int i = 0; i < W; i += 2) {
      const int c = (col0 + i) >> 1;
      if (c >= h) continue;
      const float x = acc[i] + cv[i];
      const float b = acc[i + 1] + cv[i + 1];
      const size_t zi = (size_t)row * zld + c_off + h + c;
      float outv = 0.f, ls = 0.f;
      if (ok) {
        float sc, lsv;
        scale_of(x, sc, lsv);
        const float z1 = zsrc[zi];
        if (inverse) outv = (z1 - b) / sc;
        else { outv = sc * z1 + b; ls = lsv; }
      }
      zdst[zi] = outv;
      if (log_s) log_s[(size_t)row * (zld / 2) + c] = ls;
      if (params) { params[(size_t)row * zld + 2 * c] = ok ? x : 0.f; params[(size_t)row * zld + 2 * c + 1] = ok ? b : 0.f; }
    }
  }
};

// ------------------------------------------------------------------------------------------------------
// backward epilogues
// ------------------------------------------------------------------------------------------------------
// dgrad through a softplus'ed (partial) conv output x:  g_v = acc * softplus'(x) * ratio   (valid rows)
// with act == ACT_NONE / partial == 0 it is a plain masked store (g_x0 of `start`).
template <typename T, int ACT = -1>
struct EpiDgradAct {
  const T* x; int ldx;    // forward OUTPUT of the layer whose pre-activation gradient is produced
  T* out; int ldo;
  RowMeta meta;
  int act, partial, log2d, ksize;
  __device__ __forceinline__ RowState prep(int row) const {
    RowState rs;
    rs.ok = meta.valid(row);
    rs.rt = (partial && rs.ok) ? meta.ratio(row, log2d, ksize) : 1.f;
    return rs;
  }
  __device__ __forceinline__ const float* colvec() const { return nullptr; }
  template <int W>
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[W], const RowState& rs,
                                             const float (&)[W]) const {
    const bool ok = rs.ok;
    float y[W];
    if (ok) {
      float xv[W];
      const int a = ACT >= 0 ? ACT : act;
      if (a == ACT_SOFTPLUS) Act<T>::template ldv<W>(x + (size_t)row * ldx + col0, xv);
      const float rt = rs.rt;
#pragma unroll
      for (int i = 0; i < W; ++i) y[i] = acc[i] * (a == ACT_SOFTPLUS ? softplus_grad_t<T>(xv[i]) : 1.f) * rt;
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) y[i] = 0.f;
    }
    Act<T>::template stv<W>(out + (size_t)row * ldo + col0, y);
  }
};

// dgrad of `start`: columns [0, ctx_ld) -> g_ctx (fp32, optionally accumulated over flows),
// columns [ctx_ld, ctx_ld + h) -> g_zmid[:, c_off + c] += acc   (gradient reaching z0 through the WN)
struct EpiStartDgrad {
  float* g_ctx; int n_ctx; int ctx_ld;
  float* g_zmid; int c_off, h, zld;
  int accumulate;
  RowMeta meta;
  __device__ __forceinline__ RowState prep(int row) const { return RowState{meta.valid(row), 1.f}; }
  __device__ __forceinline__ const float* colvec() const { return nullptr; }
  template <int W>
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[W], const RowState& rs,
                                             const float (&)[W]) const {
    const bool ok = rs.ok;
    if (col0 + W <= ctx_ld) {
      float y[W];
      float* dst = g_ctx + (size_t)row * ctx_ld + col0;
      if (accumulate) Act<float>::ldv<W>(dst, y);
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const float v = (ok && col0 + i < n_ctx) ? acc[i] : 0.f;
        y[i] = accumulate ? y[i] + v : v;
      }
      Act<float>::stv<W>(dst, y);
    } else {
#pragma unroll
      for (int i = 0; i < W; ++i) {
        const int c = col0 + i - ctx_ld;
        if (ok && c >= 0 && c < h) g_zmid[(size_t)row * zld + c_off + c] += acc[i];
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------------
// SIMT fp32 engine: 128 x 128 output tile, BK = 16, 256 threads x (8 x 8) micro-tiles, register prefetch.
// ------------------------------------------------------------------------------------------------------
constexpr int kSimtBM = 128, kSimtBN = 128, kSimtBK = 16, kSimtThreads = 256;

template <typename Epi>
__global__ void __launch_bounds__(kSimtThreads) rowgemm_simt(GemmDesc d, Epi epi) {
  const int rows_used = d.plan ? d.plan[0] : d.rows_alloc;
  const int row0 = blockIdx.x * kSimtBM;
  if (row0 >= rows_used) return;
  const int n0 = blockIdx.y * kSimtBN;

  __shared__ __align__(16) float As[2][kSimtBK][kSimtBM + 4];
  __shared__ __align__(16) float Bs[2][kSimtBK][kSimtBN + 4];

  const int tid = threadIdx.x;
  const int lr = tid & 127;   // row (A) / column (W) this thread stages
  const int lk = (tid >> 7) * 8;  // k half
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // flat list of k-steps over all segments
  int total_steps = 0;
  for (int s = 0; s < d.nseg; ++s) total_steps += d.seg[s].klen / kSimtBK;

  float ra[8], rb_[8];
  int seg_i = 0, seg_step = 0, kofs = 0;

  auto fetch = [&](int si, int sstep, int kof) {
    const Seg& sg = d.seg[si];
    const int ar = row0 + lr + sg.shift;
    if (ar >= 0 && ar < d.rows_alloc) {
      const float* p = reinterpret_cast<const float*>(sg.a) + (size_t)ar * sg.lda + sg.kcol + sstep * kSimtBK + lk;
      const float4 t0 = *reinterpret_cast<const float4*>(p);
      const float4 t1 = *reinterpret_cast<const float4*>(p + 4);
      ra[0] = t0.x; ra[1] = t0.y; ra[2] = t0.z; ra[3] = t0.w; ra[4] = t1.x; ra[5] = t1.y; ra[6] = t1.z; ra[7] = t1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) ra[i] = 0.f;
    }
    const int wn = n0 + lr;
    if (wn < d.N) {
      const float* p = reinterpret_cast<const float*>(d.w) + (size_t)wn * d.ldw + kof + lk;
      const float4 t0 = *reinterpret_cast<const float4*>(p);
      const float4 t1 = *reinterpret_cast<const float4*>(p + 4);
      rb_[0] = t0.x; rb_[1] = t0.y; rb_[2] = t0.z; rb_[3] = t0.w; rb_[4] = t1.x; rb_[5] = t1.y; rb_[6] = t1.z; rb_[7] = t1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) rb_[i] = 0.f;
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      As[buf][lk + i][lr] = ra[i];
      Bs[buf][lk + i][lr] = rb_[i];
    }
  };
  auto advance = [&]() {
    kofs += kSimtBK;
    if (++seg_step == d.seg[seg_i].klen / kSimtBK) { seg_step = 0; ++seg_i; }
  };

  fetch(seg_i, seg_step, kofs);
  advance();
  stage(0);
  __syncthreads();
  for (int step = 0; step < total_steps; ++step) {
    const int buf = step & 1;
    if (step + 1 < total_steps) { fetch(seg_i, seg_step, kofs); advance(); }
#pragma unroll
    for (int k = 0; k < kSimtBK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (step + 1 < total_steps) {
      stage(buf ^ 1);
      __syncthreads();
    }
  }
  const int col0 = n0 + tx * 8;
  if (col0 < d.N) {
    const float* cv = epi.colvec();
    float cvr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) cvr[j] = cv ? cv[col0 + j] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row0 + ty * 8 + i;  // rows past the packed range are gap rows (functors write zeros)
      epi.template operator()<8>(row, col0, acc[i], epi.prep(row), cvr);
    }
  }
}

template <typename Epi>
inline int launch_rowgemm_simt(const GemmDesc& d, const Epi& epi, cudaStream_t stream) {
  dim3 grid(d.rows_alloc / kSimtBM, ceil_div(d.N, kSimtBN));
  rowgemm_simt<Epi><<<grid, kSimtThreads, 0, stream>>>(d, epi);
  return after_launch();
}

}  // namespace rb
