// bf16 tcgen05 engine of the row GEMM (see rowgemm.cuh for the contraction it computes).
//
// Persistent, warp-specialised, one CTA per SM (roles on the warps the kernel body assigns: see the note there):
//   warps 0-3   epilogue: tcgen05.ld of the fp32 accumulator (lane == output row), fused epilogue functor
//               (bias / partial-conv renormalisation / softplus / affine coupling / ...), direct global stores
//   warp 4      TMA producer: per k-block one 128x64 activation box (row-shifted per segment -- the dilated conv
//               taps are nothing but different row coordinates of the same tensor map; out-of-range rows are
//               zero-filled by the TMA unit) and one BNx64 weight box, 128B-swizzled, into a 4-stage smem ring
//   warp 5      MMA issuer: one elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16) x4 per k-block,
//               accumulating in TMEM; tcgen05.commit releases smem stages and publishes finished accumulators
// Two 256-column TMEM accumulators (all 512 columns) double-buffer the epilogue against the next tile's MMAs.
//
// Tile width and ring depth are chosen ON THE DEVICE, per launch: the number of row tiles comes from the frame plan (device
// memory, no host sync), and 85 row tiles x 4 column tiles of 256 on 148 SMs are 2.30 waves that run as 3 (the measured
// 0.72-of-peak of the in_layer GEMM was this and nothing else).  The host passes up to kTcCand tile widths with a weight
// tensor map each; every CTA evaluates rounds(bn) x (bn + fixed) for each and takes the cheapest (e.g. 208 -> 425 tiles
// = 2.87 waves, 19 % fewer MMA cycles per CTA), and sizes the TMA ring to what a stage of that width leaves room for
// (bn = 256: 4 stages, 208 / 176: 5, <= 144: 6 -- a narrower tile needs the deeper ring to stay MMA- and not
// fetch-latency bound).  The accumulation order of every output element is the same for every width: results are
// bit-identical whichever candidate wins.
#pragma once
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "ptx.cuh"
#include "rowgemm.cuh"

namespace rb {

constexpr int kTcBM = 128, kTcBK = 64, kTcMaxBN = 256, kTcMaxStages = 6, kTcThreads = 192;
constexpr int kTcABytes = kTcBM * kTcBK * 2;       // 16 KB
constexpr int kTcRingBytes = 208 * 1024;           // 4 x 48 KB (bn 256), 5 x 41.6 (208), 5 x 38 (176), 6 x 34 (144), 6 x 32 (128)
constexpr int kTcSmemBytes = kTcRingBytes + 1024 /*align slack*/ + 256 /*barriers*/ +
                             2 * kTcMaxBN * 4 /*per-tile column vector (bias), double buffered*/;
constexpr int kTcMaxMaps = 4;
constexpr int kTcCand = 4;                         // tile-width candidates per launch
constexpr int kTcTileFixed = 20;                   // per-tile fixed cost in units of one output column's MMA time

struct TcSeg {
  int map;    // index into TcParams::amap
  int shift;  // row shift
  int kcol;   // first column
  int kblocks;
};

struct TcParams {
  CUtensorMap amap[kTcMaxMaps];
  CUtensorMap wmap[kTcCand];   // weight map per candidate tile width (box = bn rows x 64 columns)
  TcSeg seg[kMaxSeg];
  int nseg;
  int N;                // output columns
  int bn[kTcCand];      // candidate tile widths (multiples of 16, <= 256); N is covered by ceil(N / bn) tiles
  int ncand;
  int rows_alloc;
  const int* plan;
};

__host__ __device__ inline int tc_stages_for(int bn) {
  const int st = kTcRingBytes / (kTcABytes + bn * kTcBK * 2);
  return st > kTcMaxStages ? kTcMaxStages : st;
}

// ----------------------------------------------------------------------------------------------------------
// host: tensor-map construction (driver entry point fetched through the runtime; no libcuda link dependency)
// ----------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// 2-D bf16 row-major [rows][ld] tensor, box = box_rows x 64 columns, 128-byte swizzle.
inline int make_map_bf16(const void* base, int ld, int cols, int rows, int box_rows, CUtensorMap* out) {
  struct Key {
    const void* p; int ld, cols, rows, box;
    bool operator==(const Key& o) const { return p == o.p && ld == o.ld && cols == o.cols && rows == o.rows && box == o.box; }
  };
  struct Hash {
    size_t operator()(const Key& k) const {
      size_t h = reinterpret_cast<size_t>(k.p);
      h ^= (size_t)k.ld * 0x9E3779B97F4A7C15ull + (size_t)k.rows * 0xC2B2AE3D27D4EB4Full + (size_t)k.box * 1315423911u +
           (size_t)k.cols * 2654435761u;
      return h;
    }
  };
  static std::unordered_map<Key, CUtensorMap, Hash> cache;
  static std::mutex mu;
  Key key{base, ld, cols, rows, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return RADTTS_ERR_NOT_SM100;
  if (((uintptr_t)base & 15) || (ld % 8)) return RADTTS_ERR_INVALID_ARG;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kTcBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return RADTTS_ERR_INVALID_ARG;
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
  }
  return 0;
}

// ----------------------------------------------------------------------------------------------------------
// device
// ----------------------------------------------------------------------------------------------------------
template <typename Epi>
__global__ void __launch_bounds__(kTcThreads, 1) rowgemm_tc_kernel(const __grid_constant__ TcParams p, const Epi epi) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_used = p.plan ? p.plan[0] : p.rows_alloc;
  const int n_tiles_m = (rows_used + kTcBM - 1) / kTcBM;
  // tile width: fewest MMA cycles per CTA over the whole launch (every CTA computes the same answer)
  int cand = 0;
  {
    long long best = -1;
    for (int c = 0; c < p.ncand; ++c) {
      const int tiles = n_tiles_m * ((p.N + p.bn[c] - 1) / p.bn[c]);
      const long long cost = (long long)((tiles + (int)gridDim.x - 1) / (int)gridDim.x) * (p.bn[c] + kTcTileFixed);
      if (best < 0 || cost < best) { best = cost; cand = c; }
    }
  }
  const int bn = p.bn[cand];
  const int n_tiles_n = (p.N + bn - 1) / bn;
  const int n_tiles = n_tiles_m * n_tiles_n;
  const int nst = tc_stages_for(bn);
  const int b_bytes = bn * kTcBK * 2;
  const CUtensorMap* wmap = &p.wmap[cand];
  int kb_total = 0;
  for (int s = 0; s < p.nseg; ++s) kb_total += p.seg[s].kblocks;

  uint8_t* sA = smem;
  uint8_t* sB = smem + nst * kTcABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTcRingBytes);
  uint64_t* full = bars;                       // [kTcMaxStages]
  uint64_t* empty = bars + kTcMaxStages;       // [kTcMaxStages]
  uint64_t* tfull = bars + 2 * kTcMaxStages;   // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* cvec_s = reinterpret_cast<float*>(smem + kTcRingBytes + 256);  // [2][kTcMaxBN]

  if (threadIdx.x == 0) {
    for (int s = 0; s < kTcMaxStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_fence_init();
  }
  if (warp == 4 && lane == 0) {
    for (int m = 0; m < kTcMaxMaps; ++m) tma_prefetch_desc(&p.amap[m]);
    tma_prefetch_desc(wmap);
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Role -> warp mapping: the SM's arbiter prefers HIGHER warp ids, so the two latency-critical single-thread roles
  // (TMA producer, MMA issuer) sit on warps 4 and 5 and win issue slots over the instruction-heavy epilogue warps 0-3.
  if (warp == 4) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)kTcABytes + (uint32_t)b_bytes;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tm = tile % n_tiles_m, tn = tile / n_tiles_m;
        const int row0 = tm * kTcBM, n0 = tn * bn;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const TcSeg sg = p.seg[s];
          for (int kk = 0; kk < sg.kblocks; ++kk, ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], tx_bytes);
            tma_load_2d(sA + stage * kTcABytes, &p.amap[sg.map], &full[stage], sg.kcol + kk * kTcBK, row0 + sg.shift);
            tma_load_2d(sB + stage * b_bytes, wmap, &full[stage], kb * kTcBK, n0);
            if (++stage == nst) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16(kTcBM, bn, 0, 0);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kTcMaxBN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * kTcABytes);
          const uint32_t b_addr = smem_u32(sB + stage * b_bytes);
#pragma unroll
          for (int k = 0; k < kTcBK / 16; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * 32, 0, 1024, 2);
            const uint64_t db = umma_smem_desc(b_addr + k * 32, 0, 1024, 2);
            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int tm = tile % n_tiles_m, tn = tile / n_tiles_m;
      const int row = tm * kTcBM + quad * 32 + lane;
      const int n0 = tn * bn;
      // per-tile staging, off the TMEM critical path: row state in registers, column vector (bias) in smem
      const RowState rs = epi.prep(row);
      float* cv = cvec_s + acc * kTcMaxBN;
      {
        const float* gv = epi.colvec();
        const int et = threadIdx.x;       // 0..127: the epilogue warps are warps 0-3
        for (int i = et; i < bn; i += 128) cv[i] = (gv && n0 + i < p.N) ? gv[n0 + i] : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * kTcMaxBN + ((uint32_t)(quad * 32) << 16);
      // TMEM -> registers in 64-column groups, software pipelined: the tcgen05.ld of group g+1 is in flight while
      // group g runs through the epilogue functor (a lone ld + wait per 16 columns costs a full TMEM round trip
      // each time and made the epilogue, not the MMAs, the critical path of the K = 1024 GEMMs).
      uint32_t ra[4][16], rbuf[4][16];
      const int ngroups = (bn + 63) / 64;
      auto load_group = [&](uint32_t (&dst)[4][16], int g) {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (g * 64 + q * 16 < bn) tmem_ld16(t_addr + g * 64 + q * 16, dst[q]);
      };
      auto run_group = [&](uint32_t (&src)[4][16], int g) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = g * 64 + q * 16;
          if (c < bn && n0 + c < p.N) {
            float v[16], cvr[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(src[q][i]);
#pragma unroll
            for (int i = 0; i < 16; i += 4) {   // bias of this chunk: 4 x LDS.128 from the per-tile staging buffer
              const float4 t4 = *reinterpret_cast<const float4*>(cv + c + i);
              cvr[i] = t4.x; cvr[i + 1] = t4.y; cvr[i + 2] = t4.z; cvr[i + 3] = t4.w;
            }
            epi.template operator()<16>(row, n0 + c, v, rs, cvr);
          }
        }
      };
      load_group(ra, 0);
      for (int g = 0; g < ngroups; g += 2) {
        tmem_ld_wait();
        if (g + 1 < ngroups) load_group(rbuf, g + 1);
        run_group(ra, g);
        if (g + 1 < ngroups) {
          tmem_ld_wait();
          if (g + 2 < ngroups) load_group(ra, g + 2);
          run_group(rbuf, g + 1);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <typename Epi>
inline int launch_rowgemm_tc(const GemmDesc& d, const Epi& epi, cudaStream_t stream) {
  TcParams p{};
  p.nseg = d.nseg;
  p.N = d.N;
  p.rows_alloc = d.rows_alloc;
  p.plan = d.plan;
  if (d.N % 16) return RADTTS_ERR_INVALID_ARG;
  // candidate tile widths; a partial last column tile reads zero-filled weight rows and is masked on store
  if (d.N <= kTcMaxBN) {
    p.bn[0] = d.N;
    p.ncand = 1;
  } else {
    static const int kWidths[kTcCand] = {256, 208, 176, 144};
    p.ncand = g_gemm_tile_select ? kTcCand : 1;
    for (int c = 0; c < p.ncand; ++c) p.bn[c] = kWidths[c];
  }
  const void* bases[kTcMaxMaps];
  int lds[kTcMaxMaps];
  int nmaps = 0;
  int ktotal = 0;
  for (int s = 0; s < d.nseg; ++s) {
    const Seg& sg = d.seg[s];
    if (sg.klen % kTcBK) return RADTTS_ERR_INVALID_ARG;
    int m = -1;
    for (int i = 0; i < nmaps; ++i)
      if (bases[i] == sg.a && lds[i] == sg.lda) m = i;
    if (m < 0) {
      if (nmaps == kTcMaxMaps) return RADTTS_ERR_UNSUPPORTED;
      m = nmaps++;
      bases[m] = sg.a;
      lds[m] = sg.lda;
      RB_TRY(make_map_bf16(sg.a, sg.lda, sg.lda, d.rows_alloc, kTcBM, &p.amap[m]));
    }
    p.seg[s] = TcSeg{m, sg.shift, sg.kcol, sg.klen / kTcBK};
    ktotal += sg.klen;
  }
  for (int i = nmaps; i < kTcMaxMaps; ++i) p.amap[i] = p.amap[0];
  if (ktotal > d.ldw) return RADTTS_ERR_INVALID_ARG;
  int min_tiles_n = ceil_div(d.N, p.bn[0]), max_tiles_n = min_tiles_n;
  for (int c = 0; c < p.ncand; ++c) {
    RB_TRY(make_map_bf16(d.w, d.ldw, d.ldw, d.N, p.bn[c], &p.wmap[c]));
    const int tn = ceil_div(d.N, p.bn[c]);
    if (tn > max_tiles_n) max_tiles_n = tn;
  }
  for (int c = p.ncand; c < kTcCand; ++c) { p.wmap[c] = p.wmap[0]; p.bn[c] = p.bn[0]; }
  static bool configured = false;  // per Epi instantiation
  if (!configured) {
    RB_CUDA(cudaFuncSetAttribute(rowgemm_tc_kernel<Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    configured = true;
  }
  const int max_tiles = (d.rows_alloc / kTcBM) * max_tiles_n;
  const int grid = max_tiles < kNumSMs ? max_tiles : kNumSMs;
  rowgemm_tc_kernel<Epi><<<grid, kTcThreads, kTcSmemBytes, stream>>>(p, epi);
  return after_launch();
}

}  // namespace rb
