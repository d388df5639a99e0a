// tcgen05 engine for the weight-gradient contractions (see wgrad.cuh).  One persistent launch processes a
// whole LIST of problems (all taps of all WN layers of a flow: ~850 equal-cost tiles -> 5.7 waves on 148 SMs,
// no split-K, plain stores).
//
// K is the packed row axis, so both operands are MN-major: a TMA box of {64 channels, 64 rows} with 128-byte
// swizzle lands in smem as 64 rows x 128 B, which is exactly the canonical "MN-major, SW128" UMMA layout
// (8-row groups of 1024 B; the next 64-channel block 8 KB further):  LBO = 8192 (MN block stride),
// SBO = 1024 (8-row K-group stride); one K=16 MMA spans two K-groups, so the descriptor advances 2048 B per MMA.
#pragma once
#include <cstdlib>

#include "rowgemm_tc.cuh"
#include "wgrad.cuh"

namespace rb {

constexpr int kWgMaxMaps = 24, kWgMaxProbs = 40;
constexpr int kWgBM = 128, kWgBN = 256, kWgBK = 64, kWgStages = 4, kWgThreads = 192;
constexpr int kWgABytes = kWgBM * kWgBK * 2, kWgBBytes = kWgBN * kWgBK * 2;
constexpr int kWgSmemBytes = kWgStages * (kWgABytes + kWgBBytes) + 1024 + 256;

struct WgTcProb {
  int gmap, xmap, gcol, xcol, N, C, shift, tiles_c, tile_begin, so_c, atomic, pad_;
  long so_n;
  float* out;
};
struct WgTcParams {
  CUtensorMap maps[kWgMaxMaps];
  WgTcProb prob[kWgMaxProbs];
  int nprob, total_tiles, rows_alloc, pad_;
  const int* plan;
};

// The kernel has ONE definition in the library (flow_bwd.cu defines RB_WGRAD_TC_DEFINE before including this header);
// other translation units (convnet.cu) launch it through this declaration.
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgTcParams p);
#ifdef RB_WGRAD_TC_DEFINE
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const __grid_constant__ WgTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kWgStages * kWgABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kWgStages * (kWgABytes + kWgBBytes));
  uint64_t* full = bars;
  uint64_t* empty = bars + kWgStages;
  uint64_t* tfull = bars + 2 * kWgStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows_used = p.plan ? p.plan[0] : p.rows_alloc;
  const int kb_total = (rows_used + kWgBK - 1) / kWgBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 4); }
    mbar_fence_init();
  }
  if (warp == 5) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto find_prob = [&](int tile) {
    int q = 0;
    while (q + 1 < p.nprob && p.prob[q + 1].tile_begin <= tile) ++q;
    return q;
  };

  if (warp == 4) {   // producer / MMA issuer on the high warp ids (arbiter priority), epilogue on warps 0-3
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const WgTcProb& pr = p.prob[find_prob(tile)];
        const int lt = tile - pr.tile_begin;
        const int n0 = (lt / pr.tiles_c) * kWgBM, c0 = (lt % pr.tiles_c) * kWgBN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], kWgABytes + kWgBBytes);
          uint8_t* a = sA + stage * kWgABytes;
          uint8_t* b = sB + stage * kWgBBytes;
#pragma unroll
          for (int nb = 0; nb < kWgBM / 64; ++nb)
            tma_load_2d(a + nb * 8192, &p.maps[pr.gmap], &full[stage], pr.gcol + n0 + nb * 64, kb * kWgBK);
#pragma unroll
          for (int cb = 0; cb < kWgBN / 64; ++cb)
            tma_load_2d(b + cb * 8192, &p.maps[pr.xmap], &full[stage], pr.xcol + c0 + cb * 64, kb * kWgBK + pr.shift);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16(kWgBM, kWgBN, 1, 1);
      const uint32_t lbo = 8192u, sbo = 1024u;   // MN-major SW128: 8 KB between 64-channel blocks, 1 KB between 8-row groups
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kWgBN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * kWgABytes);
          const uint32_t b_addr = smem_u32(sB + stage * kWgBBytes);
#pragma unroll
          for (int k = 0; k < kWgBK / 16; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * 2048, lbo, sbo, 2);
            const uint64_t db = umma_smem_desc(b_addr + k * 2048, lbo, sbo, 2);
            umma_f16(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == kWgStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    int acc = 0; uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const WgTcProb& pr = p.prob[find_prob(tile)];
      const int lt = tile - pr.tile_begin;
      const int n = (lt / pr.tiles_c) * kWgBM + quad * 32 + lane;
      const int c0 = (lt % pr.tiles_c) * kWgBN;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + (uint32_t)acc * kWgBN + ((uint32_t)(quad * 32) << 16);
      float* orow = pr.out + (size_t)n * pr.so_n;
      for (int c = 0; c < kWgBN; c += 16) {
        uint32_t r[16];
        tmem_ld16(t_addr + c, r);
        tmem_ld_wait();
        if (n < pr.N && kb_total > 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int cc = c0 + c + i;
            if (cc < pr.C) {
              const float v = __uint_as_float(r[i]);
              if (pr.atomic) atomicAdd(orow + (size_t)cc * pr.so_c, v);
              else orow[(size_t)cc * pr.so_c] = v;
            }
          }
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

#endif  // RB_WGRAD_TC_DEFINE

// Host-side batch builder ----------------------------------------------------------------------------------
struct WgradBatch {
  WgTcParams params{};
  const void* bases[kWgMaxMaps];
  int lds[kWgMaxMaps];
  int nmaps = 0;
  int rows_alloc = 0;

  int map_for(const void* base, int ld) {
    for (int i = 0; i < nmaps; ++i)
      if (bases[i] == base && lds[i] == ld) return i;
    if (nmaps == kWgMaxMaps) return -1;
    if (make_map_bf16(base, ld, ld, rows_alloc, kWgBK, &params.maps[nmaps])) return -1;
    bases[nmaps] = base;
    lds[nmaps] = ld;
    return nmaps++;
  }
  // returns false if the problem cannot go on the tensor-core path
  bool add(const WgradProb& w, bool atomic) {
    if (params.nprob == kWgMaxProbs) return false;
    if ((w.ldg % 8) || (w.ldx % 8) || (w.gcol % 8) || (w.xcol % 8)) return false;
    const int gm = map_for(w.G, w.ldg);
    const int xm = map_for(w.X, w.ldx);
    if (gm < 0 || xm < 0) return false;
    WgTcProb& q = params.prob[params.nprob++];
    q.gmap = gm; q.xmap = xm; q.gcol = w.gcol; q.xcol = w.xcol; q.N = w.N; q.C = w.C; q.shift = w.shift;
    q.tiles_c = ceil_div(w.C, kWgBN);
    q.tile_begin = params.total_tiles;
    q.so_c = w.so_c; q.so_n = w.so_n; q.out = w.out; q.atomic = atomic ? 1 : 0;
    params.total_tiles += ceil_div(w.N, kWgBM) * q.tiles_c;
    return true;
  }
  int launch(const int* plan, cudaStream_t st) {
    if (params.nprob == 0) return 0;
    params.plan = plan;
    params.rows_alloc = rows_alloc;
    static bool configured = false;
    if (!configured) {
      RB_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
      configured = true;
    }
    const int grid = params.total_tiles < kNumSMs ? params.total_tiles : kNumSMs;
    wgrad_tc_kernel<<<grid, kWgThreads, kWgSmemBytes, st>>>(params);
    return after_launch();
  }
};

inline bool wgrad_tc_supported(const WgradProb& w) {
  return !((w.ldg % 8) || (w.ldx % 8) || (w.gcol % 8) || (w.xcol % 8));
}
inline int launch_wgrad_tc(const WgradProb& w, const int* plan, int rows_alloc, cudaStream_t st) {
  WgradBatch b;
  b.rows_alloc = rows_alloc;
  if (!b.add(w, true)) return RADTTS_ERR_UNSUPPORTED;
  return b.launch(plan, st);
}

}  // namespace rb
