// Micro-benchmark of the tcgen05 row GEMM engine (radtts_b200/csrc/rowgemm_tc.cuh) outside the library: the SAME
// kernel template, launched with an explicit tile width and grid, on WN-shaped problems.  It answers the planning
// questions DESIGN.md 9.1 leaves open in one GPU call:
//   * time per tile as a function of the tile width BN (is a narrower tile proportionally cheaper, or is the MMA rate
//     operand-fetch bound below BN = 256?)  -> whether wave-balanced tile widths can remove the 2.3-waves-in-3 loss;
//   * time of exactly 1, 2, 3 full waves -> the in-tile rate with no quantisation;
//   * the K = 1024 GEMMs (res_skip), where the epilogue rather than the MMAs sets the tile time.
// Every configuration is checked on sampled outputs against a double-precision host reference.
//
// Build (no GPU needed):  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
//                              -o tools/_bin/gemm_micro tools/gemm_micro.cu
// Run on the GPU box:     tools/_bin/gemm_micro            (prints one line per configuration)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../radtts_b200/csrc/rowgemm_tc.cuh"

namespace rb {
long long g_launches = 0;   // the library defines these in capi.cu
int g_gemm_tile_select = 1;
int g_gemm_prefetch_a = 1;

// bias + softplus + bf16 store: the arithmetic of the in_layer epilogue without the frame-plan lookups
struct EpiMicro {
  __nv_bfloat16* out; int ldo;
  const float* bias;
  __device__ __forceinline__ RowState prep(int) const { return RowState{true, 1.f}; }
  __device__ __forceinline__ const float* colvec() const { return bias; }
  template <int W>
  __device__ __forceinline__ void operator()(int row, int col0, const float (&acc)[W], const RowState&,
                                             const float (&cv)[W]) const {
    float y[W];
#pragma unroll
    for (int i = 0; i < W; ++i) y[i] = softplus_t<__nv_bfloat16>(acc[i] + cv[i]);
    Act<__nv_bfloat16>::template stv<W>(out + (size_t)row * ldo + col0, y);
  }
};

// launch_rowgemm_tc with the two knobs exposed
static int launch_micro(const GemmDesc& d, const EpiMicro& epi, int bn, int grid, cudaStream_t st) {
  TcParams p{};
  p.nseg = d.nseg; p.N = d.N; p.rows_alloc = d.rows_alloc; p.plan = nullptr;
  p.bn[0] = bn; p.ncand = 1;                     // one candidate: the width under test (the ring depth follows from it)
  RB_TRY(make_map_bf16(d.seg[0].a, d.seg[0].lda, d.seg[0].lda, d.rows_alloc, kTcBM, &p.amap[0]));
  for (int i = 1; i < kTcMaxMaps; ++i) p.amap[i] = p.amap[0];
  for (int s = 0; s < d.nseg; ++s) p.seg[s] = TcSeg{0, d.seg[s].shift, d.seg[s].kcol, d.seg[s].klen / kTcBK};
  RB_TRY(make_map_bf16(d.w, d.ldw, d.ldw, d.N, bn, &p.wmap[0]));
  for (int c = 1; c < kTcCand; ++c) { p.wmap[c] = p.wmap[0]; p.bn[c] = bn; }
  static bool configured = false;
  if (!configured) {
    RB_CUDA(cudaFuncSetAttribute(rowgemm_tc_kernel<EpiMicro>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
    configured = true;
  }
  rowgemm_tc_kernel<EpiMicro><<<grid, kTcThreads, kTcSmemBytes, st>>>(p, epi);
  return after_launch();
}
}  // namespace rb

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

struct Problem {
  int rows, C, N, taps, dil;   // A [rows][C]; K = taps * C; out [rows][N]
};

int main(int argc, char** argv) {
  using namespace rb;
  const int reps = argc > 1 ? atoi(argv[1]) : 20;
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  // L2 flush buffer (larger than the 126 MB L2), written between timed launches
  const size_t flush_bytes = 256u << 20;
  void* flush; cudaMalloc(&flush, flush_bytes);

  std::vector<Problem> probs;
  // in_layer, cfg2-sized: 85 row tiles (the benchmarked batch), then exactly 1 / 2 / 3 waves at BN = 256 (N = 1024 -> 4
  // column tiles: 37, 74, 111 row tiles)
  for (int tiles_m : {85, 37, 74, 111}) probs.push_back({tiles_m * 128, 1024, 1024, 5, 2});
  probs.push_back({85 * 128, 1024, 1024, 1, 1});   // res_skip: K = 1024
  probs.push_back({85 * 128, 4096, 160, 1, 1});    // `end`: K = 4 x 1024 concatenated skip buffer, N = 160, one tile per CTA
  const int bns_all[] = {256, 240, 224, 208, 192, 176, 160, 144, 128, 96, 64};
  const int bns_quick[] = {256, 208, 160};
  const bool quick = argc > 2;
  std::vector<int> bns(quick ? std::begin(bns_quick) : std::begin(bns_all), quick ? std::end(bns_quick) : std::end(bns_all));

  for (const Problem& pr : probs) {
    const int K = pr.taps * pr.C;
    std::vector<__nv_bfloat16> hA((size_t)pr.rows * pr.C), hW((size_t)pr.N * K);
    std::vector<float> hb(pr.N);
    uint32_t s = 12345u + pr.rows;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.f - 0.5f; };
    for (auto& v : hA) v = __float2bfloat16(rnd());
    for (auto& v : hW) v = __float2bfloat16(rnd() * 0.05f);
    for (auto& v : hb) v = rnd();
    __nv_bfloat16 *dA, *dW, *dO; float* db;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dW, hW.size() * 2); cudaMalloc(&dO, (size_t)pr.rows * pr.N * 2);
    cudaMalloc(&db, pr.N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, hW.data(), hW.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), pr.N * 4, cudaMemcpyHostToDevice);
    GemmDesc d{};
    d.nseg = pr.taps; d.w = dW; d.ldw = K; d.N = pr.N; d.rows_alloc = pr.rows; d.plan = nullptr;
    for (int t = 0; t < pr.taps; ++t) d.seg[t] = Seg{dA, pr.C, (t - pr.taps / 2) * pr.dil, 0, pr.C};
    EpiMicro epi{dO, pr.N, db};
    const double flops = 2.0 * pr.rows * pr.N * K;

    for (int bn : bns) {
      if (pr.taps == 1 && bn < 128) continue;
      if (bn > pr.N) continue;
      const int tiles = (pr.rows / 128) * ceil_div(pr.N, bn);
      const int grid = tiles < kNumSMs ? tiles : kNumSMs;
      cudaMemsetAsync(dO, 0, (size_t)pr.rows * pr.N * 2, st);
      int rc = launch_micro(d, epi, bn, grid, st);
      if (rc || cudaStreamSynchronize(st) != cudaSuccess) {
        printf("rows %d K %d N %d bn %d: launch failed rc=%d (%s)\n", pr.rows, K, pr.N, bn, rc,
               cudaGetErrorString(cudaGetLastError()));
        return 1;
      }
      // sampled check against a double-precision reference of the same bf16 inputs
      std::vector<__nv_bfloat16> hO((size_t)pr.rows * pr.N);
      cudaMemcpy(hO.data(), dO, hO.size() * 2, cudaMemcpyDeviceToHost);
      double max_err = 0;
      for (int i = 0; i < 96; ++i) {
        const int r = i < 8 ? i : (i < 16 ? pr.rows - 1 - (i - 8) : (int)((i * 2654435761u) % (uint32_t)pr.rows));
        const int c = i < 4 ? pr.N - 1 - i : (int)((i * 40503u + 7u) % (uint32_t)pr.N);
        double acc = hb[c];
        for (int t = 0; t < pr.taps; ++t) {
          const int rr = r + (t - pr.taps / 2) * pr.dil;
          if (rr < 0 || rr >= pr.rows) continue;
          for (int k = 0; k < pr.C; ++k)
            acc += (double)__bfloat162float(hA[(size_t)rr * pr.C + k]) *
                   (double)__bfloat162float(hW[(size_t)c * K + t * pr.C + k]);
        }
        const double want = acc > 20 ? acc : log1p(exp(acc));
        const double got = __bfloat162float(hO[(size_t)r * pr.N + c]);
        const double err = fabs(got - want) / (fabs(want) + 1e-2);
        if (err > max_err) max_err = err;
      }
      float total = 0;
      for (int it = 0; it < reps + 3; ++it) {
        cudaMemsetAsync(flush, it, flush_bytes, st);
        cudaEventRecord(e0, st);
        launch_micro(d, epi, bn, grid, st);
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 3) total += ms;
      }
      const double us = total / reps * 1e3;
      const int rounds = ceil_div(tiles, grid);
      printf("rows %6d K %5d N %4d | bn %3d tiles %4d grid %3d rounds %d | %8.2f us  %7.1f TFLOP/s | us/round %6.2f "
             "us/round/256col %6.2f | max rel err %.4f %s\n",
             pr.rows, K, pr.N, bn, tiles, grid, rounds, us, flops / us * 1e-6, us / rounds, us / rounds * 256.0 / bn,
             max_err, max_err < 2e-2 ? "ok" : "MISMATCH");
      fflush(stdout);
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dO); cudaFree(db);
  }
  // ---- two independent row ranges as two launch chains on two streams: the partial last wave of one
  //      chain's GEMM is back-filled by the other chain's next GEMM.  Compared with one chain over all rows.
  {
    const int C = 1024, N = 1024, taps = 5, K = taps * C, L = 8;
    const int tm_all = 85, tm_a = 43, tm_b = 42;
    __nv_bfloat16 *dA[3], *dO[3], *dW; float* db;
    const int tms[3] = {tm_all, tm_a, tm_b};
    cudaMalloc(&dW, (size_t)N * K * 2); cudaMemset(dW, 0, (size_t)N * K * 2);
    cudaMalloc(&db, N * 4); cudaMemset(db, 0, N * 4);
    for (int i = 0; i < 3; ++i) {
      cudaMalloc(&dA[i], (size_t)tms[i] * 128 * C * 2); cudaMemset(dA[i], 0, (size_t)tms[i] * 128 * C * 2);
      cudaMalloc(&dO[i], (size_t)tms[i] * 128 * N * 2);
    }
    cudaStream_t st2; cudaStreamCreate(&st2);
    cudaEvent_t fork, join; cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&join, cudaEventDisableTiming);
    auto desc = [&](int i) {
      GemmDesc d{};
      d.nseg = taps; d.w = dW; d.ldw = K; d.N = N; d.rows_alloc = tms[i] * 128; d.plan = nullptr;
      for (int t = 0; t < taps; ++t) d.seg[t] = Seg{dA[i], C, (t - taps / 2) * 2, 0, C};
      return d;
    };
    const GemmDesc d_all = desc(0), d_a = desc(1), d_b = desc(2);
    auto grid_of = [&](int tm) { const int t = tm * 4; return t < kNumSMs ? t : kNumSMs; };
    for (int mode = 0; mode < 2; ++mode) {
      float total = 0;
      for (int it = 0; it < reps + 3; ++it) {
        cudaMemsetAsync(flush, it, flush_bytes, st);
        cudaEventRecord(e0, st);
        if (mode == 0) {
          for (int l = 0; l < L; ++l) launch_micro(d_all, EpiMicro{dO[0], N, db}, 256, grid_of(tm_all), st);
        } else {
          cudaEventRecord(fork, st);
          cudaStreamWaitEvent(st2, fork, 0);
          for (int l = 0; l < L; ++l) {
            launch_micro(d_a, EpiMicro{dO[1], N, db}, 256, grid_of(tm_a), st);
            launch_micro(d_b, EpiMicro{dO[2], N, db}, 256, grid_of(tm_b), st2);
          }
          cudaEventRecord(join, st2);
          cudaStreamWaitEvent(st, join, 0);
        }
        cudaEventRecord(e1, st);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (it >= 3) total += ms;
      }
      const double us = total / reps * 1e3 / L;
      printf("%s: %d x in_layer GEMM over 85 row tiles: %8.2f us per layer  %7.1f TFLOP/s\n",
             mode == 0 ? "one chain  (1 stream )" : "two chains (2 streams)", L, us,
             2.0 * tm_all * 128 * N * K / us * 1e-6);
    }
  }
  (void)bf16_round;
  return 0;
}
