// Helpers shared by the forward (flow.cu) and backward (flow_bwd.cu) translation units of kernel 2.
#pragma once
#include "common.cuh"
#include "flow_layout.cuh"
#include "frameplan.cuh"
#include "rowgemm.cuh"
#include "rowgemm_tc.cuh"

namespace rb {

inline int grid_for(size_t n, int threads = 256) {
  size_t b = (n + threads - 1) / threads;
  if (b > (size_t)kNumSMs * 16) b = (size_t)kNumSMs * 16;
  if (b < 1) b = 1;
  return (int)b;
}

// fp32 activations -> SIMT engine, bf16 activations -> tcgen05 engine
template <typename T, typename Epi>
inline int run_gemm(const GemmDesc& g, const Epi& e, cudaStream_t st) {
  if constexpr (sizeof(T) == 4) return launch_rowgemm_simt(g, e, st);
  else return launch_rowgemm_tc(g, e, st);
}

inline int check_dims(const radtts_flow_dims* d) {
  if (!d) return RADTTS_ERR_INVALID_ARG;
  if (d->z_ld <= 0 || d->z_ld % 16 || d->c_active <= 0 || d->c_active % 2 || d->c_off + d->c_active != d->z_ld)
    return RADTTS_ERR_INVALID_ARG;
  if (d->c_active / 2 > 128 || d->c_off + 128 > d->z_ld + 96) return RADTTS_ERR_UNSUPPORTED;
  if (d->n_ch <= 0 || d->n_ch % 64 || d->n_layers < 1 || d->n_layers > RADTTS_MAX_LAYERS) return RADTTS_ERR_UNSUPPORTED;
  if (d->ksize < 1 || d->ksize % 2 == 0 || d->ksize > kMaxSeg - 1 || d->n_ctx <= 0) return RADTTS_ERR_UNSUPPORTED;
  if (((d->ksize / 2) << (d->n_layers - 1)) > kGap) return RADTTS_ERR_UNSUPPORTED;
  return 0;
}

// One internal side stream (+ fork / join events) per process, created on first use -- i.e. during the eager warm-up
// that precedes any stream capture.
constexpr int kSideForks = RADTTS_MAX_LAYERS + 4;
struct SideStream {
  cudaStream_t stream;
  cudaEvent_t fork[kSideForks];
  cudaEvent_t join;
};
inline SideStream* side_stream() {
  static SideStream s{};
  static int state = 0;   // 0 = not created, 1 = ok, -1 = failed
  if (state == 0) {
    bool ok = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; ok && i < kSideForks; ++i) ok = cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) == cudaSuccess;
    state = ok ? 1 : -1;
  }
  return state == 1 ? &s : nullptr;
}


}  // namespace rb
