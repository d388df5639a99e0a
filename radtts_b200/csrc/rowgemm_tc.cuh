// bf16 tcgen05 engine of the row GEMM (placeholder until the TMA/TMEM kernel lands in this file).
#pragma once
#include "rowgemm.cuh"

namespace rb {
template <typename Epi>
inline int launch_rowgemm_tc(const GemmDesc&, const Epi&, cudaStream_t) { return RADTTS_ERR_UNSUPPORTED; }
}  // namespace rb
