// Shared host-side helpers for the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/radtts_b200.h"

namespace rb {

extern long long g_launches;  // defined in capi.cu; counts kernel launches (bench.py `gpu_launches`)
extern int g_gemm_tile_select;  // defined in capi.cu; radtts_set_gemm_tile_select

inline int after_launch() {
  ++g_launches;
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? 0 : (int)e;
}

#define RB_CUDA(expr)                      \
  do {                                     \
    cudaError_t _e = (expr);               \
    if (_e != cudaSuccess) return (int)_e; \
  } while (0)

#define RB_TRY(expr)       \
  do {                     \
    int _rc = (expr);      \
    if (_rc) return _rc;   \
  } while (0)

constexpr int kSmemBudget = 232448;  // 227 KB opt-in maximum per CTA on sm_100
constexpr int kNumSMs = 148;

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }
template <typename T>
__host__ __device__ constexpr T round_up(T a, T b) { return ceil_div(a, b) * b; }

}  // namespace rb
