// Feasibility test: can a tcgen05.mma A operand start at an arbitrary ROW of a K-major, 128-byte-swizzled shared-memory
// window?  (Conv taps are row-shifted views of the same activation tile: if they can share one TMA window, the A-operand
// traffic of the 5-tap GEMMs falls 4-5x.)  One CTA: TMA-loads a 160-row x 64-column bf16 window and a 64 x 64 weight tile,
// runs D[128 x 64] = A[s : s + 128, :] * B^T for several row shifts s and two settings of the descriptor's
// "matrix base offset" field, and compares with a host reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -I radtts_b200/csrc \
//        -o tools/_bin/umma_rowshift_test tools/umma_rowshift_test.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../radtts_b200/csrc/rowgemm_tc.cuh"

namespace rb {
long long g_launches = 0;
int g_gemm_tile_select = 1;

__device__ __forceinline__ uint64_t desc_bo(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;   // 128B swizzle
  return d;
}

constexpr int kWinRows = 160, kN = 64;

__global__ void __launch_bounds__(128, 1) rowshift_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap,
                                                          int shift, int use_bo, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = sm;                         // 160 rows x 128 B = 20 KB
  uint8_t* sB = sm + 20480;                 // 64 rows x 128 B = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 20480 + 8192);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<64>(slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar[0], kWinRows * 128 + kN * 128);
    tma_load_2d(sA, &amap, &bar[0], 0, 0);
    tma_load_2d(sB, &bmap, &bar[0], 0, 0);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_bf16(128, kN, 0, 0);
    const uint32_t a0 = smem_u32(sA) + (uint32_t)shift * 128u, b0 = smem_u32(sB);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = desc_bo(a0 + k * 32, 1024, use_bo ? ((a0 >> 7) & 7) : 0);
      const uint64_t db = desc_bo(b0 + k * 32, 1024, 0);
      umma_f16(tmem, da, db, idesc, k != 0);
    }
    umma_commit(&bar[1]);
  }
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  for (int c = 0; c < kN; c += 16) {
    uint32_t r[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(size_t)(warp * 32 + lane) * kN + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem); }
}
}  // namespace rb

int main() {
  using namespace rb;
  std::vector<__nv_bfloat16> hA((size_t)kWinRows * 64), hB((size_t)kN * 64);
  uint32_t s = 777u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((s >> 8) & 0xffff) / 65536.f - 0.5f; };
  for (auto& v : hA) v = __float2bfloat16(rnd());
  for (auto& v : hB) v = __float2bfloat16(rnd());
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * kN * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap am, bm;
  if (make_map_bf16(dA, 64, 64, kWinRows, kWinRows, &am) || make_map_bf16(dB, 64, 64, kN, kN, &bm)) { printf("map failed\n"); return 1; }
  const int smem = 20480 + 8192 + 64 + 1024;
  cudaFuncSetAttribute(rowshift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hO(128 * kN);
  for (int use_bo = 0; use_bo < 2; ++use_bo)
    for (int shift : {0, 1, 2, 3, 4, 6, 8, 12, 16, 20, 31, 32}) {
      cudaMemset(dO, 0, 128 * kN * 4);
      rowshift_kernel<<<1, 128, smem>>>(am, bm, shift, use_bo, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("shift %d bo %d: %s\n", shift, use_bo, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      double max_err = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < kN; ++n) {
          double acc = 0;
          for (int k = 0; k < 64; ++k)
            acc += (double)__bfloat162float(hA[(size_t)(m + shift) * 64 + k]) * (double)__bfloat162float(hB[(size_t)n * 64 + k]);
          const double err = fabs(acc - hO[(size_t)m * kN + n]);
          if (err > max_err) max_err = err;
        }
      printf("row shift %2d  base_offset %s : max abs err %.3e  %s\n", shift, use_bo ? "(addr>>7)&7" : "0          ", max_err,
             max_err < 1e-3 ? "OK" : "WRONG");
    }
  return 0;
}
