// Micro-benchmark for the MAS row recurrence: cycles per row of one warp for incremental variants of the loop body.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mas_micro tools/mas_micro.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ROWS 2048

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
               : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}

// variant bits: 1 shuffle, 2 ballot, 4 lane-0 bit store (branch), 8 lane-31 edge store, 16 row load from smem,
//               32 predicated (asm) stores instead of branches, 64 a second warp spins on an mbarrier try_wait,
//               128 row loop unrolled by 4 with the row==0 test hoisted
template <int V>
__global__ void __launch_bounds__(64) micro(float* out, long long* cycles, int T2) {
  extern __shared__ float sm[];
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  float* rows = sm;                       // [64][T2] ring
  uint32_t* bits = reinterpret_cast<uint32_t*>(sm + 64 * T2);   // [ROWS]
  float* edge = reinterpret_cast<float*>(bits + ROWS);          // [128]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * T2; i += blockDim.x) rows[i] = -1.f - (float)((i * 7919) % 13) * 0.25f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    stop = 0;
  }
  __syncthreads();
  if (warp == 1) {
    if ((V & 64) && lane == 0) {
      while (!stop) { if (try_wait(&bar, 0)) break; }
    }
    return;
  }
  const float nanv = __int_as_float(0x7fc00000);
  float v = lane == 0 ? 0.f : -INFINITY;
  float a = rows[lane];
  const long long t0 = clock64();
  for (int row = 1; row < ROWS; ++row) {
    float an = 0.f;
    if (V & 16) an = rows[(row & 63) * T2 + lane];
    float left = (V & 1) ? __shfl_up_sync(0xffffffffu, v, 1) : v * 0.5f;
    if (lane == 0) left = nanv;
    const bool diag = left >= v;
    v = __fadd_rn(a, diag ? left : v);
    if (V & 2) {
      const uint32_t w = __ballot_sync(0xffffffffu, diag);
      if (V & 4) {
        if (V & 32) {
          asm volatile("{ .reg .pred p; setp.eq.s32 p, %2, 0; @p st.shared.u32 [%0], %1; }" ::"r"(smem_u32(&bits[row])), "r"(w), "r"(lane) : "memory");
        } else if (lane == 0) bits[row] = w;
      }
    }
    if (V & 8) {
      if (V & 32) {
        asm volatile("{ .reg .pred p; setp.eq.s32 p, %2, 31; @p st.shared.f32 [%0], %1; }" ::"r"(smem_u32(&edge[row & 127])), "f"(v), "r"(lane) : "memory");
      } else if (lane == 31) edge[row & 127] = v;
    }
    if (V & 16) a = an;
  }
  const long long t1 = clock64();
  stop = 1;
  out[lane] = v + (float)bits[ROWS - 1] + edge[5];
  if (lane == 0) cycles[0] = t1 - t0;
}

// Chunked variant: the row loop runs R rows at a time between an mbarrier wait (already complete) and an arrive +
// release store, like the real kernel.  C bits: 1 try_wait per chunk, 2 arrive + release per chunk, 4 runtime
// modulo for the stage index, 8 clock64 reads per chunk
__device__ __forceinline__ float lds(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
template <int C>
__global__ void __launch_bounds__(64) micro_chunked(float* out, long long* cycles, int T2, int R, int stages) {
  extern __shared__ float sm[];
  __shared__ uint64_t full[32], empty[32];
  __shared__ int done[4];
  float* rows = sm;
  uint32_t* bits = reinterpret_cast<uint32_t*>(sm + 64 * T2);
  float* edge = reinterpret_cast<float*>(bits + ROWS);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * T2; i += blockDim.x) rows[i] = -1.f - (float)((i * 7919) % 13) * 0.25f;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < 32; ++i)
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[i])) : "memory");   // phase 0 complete
  }
  __syncthreads();
  if (warp == 1) return;
  const float nanv = __int_as_float(0x7fc00000);
  float v = lane == 0 ? 0.f : -INFINITY;
  const int is_l0 = lane == 0, is_l31 = lane == 31;
  long long acc = 0;
  const long long t0 = clock64();
  for (int c = 0; c < ROWS / R; ++c) {
    const int s = (C & 4) ? c % stages : (c & 31);
    long long q0 = 0;
    if (C & 8) q0 = clock64();
    if (C & 1) { while (!try_wait(&full[s], 0)) {} }
    const int r0 = c * R, r1 = r0 + R;
    const uint32_t st = smem_u32(rows + (size_t)(s & 3) * 16 * T2 + lane);
    float a = lds(st);
    int row = r0;
    if (row == 0) row = 1;
    for (; row < r1; ++row) {
      const float an = lds(st + (uint32_t)(min(row + 1, r1 - 1) - r0) * T2 * 4);
      float left = __shfl_up_sync(0xffffffffu, v, 1);
      if (lane == 0) left = nanv;
      const bool diag = left >= v;
      v = __fadd_rn(a, diag ? left : v);
      const uint32_t w = __ballot_sync(0xffffffffu, diag);
      asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; @p st.shared.u32 [%0], %1; }" ::"r"(smem_u32(&bits[row])), "r"(w), "r"(is_l0) : "memory");
      asm volatile("{ .reg .pred p; setp.ne.s32 p, %2, 0; @p st.shared.f32 [%0], %1; }" ::"r"(smem_u32(&edge[row & 127])), "f"(v), "r"(is_l31) : "memory");
      a = an;
    }
    if (C & 2) {
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
        asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(&done[0])), "r"(r1) : "memory");
      }
    }
    if (C & 8) acc += clock64() - q0;
  }
  const long long t1 = clock64();
  out[lane] = v + (float)bits[ROWS - 1] + edge[5] + (float)acc;
  if (lane == 0) cycles[0] = t1 - t0;
}

template <int C>
void run_chunked(const char* name, int R) {
  float* out; long long* cyc;
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  const int T2 = 32;
  const size_t smem = (64 * T2 + ROWS + 128) * 4;
  for (int i = 0; i < 3; ++i) micro_chunked<C><<<1, 64, smem>>>(out, cyc, T2, R, 32);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("chunked R=%2d %-33s C=%3d  %7.1f cycles/row  %7.1f cycles/chunk (%s)\n", R, name, C, (double)h / (ROWS - 1),
         (double)h / (ROWS / R), cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

template <int V>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 256); cudaMalloc(&cyc, 8);
  const int T2 = 32;
  const size_t smem = (64 * T2 + ROWS + 128) * 4;
  for (int i = 0; i < 3; ++i) micro<V><<<1, 64, smem>>>(out, cyc, T2);
  long long h = 0;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%-46s V=%3d  %7.1f cycles/row  (%s)\n", name, V, (double)h / (ROWS - 1), cudaGetErrorString(e));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("chain only (fsetp, fsel, fadd)");
  run<1>("+ shfl.up");
  run<3>("+ ballot");
  run<7>("+ lane0 bits store (branch)");
  run<15>("+ lane31 edge store (branch)");
  run<31>("+ row load LDS");
  run<31 + 32>("same, predicated stores");
  run<31 + 64>("branchy + spinning try_wait warp");
  run<31 + 32 + 64>("predicated + spinning try_wait warp");
  for (int R : {16, 8, 4}) {
    run_chunked<0>("loop only", R);
    run_chunked<1>("+ try_wait", R);
    run_chunked<3>("+ arrive/release", R);
    run_chunked<7>("+ runtime modulo", R);
    run_chunked<15>("+ clock64", R);
  }
  return 0;
}
