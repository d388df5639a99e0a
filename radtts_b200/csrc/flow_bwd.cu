// Kernel 2, backward direction: everything PyTorch autograd derives for one reference FlowStep
// (SURVEY Appendix C), as explicit kernels on the packed frame layout.
//   1. coupling backward (elementwise)            -> g_params, g_zmid[z1]
//   2. `end` dgrad GEMM, epilogue emits g_u_i = g_R * softplus'(r_i) for all layers
//   3. per layer, ONE fused dgrad GEMM: g_x_{i+1} = g_u_i W_rs_i + convT(g_v_{i+1}); epilogue applies
//      softplus' and the partial-conv ratio -> g_v_i
//   4. convT(g_v_0) -> g_x0;  `start` dgrad -> g_ctx (+= over flows), g_zmid[z0] +=
//   5. 1x1-conv dgrad (fp32) -> g_zin
//   6. weight gradients: wgrad GEMMs (K = packed rows) + bias column sums
#include <cuda_bf16.h>

#define RB_WGRAD_TC_DEFINE
#include "flow_common.cuh"
#include "invconv.cuh"
#include "wgrad.cuh"

namespace rb {

template <typename T>
__device__ __forceinline__ void put_act(T* p, float v);
template <>
__device__ __forceinline__ void put_act<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void put_act<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }

// ----------------------------------------------------------------------------------------------------
// 1. coupling backward.  z1' = s z1 + b, log_s = log s, s = f(x):
//      g_z1 = s g_z1',  g_b = g_z1',  g_x = (g_z1' z1 + g_log_s / s) ds/dx
//    also seeds g_zmid with the pass-through columns (exited channels and z0).
// ----------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) coupling_bwd_kernel(const float* __restrict__ g_zout, const float* __restrict__ g_log_s,
                                                           const float* __restrict__ params, const float* __restrict__ zmid,
                                                           RowMeta meta, const int* __restrict__ plan, int zld, int c_off,
                                                           int h, int scaling, int kpad, float* __restrict__ g_zmid,
                                                           T* __restrict__ g_params) {
  const int rows_used = (plan[0] + 127) / 128 * 128;  // whole tiles: rows past the packed range get zeros
  const int per_row = kpad > zld ? kpad : zld;
  const size_t total = (size_t)rows_used * per_row;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % per_row);
    const int row = (int)(i / per_row);
    const bool ok = meta.valid(row);
    // pass-through part of g_zmid
    if (j < c_off + h) g_zmid[(size_t)row * zld + j] = ok ? g_zout[(size_t)row * zld + j] : 0.f;
    if (j < kpad) {
      float gp = 0.f;
      const int c = j >> 1, which = j & 1;
      if (!ok && c < h && !which) g_zmid[(size_t)row * zld + c_off + h + c] = 0.f;
      if (ok && c < h) {
        const size_t zi = (size_t)row * zld + c_off + h + c;
        const float gz = g_zout[zi];
        if (which) {
          gp = gz;  // g_b
        } else {
          const float x = params[(size_t)row * zld + 2 * c];
          const float gl = g_log_s ? g_log_s[(size_t)row * (zld / 2) + c] : 0.f;
          const float z1 = zmid[zi];
          float s, ds;
          if (scaling == 0) { const float t = tanhf(x); s = (t + 1.f) + 1e-6f; ds = 1.f - t * t; }
          else if (scaling == 1) { s = expf(x); ds = s; }
          else if (scaling == 2) { const float sg = 1.f / (1.f + expf(-(x + 10.f))); s = sg + 1e-6f; ds = sg * (1.f - sg); }
          else { s = 1.f; ds = 0.f; }
          // log_s = log(s) for tanh/sigmoid, = x for exp (d log_s / dx = 1), = 0 for translate
          const float dls = (scaling == 1) ? 1.f : (scaling == 3 ? 0.f : ds / s);
          gp = gz * z1 * ds + gl * dls;
          g_zmid[zi] = s * gz;
        }
      }
      put_act(g_params + (size_t)row * kpad + j, gp);
    }
  }
}

// g_u_l[r][c] = g_R[r][c] * softplus'(r_l[r][c]) for every WN layer l (softplus' through the saved OUTPUT r_l).
// Fully coalesced 16-byte accesses; replaces a per-row fan-out inside the GEMM epilogue that ran at ~0.7 TB/s.
template <typename T>
__global__ void __launch_bounds__(256) end_fanout_kernel(const T* __restrict__ gR, const T* __restrict__ r, int n_layers,
                                                         int n_ch, int rows_alloc, const int* __restrict__ plan,
                                                         T* __restrict__ gu) {
  constexpr int V = 16 / sizeof(T);                       // elements per 16-byte vector
  const int rows_used = (plan[0] + 127) / 128 * 128;
  const size_t nvec = (size_t)rows_used * n_ch / V;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i * V;
    const size_t row = e / n_ch;
    const int c = (int)(e % n_ch);
    float g[V];
    Act<T>::template ldv<V>(gR + e, g);
    for (int l = 0; l < n_layers; ++l) {
      float rv[V], y[V];
      Act<T>::template ldv<V>(r + row * ((size_t)n_layers * n_ch) + (size_t)l * n_ch + c, rv);
#pragma unroll
      for (int j = 0; j < V; ++j) y[j] = g[j] * softplus_grad_t<T>(rv[j]);
      Act<T>::template stv<V>(gu + ((size_t)l * rows_alloc + row) * n_ch + c, y);
    }
  }
}

// de-interleave the `end` weight gradient: tmp[2c + p][k] -> out[p * h + c][k]; tmp[..][n_ch] holds the bias grad
__global__ void end_grad_unpack_kernel(const float* __restrict__ tmp, int ldt, int h, int n_ch, float* __restrict__ gw,
                                       float* __restrict__ gb) {
  const int total = 2 * h * (n_ch + 1);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int k = i % (n_ch + 1), r = i / (n_ch + 1);
    const int c = r >> 1, p = r & 1;
    const float v = tmp[(size_t)r * ldt + k];
    if (k < n_ch) gw[(size_t)(p * h + c) * n_ch + k] = v;
    else gb[p * h + c] = v;
  }
}

template <typename T>
static int backward_impl(const radtts_flow_dims& d, const uint8_t* base, const PlanView& pv,
                         const radtts_flow_buffers& f, const radtts_flow_grad_buffers& g, int accumulate_ctx,
                         cudaStream_t st, cudaStream_t aux) {
  const int prec = sizeof(T) == 4 ? RADTTS_PREC_FP32 : RADTTS_PREC_BF16;
  FlowLayout L = flow_layout(d, prec, 1);
  const int h = d.c_active / 2, nc = d.n_ch, k = d.ksize, nl = d.n_layers;
  const int rows = pv.rows_alloc;
  RowMeta meta{pv.pos(), pv.rem()};
  const T* x = reinterpret_cast<const T*>(f.x);
  const T* r = reinterpret_cast<const T*>(f.r);
  T* gu = reinterpret_cast<T*>(g.g_u);
  T* gv = reinterpret_cast<T*>(g.g_v);
  T* gx0 = reinterpret_cast<T*>(g.g_x0);
  T* gparams = reinterpret_cast<T*>(g.g_params);

  // 1. coupling backward
  coupling_bwd_kernel<T><<<grid_for((size_t)rows * L.end_kpad), 256, 0, st>>>(
      g.g_zout, g.g_log_s, f.params, f.zmid, meta, pv.hdr(), d.z_ld, d.c_off, h, d.scaling, L.end_kpad, g.g_zmid, gparams);
  RB_TRY(after_launch());

  GemmDesc gd{};
  gd.rows_alloc = rows;
  gd.plan = pv.hdr();
  // 2. end dgrad
  gd.nseg = 1;
  gd.seg[0] = Seg{gparams, L.end_kpad, 0, 0, L.end_kpad};
  gd.w = base + L.w_end_t; gd.ldw = L.end_kpad; gd.N = nc;
  {
    // g_R lands in g_x0 (free until step 4), then one coalesced pass fans it out to g_u_0..L-1
    EpiDgradAct<T, ACT_NONE> e{nullptr, nc, gx0, nc, meta, ACT_NONE, 0, 0, k};
    RB_TRY((run_gemm<T>(gd, e, st)));
    end_fanout_kernel<T><<<grid_for((size_t)rows * nc / (16 / sizeof(T))), 256, 0, st>>>(gx0, r, nl, nc, rows, pv.hdr(), gu);
    RB_TRY(after_launch());
  }
  // 3. layers, last to first
  for (int i = nl - 1; i >= 0; --i) {
    gd.nseg = 1;
    gd.seg[0] = Seg{gu + (size_t)i * rows * nc, nc, 0, 0, nc};
    if (i + 1 < nl) {
      for (int t = 0; t < k; ++t)
        gd.seg[1 + t] = Seg{gv + (size_t)(i + 1) * rows * nc, nc, -((t - k / 2) << (i + 1)), 0, nc};
      gd.nseg = 1 + k;
    }
    gd.w = base + L.w_dg[i]; gd.ldw = L.dg_k[i]; gd.N = nc;
    EpiDgradAct<T, ACT_SOFTPLUS> e{x + (size_t)(i + 1) * rows * nc, nc, gv + (size_t)i * rows * nc, nc, meta, ACT_SOFTPLUS,
                                   d.partial_padding, i, k};
    RB_TRY((run_gemm<T>(gd, e, st)));
  }
  // 4. g_x0 and start dgrad
  gd.nseg = k;
  for (int t = 0; t < k; ++t) gd.seg[t] = Seg{gv, nc, -(t - k / 2), 0, nc};
  gd.w = base + L.w_dg0; gd.ldw = k * nc; gd.N = nc;
  {
    EpiDgradAct<T, ACT_NONE> e{nullptr, nc, gx0, nc, meta, ACT_NONE, 0, 0, k};
    RB_TRY((run_gemm<T>(gd, e, st)));
  }
  gd.nseg = 1;
  gd.seg[0] = Seg{gx0, nc, 0, 0, nc};
  gd.w = base + L.w_start_t; gd.ldw = nc; gd.N = L.ctx_ld + 128;
  {
    EpiStartDgrad e{g.g_ctx, d.n_ctx, L.ctx_ld, g.g_zmid, d.c_off, h, d.z_ld, accumulate_ctx, meta};
    RB_TRY((run_gemm<T>(gd, e, st)));
  }
  // 5. 1x1 conv dgrad (fp32 always)
  gd.nseg = 1;
  gd.seg[0] = Seg{g.g_zmid, d.z_ld, 0, 0, d.z_ld};
  gd.w = base + L.w_inv_t; gd.ldw = d.z_ld; gd.N = d.z_ld;
  if (d.z_ld % 4 == 0 && d.z_ld <= kInvMaxLd) {
    RB_TRY(launch_invconv_rows<float>(g.g_zmid, reinterpret_cast<const float*>(base + L.w_inv_t), d.z_ld, pv.hdr(), rows,
                                      meta, 1, g.g_zin, nullptr, 0, nullptr, 0, 0, st));
  } else {
    EpiStoreF32 e{g.g_zin, d.z_ld, meta, 1};
    RB_TRY(launch_rowgemm_simt(gd, e, st));
  }

  // 6. weight gradients -------------------------------------------------------------------------------
  // bf16: every problem of this flow goes into ONE batched tcgen05 launch; fp32: SIMT launches with atomics.
  // With an auxiliary stream (bf16 only) the memory-bound helpers -- the fp32 1x1-conv weight gradient, the bias
  // column sums and the `end` de-interleave -- leave the caller's stream: nothing on the dgrad chain of the NEXT flow
  // needs them, so they run underneath its tensor-bound GEMMs.  The caller joins `aux` before it reads any gradient.
  SideStream* sev = nullptr;
  cudaStream_t sa = st;
  if (aux && aux != st && sizeof(T) == 2) {
    sev = side_stream();
    if (!sev) return RADTTS_ERR_UNSUPPORTED;
    sa = aux;
  }
  float* tmp = g.scratch_f32;
  const int ldt = nc + 1;
  RB_CUDA(cudaMemsetAsync(tmp, 0, (size_t)d.z_ld * ldt * sizeof(float), st));
  if (sev) {
    RB_CUDA(cudaEventRecord(sev->fork[kSideForks - 1], st));
    RB_CUDA(cudaStreamWaitEvent(sa, sev->fork[kSideForks - 1], 0));
  }
  WgradBatch batch;
  batch.rows_alloc = rows;
  ColsumBatch<T> sums;
  // the tcgen05 wgrad stores every element of its output (no split-K, no atomics): only the SIMT path needs zeros
  constexpr bool kNeedZero = sizeof(T) != 2;
  auto wgrad = [&](const WgradProb& p, bool accumulate) -> int {
    if constexpr (sizeof(T) == 2) {
      if (batch.add(p, accumulate)) return 0;
      // batch full: SIMT fallback accumulates with atomics, so a non-accumulating problem needs its window zeroed
      if (!accumulate)
        RB_CUDA(cudaMemset2DAsync(p.out, (size_t)p.so_n * sizeof(float), 0, (size_t)p.C * sizeof(float), p.N, st));
    }
    return launch_wgrad_simt<T, T>(p, pv.hdr(), rows, st);
  };
  // 1x1 conv: dW_full[c][j] = sum_r g_zmid[r][c] zin[r][j]   (fp32 always)
  {
    RB_CUDA(cudaMemsetAsync(g.g_w_inv_full, 0, (size_t)d.z_ld * d.z_ld * sizeof(float), sa));
    if (d.z_ld == kInvMaxLd) {
      RB_TRY(launch_invconv_wgrad(g.g_zmid, f.zin, pv.hdr(), rows, g.g_w_inv_full, sa));
    } else {
      WgradProb p{g.g_zmid, d.z_ld, 0, d.z_ld, f.zin, d.z_ld, 0, d.z_ld, 0, g.g_w_inv_full, d.z_ld, 1};
      RB_TRY((launch_wgrad_simt<float, float>(p, pv.hdr(), rows, sa)));
    }
  }
  // end: interleaved rows into scratch, the n_layers K-blocks of r accumulate; bias = column sums of g_params
  for (int l = 0; l < nl; ++l) {
    WgradProb p{gparams, L.end_kpad, 0, d.z_ld, r, nl * nc, l * nc, nc, 0, tmp, ldt, 1};
    RB_TRY(wgrad(p, true));
  }
  RB_TRY(sums.add(gparams, L.end_kpad, 0, d.z_ld, 0, 0, tmp + nc, ldt, false, sa));
  for (int i = 0; i < nl; ++i) {
    const T* gui = gu + (size_t)i * rows * nc;
    const T* gvi = gv + (size_t)i * rows * nc;
    // res_skip_i: dW[n][c] = sum_r g_u_i[r][n] x_{i+1}[r][c]
    if (kNeedZero) RB_CUDA(cudaMemsetAsync(g.g_w_rs[i], 0, (size_t)nc * nc * sizeof(float), st));
    {
      WgradProb p{gui, nc, 0, nc, x + (size_t)(i + 1) * rows * nc, nc, 0, nc, 0, g.g_w_rs[i], nc, 1};
      RB_TRY(wgrad(p, false));
    }
    RB_TRY(sums.add(gui, nc, 0, nc, 0, 0, g.g_b_rs[i], 1, true, sa));
    // in_layer_i (tap-major output [t][n][c]): dW[t][n][c] = sum_r g_v_i[r][n] x_i[r + (t - half) d][c]
    if (kNeedZero) RB_CUDA(cudaMemsetAsync(g.g_w_in[i], 0, (size_t)nc * nc * k * sizeof(float), st));
    for (int t = 0; t < k; ++t) {
      WgradProb p{gvi, nc, 0, nc, x + (size_t)i * rows * nc, nc, 0, nc, (t - k / 2) << i,
                  g.g_w_in[i] + (size_t)t * nc * nc, nc, 1};
      RB_TRY(wgrad(p, false));
    }
    // bias sits outside the partial-conv renormalisation: g_bias = sum_r g_v / ratio
    RB_TRY(sums.add(gvi, nc, 0, nc, d.partial_padding, i, g.g_b_in[i], 1, true, sa));
  }
  // start: the two K segments land directly in the reference layout [z0 (h) | ctx (n_ctx)]
  {
    const int ldo = h + d.n_ctx;
    if (kNeedZero) RB_CUDA(cudaMemsetAsync(g.g_w_start, 0, (size_t)nc * ldo * sizeof(float), st));
    WgradProb pc{gx0, nc, 0, nc, f.ctx, L.ctx_ld, 0, d.n_ctx, 0, g.g_w_start + h, ldo, 1};
    RB_TRY(wgrad(pc, false));
    WgradProb pz{gx0, nc, 0, nc, f.z0, 128, 0, h, 0, g.g_w_start, ldo, 1};
    RB_TRY(wgrad(pz, false));
    RB_TRY(sums.add(gx0, nc, 0, nc, 0, 0, g.g_b_start, 1, true, sa));
  }
  RB_TRY(sums.launch(meta, k, pv.hdr(), rows, sa));
  RB_TRY(batch.launch(pv.hdr(), st));
  if (sev) {
    RB_CUDA(cudaEventRecord(sev->fork[kSideForks - 2], st));
    RB_CUDA(cudaStreamWaitEvent(sa, sev->fork[kSideForks - 2], 0));
  }
  end_grad_unpack_kernel<<<grid_for((size_t)2 * h * ldt), 256, 0, sa>>>(tmp, ldt, h, nc, g.g_w_end, g.g_b_end);
  RB_TRY(after_launch());
  return 0;
}

// ----------------------------------------------------------------------------------------------------
// 7. weight-norm backward, all convs of the flow in one launch (replaces 9 torch._weight_norm_interface_backward
//    launches + the tap-major -> (n, c, k) permute copies).  One CTA per (tensor, output channel).
// ----------------------------------------------------------------------------------------------------
constexpr int kWnMaxItems = 1 + 2 * RADTTS_MAX_LAYERS;
struct WnBwdItem {
  const float* v;    // weight_v (N, C, k) -- or the effective weight when g == NULL
  const float* g;    // weight_g (N) or NULL
  const float* gw;   // gradient of the effective weight: tap_major ? [k][N][C] : [N][C * k]
  float* gv;         // out (N, C, k)
  float* gg;         // out (N), unused when g == NULL
  int C, k, tap_major, pad_;
};
struct WnBwdParams {
  WnBwdItem it[kWnMaxItems];
  int N, accumulate;
};
__global__ void __launch_bounds__(256) wn_bwd_kernel(const WnBwdParams p) {
  extern __shared__ float gws[];                 // the channel's gradient row in (c, t) order
  __shared__ float red[2][8];
  const WnBwdItem& it = p.it[blockIdx.y];
  const int n = blockIdx.x, N = p.N, C = it.C, k = it.k, CK = C * k;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (it.tap_major) {
    for (int t = 0; t < k; ++t)
      for (int c = tid; c < C; c += 256) gws[c * k + t] = it.gw[((size_t)t * N + n) * C + c];
  } else {
    for (int i = tid; i < CK; i += 256) gws[i] = it.gw[(size_t)n * CK + i];
  }
  __syncthreads();
  float* gv = it.gv + (size_t)n * CK;
  if (!it.g) {                                   // not weight-normed: gradient of the weight itself
    for (int i = tid; i < CK; i += 256) gv[i] = p.accumulate ? gv[i] + gws[i] : gws[i];
    return;
  }
  const float* v = it.v + (size_t)n * CK;
  float dot = 0.f, nrm2 = 0.f;
  for (int i = tid; i < CK; i += 256) {
    const float x = v[i];
    dot = fmaf(x, gws[i], dot);
    nrm2 = fmaf(x, x, nrm2);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    dot += __shfl_xor_sync(0xffffffffu, dot, o);
    nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
  }
  if (lane == 0) { red[0][warp] = dot; red[1][warp] = nrm2; }
  __syncthreads();
  dot = 0.f; nrm2 = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) { dot += red[0][w]; nrm2 += red[1][w]; }
  const float norm = sqrtf(nrm2);
  const float scale = it.g[n] / norm;            // w = v * scale
  const float proj = dot / nrm2;
  for (int i = tid; i < CK; i += 256) {
    const float val = scale * (gws[i] - v[i] * proj);
    gv[i] = p.accumulate ? gv[i] + val : val;
  }
  if (tid == 0) {
    const float val = dot / norm;
    it.gg[n] = p.accumulate ? it.gg[n] + val : val;
  }
}

}  // namespace rb

using namespace rb;

extern "C" int radtts_flowstep_backward_ex(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                           int Tmax, const radtts_flow_buffers* fwd, const radtts_flow_grad_buffers* g,
                                           int accumulate_ctx, int precision, void* stream, void* aux_stream);
extern "C" int radtts_flowstep_backward(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                        int Tmax, const radtts_flow_buffers* fwd, const radtts_flow_grad_buffers* g,
                                        int accumulate_ctx, int precision, void* stream) {
  return radtts_flowstep_backward_ex(dims, prepared, plan, B, Tmax, fwd, g, accumulate_ctx, precision, stream, nullptr);
}

extern "C" int radtts_flowstep_backward_ex(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                           int Tmax, const radtts_flow_buffers* fwd, const radtts_flow_grad_buffers* g,
                                           int accumulate_ctx, int precision, void* stream, void* aux_stream) {
  RB_TRY(check_dims(dims));
  if (!prepared || !plan || !fwd || !g || B <= 0 || Tmax <= 0) return RADTTS_ERR_INVALID_ARG;
  if (!fwd->ctx || !fwd->zin || !fwd->zmid || !fwd->z0 || !fwd->x || !fwd->r || !fwd->params) return RADTTS_ERR_INVALID_ARG;
  if (!g->g_zout || !g->g_zin || !g->g_ctx || !g->g_zmid || !g->g_params || !g->g_u || !g->g_v || !g->g_x0 ||
      !g->g_w_inv_full || !g->g_w_start || !g->g_b_start || !g->g_w_end || !g->g_b_end || !g->scratch_f32)
    return RADTTS_ERR_INVALID_ARG;
  for (int i = 0; i < dims->n_layers; ++i)
    if (!g->g_w_in[i] || !g->g_b_in[i] || !g->g_w_rs[i] || !g->g_b_rs[i]) return RADTTS_ERR_INVALID_ARG;
  PlanView pv = make_plan_view(plan, B, Tmax);
  const uint8_t* base = reinterpret_cast<const uint8_t*>(prepared);
  if (precision == RADTTS_PREC_FP32)
    return backward_impl<float>(*dims, base, pv, *fwd, *g, accumulate_ctx, (cudaStream_t)stream, nullptr);
  if (precision == RADTTS_PREC_BF16)
    return backward_impl<__nv_bfloat16>(*dims, base, pv, *fwd, *g, accumulate_ctx, (cudaStream_t)stream,
                                        (cudaStream_t)aux_stream);
  return RADTTS_ERR_INVALID_ARG;
}

extern "C" int radtts_flow_weight_norm_backward(const radtts_flow_dims* dims, const radtts_flow_weights* w,
                                                const radtts_flow_grad_buffers* g, const radtts_flow_wn_grads* out,
                                                int accumulate, void* stream) {
  RB_TRY(check_dims(dims));
  if (!w || !g || !out) return RADTTS_ERR_INVALID_ARG;
  const int nl = dims->n_layers, nc = dims->n_ch, k = dims->ksize, h = dims->c_active / 2;
  WnBwdParams p{};
  p.N = nc;
  p.accumulate = accumulate ? 1 : 0;
  auto set = [&](int idx, const float* v, const float* wg, const float* gw, float* gv, float* gg, int C, int kk,
                 int tap_major) -> int {
    if (!v || !gw || !gv || (wg && !gg)) return RADTTS_ERR_INVALID_ARG;
    p.it[idx] = WnBwdItem{v, wg, gw, gv, gg, C, kk, tap_major, 0};
    return 0;
  };
  RB_TRY(set(0, w->w_start, w->wg_start, g->g_w_start, out->gv_start, out->gg_start, h + dims->n_ctx, 1, 0));
  for (int i = 0; i < nl; ++i) {
    RB_TRY(set(1 + i, w->w_in[i], w->wg_in[i], g->g_w_in[i], out->gv_in[i], out->gg_in[i], nc, k, 1));
    RB_TRY(set(1 + nl + i, w->w_rs[i], w->wg_rs[i], g->g_w_rs[i], out->gv_rs[i], out->gg_rs[i], nc, 1, 0));
  }
  const int maxck = nc * k > h + dims->n_ctx ? nc * k : h + dims->n_ctx;
  const size_t smem = (size_t)maxck * sizeof(float);
  if (smem > 48 * 1024) return RADTTS_ERR_UNSUPPORTED;
  wn_bwd_kernel<<<dim3(nc, 1 + 2 * nl), 256, smem, (cudaStream_t)stream>>>(p);
  return after_launch();
}
