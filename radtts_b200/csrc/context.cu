// Hard-attention context (reference radtts.py:393-399): with a binarized alignment,
//     context = bmm(text_enc, attn_hard^T)
// multiplies by a one-hot matrix -- a gather of text-encoder columns by the frame -> token index that kernel 1 (MAS)
// already produces, and its backward is a segment sum over each token's (contiguous) frames.  HBM-bound: the forward
// writes B*C*T1 floats once, the backward reads them once.
//   * frames t >= out_len carry frame_to_token = -1 and get zeros (their rows of the hard map are empty);
//   * the reference sets opt[0, 0] = 1 unconditionally (alignment.py:59): when the path does not start on token 0,
//     frame 0 has TWO ones and its context is text[:, 0] + text[:, path[0]] -- reproduced here.
#include "common.cuh"

namespace rb {

__global__ void __launch_bounds__(256) context_gather_kernel(const float* __restrict__ text, const int* __restrict__ f2t,
                                                             int C, int T1, int T2, float* __restrict__ ctx) {
  const int b = blockIdx.z, c = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T1) return;
  const float* trow = text + ((size_t)b * C + c) * T2;
  const int j = f2t[(size_t)b * T1 + t];
  float v = 0.f;
  if (j >= 0) {
    v = trow[j];
    if (t == 0 && j != 0) v += trow[0];
  }
  ctx[((size_t)b * C + c) * T1 + t] = v;
}

// g_text[b][c][j] = sum over the frames of token j of g_ctx[b][c][t]   (+ g_ctx[b][c][0] for j == 0, see above)
// One CTA per (b, 8 channels): token boundaries by binary search in the monotone frame_to_token row (staged in smem).
__global__ void __launch_bounds__(256) context_scatter_kernel(const float* __restrict__ g_ctx, const int* __restrict__ f2t,
                                                              int C, int T1, int T2, float* __restrict__ g_text) {
  extern __shared__ int f2t_s[];   // [T1]
  const int b = blockIdx.y, c0 = blockIdx.x * 8;
  for (int t = threadIdx.x; t < T1; t += blockDim.x) f2t_s[t] = f2t[(size_t)b * T1 + t];
  __syncthreads();
  // valid frames = prefix with f2t >= 0
  int lo = 0, hi = T1;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (f2t_s[mid] >= 0) lo = mid + 1; else hi = mid; }
  const int olen = lo;
  for (int idx = threadIdx.x; idx < 8 * T2; idx += blockDim.x) {
    const int cl = idx / T2, j = idx - cl * T2;
    const int c = c0 + cl;
    if (c >= C) continue;
    // first frame with f2t >= j, first frame with f2t >= j + 1
    int a0 = 0, a1 = olen;
    while (a0 < a1) { const int mid = (a0 + a1) >> 1; if (f2t_s[mid] < j) a0 = mid + 1; else a1 = mid; }
    int e0 = a0, e1 = olen;
    while (e0 < e1) { const int mid = (e0 + e1) >> 1; if (f2t_s[mid] < j + 1) e0 = mid + 1; else e1 = mid; }
    const float* grow = g_ctx + ((size_t)b * C + c) * T1;
    float s = 0.f;
    for (int t = a0; t < e0; ++t) s += grow[t];
    if (j == 0 && olen > 0 && f2t_s[0] != 0) s += grow[0];
    g_text[((size_t)b * C + c) * T2 + j] = s;
  }
}

}  // namespace rb

using namespace rb;

extern "C" int radtts_context_gather(const float* text_enc, const int32_t* frame_to_token, int B, int C, int T1, int T2,
                                     float* context, void* stream) {
  if (!text_enc || !frame_to_token || !context || B <= 0 || C <= 0 || T1 <= 0 || T2 <= 0) return RADTTS_ERR_INVALID_ARG;
  if (C > 65535 || B > 65535) return RADTTS_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(T1, 256), C, B);
  context_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(text_enc, frame_to_token, C, T1, T2, context);
  return after_launch();
}

extern "C" int radtts_context_scatter(const float* grad_context, const int32_t* frame_to_token, int B, int C, int T1,
                                      int T2, float* grad_text_enc, void* stream) {
  if (!grad_context || !frame_to_token || !grad_text_enc || B <= 0 || C <= 0 || T1 <= 0 || T2 <= 0)
    return RADTTS_ERR_INVALID_ARG;
  if (B > 65535 || (size_t)T1 * sizeof(int) > 48 * 1024) return RADTTS_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(C, 8), B);
  context_scatter_kernel<<<grid, 256, (size_t)T1 * sizeof(int), (cudaStream_t)stream>>>(grad_context, frame_to_token, C, T1,
                                                                                        T2, grad_text_enc);
  return after_launch();
}
