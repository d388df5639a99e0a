/*
 * radtts_b200 -- C ABI of the B200-native (sm_100a) RADTTS hot path.
 *
 * The reference (duj12/radtts) has no FFI: its boundary is the Python module API
 * (RADTTS.forward / RADTTS.infer and the layers they call).  Each entry point below replaces the
 * compute of the reference function cited next to it; the Python mirror in radtts_b200/*.py keeps
 * the reference's names and argument meaning and calls these through ctypes.
 *
 * Conventions
 *   - plain C, no C++/torch types; every pointer is a DEVICE pointer owned by the caller unless the
 *     name ends in _host; the library never allocates or frees device memory (workspace is passed in);
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - all launches are asynchronous on `stream`; return value 0 = ok, >0 = cudaError_t of the failing
 *     runtime call, <0 = RADTTS_ERR_* below;  no exceptions, no exit();
 *   - process-global state the library keeps: cached cudaFuncSetAttribute / driver entry points, a launch counter, ONE
 *     internal side stream + two events per device (res_skip back-fill fork/join of the flow step, captured like any
 *     other fork when the caller's stream is capturing), a cache of TMA tensor maps keyed by (pointer, shape), and the
 *     row-GEMM tuning switch (radtts_set_gemm_tile_select).
 *     Calls on different streams are safe; two threads calling flow-step entry points concurrently are not.
 */
#ifndef RADTTS_B200_H_
#define RADTTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RADTTS_API __attribute__((visibility("default")))
#else
#define RADTTS_API
#endif

#define RADTTS_ERR_INVALID_ARG (-1)
#define RADTTS_ERR_UNSUPPORTED (-2)
#define RADTTS_ERR_WORKSPACE (-3)
#define RADTTS_ERR_NOT_SM100 (-4)

/* ABI version of this header; bumped whenever a signature changes. */
RADTTS_API int radtts_abi_version(void);
/* Number of kernels launched by this library since load (all entry points); used by bench.py to
 * report `gpu_launches`. */
RADTTS_API long long radtts_launch_count(void);
/* Human-readable text for a negative RADTTS_ERR_* or a positive cudaError_t. */
RADTTS_API const char* radtts_error_string(int code);
/* Tuning switch of the tcgen05 row GEMM (csrc/rowgemm_tc.cuh): let the kernel pick, on the device, the tile width (256 /
 * 208 / 176 / 144 output columns) that needs the fewest rounds x width for the number of row tiles the frame plan holds
 * (default on; environment RADTTS_GEMM_TILE_SELECT=0 turns it off at load: always 256).  Results are bit-identical
 * either way.  Returns the previous setting. */
RADTTS_API int radtts_set_gemm_tile_select(int enabled);

/* ------------------------------------------------------------------------------------------------
 * Kernel 1 -- Monotonic Alignment Search + binarize_attention.
 * Replaces: RADTTS.binarize_attention (reference radtts.py:320-334) and alignment.mas_width1
 * (reference alignment.py:31-59), i.e. the D2H copy, the serial Numba DP and the per-utterance H2D.
 *
 *   attn            (B,1,T1,T2) float32 dense.  is_prob=1: probabilities (attn_soft), the log is taken
 *                   on the device; is_prob=0: log-probabilities (the bit-exact parity boundary).
 *   in_lens,out_lens (B) int64: text / mel lengths; utterance b is cropped to [:out_len,:in_len].
 *   attn_hard       (B,1,T1,T2) float32, fully written: exactly {0,1}, zero outside the crop.
 *   frame_to_token  optional (B,T1) int32: token index of every frame (-1 beyond out_len).
 *   durations       optional (B,T2) int32: == attn_hard.sum over T1 (includes the reference's extra
 *                   opt[0,0]=1 cell, alignment.py:59).
 *   ws              device workspace of at least radtts_mas_workspace_bytes(...) bytes.
 * ---------------------------------------------------------------------------------------------- */
RADTTS_API size_t radtts_mas_workspace_bytes(int B, int T1, int T2, int is_prob);
RADTTS_API int radtts_mas_forward(const float* attn, int is_prob, const int64_t* in_lens, const int64_t* out_lens, int B,
                       int T1, int T2, float* attn_hard, int32_t* frame_to_token, int32_t* durations, void* ws,
                       size_t ws_bytes, void* stream);


/* Diagnostic: phase marks of CTA 0 of the last radtts_mas_forward launch, copied to SIXTEEN host words:
 * globaltimer (ns): [0] start, [1] first DP warp done, [2] DP + fill done, [3] backtrack done, [4] end, [5] fill done;
 * clock64: [6] start, [7] first DP warp done; last DP warp, cycle sums over its chunks: [8] waiting for data, [9] row
 * loop, [11] chunks, [12] rows per chunk, [13] row-loop cycles of warp 0, [14] hand-off spins.  Synchronises the device. */
RADTTS_API int radtts_mas_debug_timeline(unsigned long long* out16_host);

/* ------------------------------------------------------------------------------------------------
 * Packed frame layout ("frame plan").
 * The reference keeps zero-padded (B, C, T') activations and re-derives masks with a host sync in every
 * layer (get_mask_from_lengths, reference common.py:86-97, called from common.py:567 etc.).  Here the
 * valid frames of all utterances are packed channels-last into one row axis with 16 zero rows between
 * utterances; the plan (row offsets, per-row position metadata) is built on the device from the lengths.
 *   lens     (B) int64 device; VALID frames per utterance = lens[b] / divisor (clamped to [0, Tmax]).
 *   geom_lens NULL, or (B) int64: rows physically reserved per utterance (>= valid).  Needed only for plain
 *            (non-partial) ConvNorm stacks, whose first layer reads the padded region too (common.py:145-154).
 *   plan     device buffer of radtts_frameplan_bytes(B, Tmax) bytes.
 *   rows     radtts_frameplan_rows(B, Tmax): allocated rows (multiple of 128) of every packed buffer.
 * ---------------------------------------------------------------------------------------------- */
RADTTS_API size_t radtts_frameplan_bytes(int B, int Tmax);
RADTTS_API int radtts_frameplan_rows(int B, int Tmax);
RADTTS_API int radtts_frameplan_build(const int64_t* lens, const int64_t* geom_lens, int divisor, int B, int Tmax,
                                      void* plan, void* stream);

/* (B, C, T) float32  <->  packed rows.  Implements the reference's squeeze/unsqueeze
 * (nn.Unfold((g,1), stride g) reference radtts.py:165-169,414 and fold radtts.py:308-318) as pure indexing:
 * packed[row0[b] + t'][col_off + c*g + k] = x[b][c][g*t' + k].
 *   pack:   dst is float32 (dst_bf16 = 0) or bfloat16 (dst_bf16 = 1), row stride ld elements; columns
 *           [col_off, col_off + ncols_pad) are written (zeros beyond C*g and on gap rows; with valid_only = 1 also
 *           on rows of the geometric span that lie past the valid length, i.e. the `x * mask` of a partial conv).
 *   unpack: src float32 packed, dst (B, C, Tmax*g) float32, zeros beyond each utterance's length. */
RADTTS_API int radtts_pack_frames(const float* src, int B, int C, int T, int g, const void* plan, int Tmax, void* dst,
                                  int dst_bf16, int ld, int col_off, int ncols_pad, int valid_only, void* stream);
RADTTS_API int radtts_unpack_frames(const float* src, int ld, int col_off, const void* plan, int B, int Tmax, int C,
                                    int g, float* dst, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Kernel 2 -- decoder flow step: Invertible1x1ConvLUS + AffineTransformationLayer(WN) + log-det pieces.
 * Replaces: FlowStep.forward (reference radtts.py:51-59), Invertible1x1ConvLUS.forward (common.py:407-428,
 * the 1x1 conv itself; W = P L U and log|det| stay a few tiny host-side torch ops), WN.forward
 * (common.py:560-578) with ConvNorm/PartialConv1d (common.py:145-154, partialconv1d.py:35-71) and
 * AffineTransformationLayer.forward with scaling_fn 'tanh' (common.py:782-784,810-832).
 *
 * precision: RADTTS_PREC_FP32 -- fp32 activations/weights, fp32 FMA (parity path, rtol 1e-3 bar);
 *            RADTTS_PREC_BF16 -- bf16 activations/weights on tcgen05 tensor cores, fp32 accumulate; z, the
 *                                1x1 conv, the coupling and log_s stay fp32 as in the reference under autocast.
 * ---------------------------------------------------------------------------------------------- */
#define RADTTS_PREC_FP32 0
#define RADTTS_PREC_BF16 1
#define RADTTS_PREC_BF16X2 2  /* LSTM recurrence only: operands split into hi + lo bf16 (16 mantissa bits), fp32 accumulate */
#define RADTTS_MAX_LAYERS 8

typedef struct radtts_flow_dims {
  int z_ld;      /* row width of the packed z buffers (n_mel_channels * n_group_size = 160) */
  int c_off;     /* first active column of this flow (2 * number of early exits so far) */
  int c_active;  /* channels this flow transforms (z_ld - c_off) */
  int n_ctx;     /* conditioning channels (1040) */
  int n_ch;      /* WN channels (1024) */
  int n_layers;  /* WN layers (4) */
  int ksize;     /* WN kernel size (5) */
  int partial_padding; /* 1: PartialConv1d renormalisation (decoder_use_partial_padding, default) */
  int scaling;   /* coupling scale fn: 0 tanh, 1 exp, 2 sigmoid, 3 translate (common.py:775-787) */
} radtts_flow_dims;

/* float32 weights in the reference's PyTorch layouts.  A weight-normed conv (torch.nn.utils.weight_norm, dim 0:
 * w = g * v / ||v||, reference common.py:540-556) is passed either as its effective weight (wg_* = NULL) or as
 * weight_v in w_* plus weight_g in wg_*: the prepare kernels then apply g / ||v|| while re-laying the weight out. */
typedef struct radtts_flow_weights {
  const float* w_inv;   /* (c_active, c_active): W for forward, W^-1 for inverse */
  const float* w_start; /* (n_ch, h + n_ctx), input channel order [z0 | context] (common.py:562) */
  const float* b_start; /* (n_ch) */
  const float* w_in[RADTTS_MAX_LAYERS]; /* (n_ch, n_ch, ksize), dilation 2^i */
  const float* b_in[RADTTS_MAX_LAYERS];
  const float* w_rs[RADTTS_MAX_LAYERS]; /* (n_ch, n_ch) */
  const float* b_rs[RADTTS_MAX_LAYERS];
  const float* w_end;   /* (2h, n_ch): rows [0,h) raw scale, [h,2h) translation */
  const float* b_end;   /* (2h) */
  const float* wg_start;                 /* (n_ch) weight_g of `start`, or NULL */
  const float* wg_in[RADTTS_MAX_LAYERS]; /* (n_ch) weight_g of in_layers[i], or NULL */
  const float* wg_rs[RADTTS_MAX_LAYERS]; /* (n_ch) weight_g of res_skip_layers[i], or NULL */
} radtts_flow_weights;

/* Packed activation buffers of one flow step; caller-allocated, `rows` = radtts_frameplan_rows().
 * "act" element type is float32 (PREC_FP32) or bfloat16 (PREC_BF16). */
typedef struct radtts_flow_buffers {
  const void* ctx; /* act [rows][ctx_ld], ctx_ld = round_up(n_ctx, 64): packed conditioning */
  const float* zin;/* [rows][z_ld] input of the flow */
  float* zmid;     /* [rows][z_ld] forward: after the 1x1 conv; inverse: after the inverse coupling */
  float* zout;     /* [rows][z_ld] output of the flow */
  void* z0;        /* act [rows][128] copy of the untransformed half (zero padded) */
  void* x;         /* act [n_layers + 1][rows][n_ch]: `start` output then every in_layer output */
  void* r;         /* act [rows][n_layers * n_ch]: softplus(res_skip) of every layer, K-concatenated */
  float* params;   /* [rows][z_ld] interleaved (raw scale, translation) pairs (forward; may be NULL) */
  float* log_s;    /* [rows][z_ld / 2] (forward only) */
} radtts_flow_buffers;

/* Re-layouts the weights for the GEMM kernels (tap-major K, interleaved `end`, bf16 cast, and -- when
 * want_backward -- the transposed copies the dgrad GEMMs read).  Run once per optimizer step (training) or
 * once per model (inference). */
RADTTS_API size_t radtts_flow_prepared_bytes(const radtts_flow_dims* dims, int precision, int want_backward);
RADTTS_API int radtts_flow_prepare(const radtts_flow_dims* dims, const radtts_flow_weights* w, int inverse,
                                   int precision, int want_backward, void* prepared, size_t prepared_bytes,
                                   void* stream);
RADTTS_API int radtts_flowstep_forward(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                       int Tmax, const radtts_flow_buffers* buf, int precision, void* stream);
RADTTS_API int radtts_flowstep_inverse(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                       int Tmax, const radtts_flow_buffers* buf, int precision, void* stream);

/* One WN in_layer (ConvNorm + PartialConv1d k5, dilation 2^layer, + softplus; reference common.py:573) in isolation:
 * x[layer] -> x[layer + 1] inside an act buffer [n_layers + 1][rows][n_ch].  This is the dominant kernel of the
 * step; the entry point exists so that it can be timed and profiled by itself. */
RADTTS_API int radtts_wn_layer_forward(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                       int Tmax, void* x, int layer, int precision, void* stream);

/* Backward of radtts_flowstep_forward (what PyTorch autograd derives for the reference).
 * Consumes the activation buffers the forward call filled (`fwd`) and a prepared blob built with
 * want_backward = 1.  Weight gradients are float32 in the reference's PyTorch layouts and are OVERWRITTEN. */
typedef struct radtts_flow_grad_buffers {
  const float* g_zout;  /* [rows][z_ld] dL/d zout */
  const float* g_log_s; /* [rows][z_ld/2] dL/d log_s (NULL = zeros) */
  float* g_zin;         /* [rows][z_ld] out */
  float* g_ctx;         /* [rows][ctx_ld] out, float32; += when accumulate_ctx */
  float* g_zmid;        /* scratch [rows][z_ld] */
  void* g_params;       /* scratch act [rows][round_up(z_ld, 64)] */
  void* g_u;            /* scratch act [n_layers][rows][n_ch] */
  void* g_v;            /* scratch act [n_layers][rows][n_ch] */
  void* g_x0;           /* scratch act [rows][n_ch] */
  float* g_w_inv_full;  /* out [z_ld][z_ld]: gradient of the identity-embedded 1x1 matrix (active block = dL/dW) */
  float* g_w_start;     /* out (n_ch, h + n_ctx) */
  float* g_b_start;
  float* g_w_in[RADTTS_MAX_LAYERS]; /* out TAP-MAJOR (ksize, n_ch, n_ch): [t][out][in]; permute(1,2,0) = reference layout */
  float* g_b_in[RADTTS_MAX_LAYERS];
  float* g_w_rs[RADTTS_MAX_LAYERS]; /* out (n_ch, n_ch) */
  float* g_b_rs[RADTTS_MAX_LAYERS];
  float* g_w_end;       /* out (2h, n_ch) */
  float* g_b_end;       /* out (2h) */
  float* scratch_f32;   /* scratch [z_ld][n_ch + 1] */
} radtts_flow_grad_buffers;
RADTTS_API int radtts_flowstep_backward(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                        int Tmax, const radtts_flow_buffers* fwd, const radtts_flow_grad_buffers* g,
                                        int accumulate_ctx, int precision, void* stream);

/* The same with an auxiliary stream (bf16 only; NULL = none): the memory-bound helpers that nothing on the NEXT flow's
 * dgrad chain depends on (fp32 1x1-conv weight gradient, bias column sums, `end` de-interleave) are enqueued on
 * aux_stream, ordered after the dgrad chain / the batched wgrad launch by events.  The caller must (i) not reuse the
 * scratch or gradient buffers of this call before aux_stream has finished with them and (ii) join aux_stream before
 * anything reads the gradients. */
RADTTS_API int radtts_flowstep_backward_ex(const radtts_flow_dims* dims, const void* prepared, const void* plan, int B,
                                           int Tmax, const radtts_flow_buffers* fwd, const radtts_flow_grad_buffers* g,
                                           int accumulate_ctx, int precision, void* stream, void* aux_stream);

/* Weight-norm backward for the convs of one flow (what autograd derives for torch._weight_norm, reference
 * common.py:540-556): from the effective-weight gradients radtts_flowstep_backward wrote (g_w_start, TAP-MAJOR
 * g_w_in, g_w_rs) to the gradients of weight_v (reference layout) and weight_g:
 *     grad_g[n] = <v_n, gw_n> / ||v_n||,   grad_v[n] = g[n]/||v_n|| * (gw_n - v_n <v_n, gw_n> / ||v_n||^2).
 * `w` holds weight_v / weight_g as for radtts_flow_prepare; a conv with wg_* = NULL just gets its gradient copied
 * (re-laid out) into gv_*.  accumulate != 0: += into gv_x / gg_x (gradient accumulation straight into .grad). */
typedef struct radtts_flow_wn_grads {
  float* gv_start; float* gg_start;                               /* (n_ch, h + n_ctx), (n_ch) */
  float* gv_in[RADTTS_MAX_LAYERS]; float* gg_in[RADTTS_MAX_LAYERS]; /* (n_ch, n_ch, ksize), (n_ch) */
  float* gv_rs[RADTTS_MAX_LAYERS]; float* gg_rs[RADTTS_MAX_LAYERS]; /* (n_ch, n_ch), (n_ch) */
} radtts_flow_wn_grads;
RADTTS_API int radtts_flow_weight_norm_backward(const radtts_flow_dims* dims, const radtts_flow_weights* w,
                                                const radtts_flow_grad_buffers* g, const radtts_flow_wn_grads* out,
                                                int accumulate, void* stream);

/* LU-parameterised 1x1-conv weights of a stack of n_flows flows (Invertible1x1ConvLUS, reference common.py:407-428),
 * batched over the flows:  W_k = P_k (tril(lower_k,-1) + diag(lower_diag_k)) (triu(upper_k,1) + diag(upper_diag_k)),
 * log_det_k = sum log|upper_diag_k|.  Arrays suffixed _host are HOST arrays of n_flows entries (device pointers / ints);
 * n_host[k] = channel count C_k <= 256; every matrix is dense row-major (C_k, C_k) float32; tmp_k is a (C_k, C_k) scratch.
 * backward (what autograd derives): g_w_k = dL/dW_k with row stride g_w_ld_host[k], g_log_det_k a
 * device scalar or NULL -> g_lower_k (strictly lower, zeros elsewhere), g_upper_k (strictly upper), g_upper_diag_k;
 * all OVERWRITTEN.  n_flows <= 32 (compose) / 16 (backward). */
RADTTS_API int radtts_lus_compose(int n_flows, const int* n_host, const float* const* lower_host,
                                  const float* const* upper_host, const float* const* upper_diag_host,
                                  const float* const* lower_diag_host, const float* const* p_host,
                                  float* const* tmp_host, float* const* w_host, float* const* log_det_host, void* stream);
RADTTS_API int radtts_lus_backward(int n_flows, const int* n_host, const float* const* lower_host,
                                   const float* const* upper_host, const float* const* upper_diag_host,
                                   const float* const* lower_diag_host, const float* const* p_host,
                                   const float* const* g_w_host, const int* g_w_ld_host,
                                   const float* const* g_log_det_host, float* const* tmp_host, float* const* g_lower_host,
                                   float* const* g_upper_host, float* const* g_upper_diag_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Kernel 3 -- ConvAttention core: pairwise L2 distance + log-softmax over text + log prior + masked softmax.
 * Replaces: ConvAttention.forward after the key/query projections (reference common.py:907-923) and its
 * autograd backward.  The (B, C, T1, T2) broadcast temporary of the reference is never materialised.
 *   q_enc (B, C, T1), k_enc (B, C, T2): projected queries / keys, float32, C <= 128 (80), T2 <= 576.
 *   prior (B, T1, T2) or NULL; key_lens (B) int64 or NULL (keys >= key_lens[b] are masked in the softmax,
 *   but -- like the reference -- still take part in the log-softmax).
 *   attn, attn_logprob (B,1,T1,T2) out; lse (B,T1) out: log-sum-exp of the distances, saved for backward.
 * backward: g_attn / g_logprob (either may be NULL) -> g_q (may be NULL), g_k; gd_ws is a (B,T1,T2) scratch.
 * ---------------------------------------------------------------------------------------------- */
RADTTS_API int radtts_convattn_forward(const float* q_enc, const float* k_enc, const float* prior,
                                       const int64_t* key_lens, int B, int C, int T1, int T2, float temp, float* attn,
                                       float* attn_logprob, float* lse, void* stream);
RADTTS_API int radtts_convattn_backward(const float* q_enc, const float* k_enc, const float* lse, const float* attn,
                                        const float* g_attn, const float* g_logprob, int has_prior, int B, int C, int T1,
                                        int T2, float temp, float* gd_ws, float* g_q, float* g_k, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Attribute flows (BGAP, config_ljs_bgap): SimpleConvNet layers, RQ-spline / affine coupling, tiny 1x1 conv.
 * ---------------------------------------------------------------------------------------------- */
/* One ConvNorm layer on packed rows (reference common.py:145-154 with partialconv1d.py:35-71 when partial = 1):
 * y = act(ratio * conv_k,dil(x) + bias) on valid rows, zeros on gap rows when mask_rows = 1.
 * act: 0 none, 1 softplus, 2 relu.  x: act [rows][ld_x] (first c_in_pad columns, c_in_pad % 64 == 0, zero padded);
 * y: act [rows][ld_y], columns [y_col_off, y_col_off + round_up(c_out, 16)) are written.
 * The weight (c_out, c_in, ksize) float32 is re-laid out once by radtts_conv_prepare. */
RADTTS_API size_t radtts_conv_prepared_bytes(int c_out, int c_in_pad, int ksize, int precision);
RADTTS_API int radtts_conv_prepare(const float* w, const float* bias, int c_out, int c_in, int c_in_pad, int ksize,
                                   int precision, void* prepared, size_t prepared_bytes, void* stream);
RADTTS_API int radtts_conv_rows(const void* prepared, int c_out, int c_in_pad, int ksize, int dilation, int act,
                                int partial, int mask_rows, const void* x, int ld_x, void* y, int ld_y, int y_col_off,
                                const void* plan, int B, int Tmax, int precision, void* stream);
/* Backward of radtts_conv_rows (what autograd derives for ConvNorm / PartialConv1d + activation, reference
 * common.py:145-154, partialconv1d.py:35-71): from g_y (act [rows][ld_gy], gradient w.r.t. the layer OUTPUT; zeros on gap
 * rows) and the saved forward input x / output y to
 *   g_x   act [rows][ld_gx] (first c_in_pad columns written; may be NULL for the first layer of a stack),
 *   g_w   float32 (c_out, c_in, ksize), the reference layout, OVERWRITTEN,
 *   g_b   float32, round_up(c_out, 16) floats, OVERWRITTEN (may be NULL).
 * prepared_t: the transposed weights (radtts_conv_prepare_backward).  g_pre: scratch act [rows][round_up(c_out, 64)].
 * act / partial / mask_rows / dilation as in the forward call.  The dgrad is a row GEMM with the taps reversed, the weight
 * gradient one tcgen05 problem per tap with K = packed rows, the bias gradient a streaming column sum. */
RADTTS_API size_t radtts_conv_backward_prepared_bytes(int c_out, int c_in_pad, int ksize, int precision);
RADTTS_API int radtts_conv_prepare_backward(const float* w, int c_out, int c_in, int c_in_pad, int ksize, int precision,
                                            void* prepared_t, size_t prepared_bytes, void* stream);
RADTTS_API int radtts_conv_rows_backward(const void* prepared_t, int c_out, int c_in, int c_in_pad, int ksize, int dilation,
                                         int act, int partial, int mask_rows, const void* x, int ld_x, const void* y,
                                         int ld_y, int y_col_off, const void* g_y, int ld_gy, void* g_pre, void* g_x,
                                         int ld_gx, float* g_w, float* g_b, const void* plan, int B, int Tmax,
                                         int precision, void* stream);

/* Rational-quadratic spline coupling transform (reference splines.py:221-319 through
 * SplineTransformationLayer.forward, common.py:699-743), (B, C, T) float32 tensors:
 * x: coupling input (first C/2 channels pass through, last C/2 are transformed); params (B, C/2 * (2 n_bins + 1), T)
 * = SimpleConvNet output; inverse = 1 for sampling.  y (B, C, T); log_s (B, T) (forward only, may be NULL). */
RADTTS_API int radtts_rqspline_apply(const float* x, const float* params, int B, int C, int T, int n_bins, int inverse,
                                     float left, float right, float bottom, float top, float* y, float* log_s,
                                     void* stream);
/* Backward of radtts_rqspline_apply in the forward direction (training the attribute flows): g_y (B, C, T) and g_log_s
 * (B, 1, T) may be NULL (zeros); g_x (B, C, T) = gradient w.r.t. the coupling input (pass-through half = g_y), g_params
 * (B, (C/2)(2 n_bins + 1), T) = gradient w.r.t. the raw spline parameters.  Closed form of what autograd derives for
 * splines.py:254-319 (blueprint: oracle/spline_grad.py). */
RADTTS_API int radtts_rqspline_backward(const float* x, const float* params, const float* g_y, const float* g_log_s, int B,
                                        int C, int T, int n_bins, float left, float right, float bottom, float top,
                                        float* g_x, float* g_params, void* stream);
/* Affine coupling apply (reference common.py:782-784,821-832): params (B, C, T) = [raw scale | translation];
 * scaling 0 tanh, 1 exp, 2 sigmoid, 3 translate.  log_s (B, C/2, T) forward only, may be NULL. */
RADTTS_API int radtts_affine_apply(const float* z, const float* params, int B, int C, int T, int scaling, int inverse,
                                   float* y, float* log_s, void* stream);
/* Backward of radtts_affine_apply in the forward direction (training the attribute flows): g_y (B, C, T) / g_log_s
 * (B, C/2, T) may be NULL (zeros) -> g_z (B, C, T), g_params (B, C, T). */
RADTTS_API int radtts_affine_backward(const float* z, const float* params, const float* g_y, const float* g_log_s, int B,
                                      int C, int T, int scaling, float* g_z, float* g_params, void* stream);
/* y[b,:,t] = W x[b,:,t] for a small dense W (C <= 16): the plain Invertible1x1Conv of BGAP (common.py:431-472), and its
 * backward: g_x = W^T g_y, g_w (C, C) = sum over (b, t) of g_y x^T (OVERWRITTEN). */
RADTTS_API int radtts_pointwise_conv_small(const float* x, const float* w, int B, int C, int T, float* y, void* stream);
RADTTS_API int radtts_pointwise_conv_small_backward(const float* x, const float* w, const float* g_y, int B, int C, int T,
                                                    float* g_x, float* g_w, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Hard-attention context: context = bmm(text_enc, attn_hard^T) of reference radtts.py:399 as a gather by
 * the frame -> token index radtts_mas_forward produces, and its backward (segment sum per token).
 *   text_enc (B, C, T2), frame_to_token (B, T1) int32 (-1 beyond out_len), context (B, C, T1); float32.
 *   Frame 0 also receives token 0 when the path does not start there (the reference's opt[0,0] = 1, alignment.py:59).
 * ---------------------------------------------------------------------------------------------- */
RADTTS_API int radtts_context_gather(const float* text_enc, const int32_t* frame_to_token, int B, int C, int T1, int T2,
                                     float* context, void* stream);
RADTTS_API int radtts_context_scatter(const float* grad_context, const int32_t* frame_to_token, int B, int C, int T1,
                                      int T2, float* grad_text_enc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Persistent bidirectional LSTM recurrence (context BiLSTM, reference radtts.py:284-293, and the text
 * encoder BiLSTM, common.py:359-371 -- a packed-sequence cuDNN call in the reference).  One cooperative kernel per
 * pass runs the whole time loop; the x-projection GEMM and the weight-gradient GEMMs stay outside (plain GEMMs).
 *   gx (2, T, B, 4H) float32: W_ih x_t + b_ih + b_hh per direction (gate order i, f, g, o); whh (2, 4H, H);
 *   lens (B) int32; h_all (T, B, 2H) out, zeros for t >= lens[b] (packed-sequence semantics);
 *   gates_save (2, T, B, 4H), c_save (2, T, B, H): activations saved for backward (NULL for inference);
 *   backward: dh_all (T, B, 2H) -> dgates_all (2, T, B, 4H) (pre-activation gate gradients).
 *   B <= 32, H <= 592.  ws: radtts_lstm_workspace_bytes(B, H).
 *   precision: RADTTS_PREC_FP32 -- fp32 FMA recurrence; RADTTS_PREC_BF16 (H % 8 == 0, else fp32) -- the recurrent product runs
 *   on tensor cores with bf16 W_hh and bf16 exchanged h / dgates, fp32 accumulate; state, gates and every output fp32;
 *   RADTTS_PREC_BF16X2 -- the same with every operand split into hi + lo bf16 parts (3 MMAs, 16 mantissa bits): meant
 *   for fp32 LSTMs wherever cuDNN itself would be allowed TF32 (10 mantissa bits).
 * ---------------------------------------------------------------------------------------------- */
RADTTS_API size_t radtts_lstm_workspace_bytes(int B, int H);
RADTTS_API int radtts_lstm_forward(const float* gx, const float* whh, const int* lens, int T, int B, int H,
                                   float* h_all, float* gates_save, float* c_save, void* ws, size_t ws_bytes,
                                   int precision, void* stream);
RADTTS_API int radtts_lstm_backward(const float* dh_all, const float* whh, const int* lens, const float* gates_save,
                                    const float* c_save, int T, int B, int H, float* dgates_all, void* ws,
                                    size_t ws_bytes, int precision, void* stream);

/* Text-encoder conv block epilogue (reference common.py:348-356 with Encoder.convolutions = partial-padding ConvNorm +
 * InstanceNorm1d(affine) + ReLU + dropout, run per utterance in the reference), batched and fused: from the raw k-tap conv
 * output `raw` (B, C, T) = conv1d(x * mask) + conv_bias to
 *     out = dropout(relu(instance_norm(partial_conv_renormalise(raw)) * gamma + beta)) * [t < len]
 * with per-(utterance, channel) statistics over the valid frames only.  drop: keep mask (0/1 floats, same shape) or NULL,
 * multiplied by drop_scale = 1 / (1 - p).  mean / rstd (B, C) are saved for the backward call, which returns g_raw and
 * the gradients of gamma, beta and the DIRECT conv-bias term (the bias also sits inside raw; that part comes from the conv's
 * own backward).  float32; lens (B) int64 device. */
RADTTS_API int radtts_encnorm_forward(const float* raw, const float* conv_bias, const int64_t* lens, const float* gamma,
                                      const float* beta, const float* drop, float drop_scale, int B, int C, int T,
                                      int ksize, float eps, float* out, float* mean, float* rstd, void* stream);
RADTTS_API int radtts_encnorm_backward(const float* raw, const float* conv_bias, const int64_t* lens, const float* gamma,
                                       const float* beta, const float* drop, float drop_scale, const float* mean,
                                       const float* rstd, const float* g_out, int B, int C, int T, int ksize, float* g_raw,
                                       float* g_gamma, float* g_beta, float* g_conv_bias, void* stream);

/* Diagnostic: phase cycle sums (clock64) of cluster rank 0 of the last cluster-LSTM launch, SIXTEEN host words: forward
 * [0] waiting for h, [1] MMAs, [2] gates + staging, [4] send + output stores, [5] steps.  Synchronises the device. */
RADTTS_API int radtts_lstm_debug_timeline(unsigned long long* out16_host);

/* ------------------------------------------------------------------------------------------------
 * Attention CTC loss, fused.  Replaces AttentionCTCLoss.forward (reference loss.py:118-135): blank
 * logit + per-utterance log-softmax + nn.CTCLoss(zero_infinity=True) against targets 1..K_b, and its backward.
 *   attn_logprob (B,1,T1,T2) float32 logits; in_lens / out_lens (B) int64 device; T2 <= 511.
 *   losses (B) out: -log p(target) per utterance (0 when infinite); the reference's scalar is mean_b(losses[b]/K_b).
 *   grad (B,1,T1,T2) out: d(mean_b(losses[b]/K_b)) / d attn_logprob, fully written.
 *   ws: radtts_attn_ctc_workspace_bytes (alpha lattice).
 * ---------------------------------------------------------------------------------------------- */
RADTTS_API size_t radtts_attn_ctc_workspace_bytes(int B, int T1, int T2);
RADTTS_API int radtts_attn_ctc(const float* attn_logprob, const int64_t* in_lens, const int64_t* out_lens, int B, int T1,
                               int T2, float blank_logprob, float* losses, float* grad, void* ws, size_t ws_bytes,
                               void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused RAdam over a flat float32 parameter buffer (update rule of reference radam.py:76-116).
 * p, g, m, v: device pointers to n floats (16-byte aligned); step_dev: device int64 counter (number of steps
 * taken so far; incremented by the call); grad_scale: device scalar multiplied into g (clipping), or NULL.
 * ---------------------------------------------------------------------------------------------- */
RADTTS_API int radtts_radam_step(float* p, const float* g, float* m, float* v, size_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, long long* step_dev,
                                 const float* grad_scale, void* stream);
/* The same update on a sub-range of the buffer, for a trainer that pipelines it: `enable` (device int, or NULL) turns the
 * whole call into a no-op when it holds 0 -- a captured step can carry "apply the pending update of the previous step"
 * as a fixed node; zero_grad != 0 writes zeros back over g in the same pass (saves the separate memset of the gradient
 * buffer); bump != 0 increments step_dev after the update (under the same `enable`).  All ranges updated with the same
 * step_dev value belong to one optimizer step: bump once, with the last of them. */
RADTTS_API int radtts_radam_step_ex(float* p, float* g, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                                    float eps, float weight_decay, long long* step_dev, const float* grad_scale,
                                    const int* enable, int zero_grad, int bump, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RADTTS_B200_H_ */
