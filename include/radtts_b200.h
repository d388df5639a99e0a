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
 *   - re-entrant; the only global state is cached cudaFuncSetAttribute / driver entry points.
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

#ifdef __cplusplus
}
#endif
#endif /* RADTTS_B200_H_ */
