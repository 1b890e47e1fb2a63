/*
 * b200nlp.h — C-ABI of libb200nlp.so: hand-written sm_100a kernels for the PaddleNLP LLM decoder hot path.
 *
 * This is the drop-in boundary.  Each entry point replaces one native op the reference reaches
 * through Paddle's custom-op C++ API (PD_BUILD_OP) or through a Paddle-core kernel; the reference call site is
 * cited beside each declaration (paths relative to the PaddleNLP tree).
 *
 * Conventions
 *   - plain C, no framework types: device pointers + int64 sizes + scalar attributes + cudaStream_t.
 *   - the CALLER owns every buffer (inputs, outputs, workspaces); kernels never allocate or free device memory.
 *   - all work is enqueued on `stream`; no host synchronisation inside any entry point.
 *   - return value: 0 = ok, <0 = argument error, >0 = cudaError_t of the failed launch;
 *     b200_last_error() returns a thread-local message for the last non-zero return.
 *   - bf16 tensors are `__nv_bfloat16` bit patterns (uint16), row-major, contiguous unless a leading dimension
 *     is passed.  "T" below is the token count (batch * seq).
 */
#ifndef B200NLP_H_
#define B200NLP_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200NLP_ABI_VERSION 1

#ifndef __CUDA_RUNTIME_API_H__
typedef struct CUstream_st* cudaStream_t;
#endif

/* ---- plumbing ------------------------------------------------------------------------------------------ */
const char* b200_last_error(void);
int b200_abi_version(void);
/* 0 if the current device is compute capability 10.x (B200); error otherwise. */
int b200_device_check(void);

/* ---- GEMM: replaces paddle.matmul / nn.Linear (cuBLASLt) --------------------------------------------------
 * C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N]);  bf16 operands, fp32 accumulation in TMEM, ONE rounding to bf16.
 *   a_mn_major = 0 : A is stored [M,K] row-major (lda = row stride in elements)          — activations
 *   a_mn_major = 1 : A is stored [K,M] row-major (i.e. the caller passes A^T)            — dW = X^T * dY
 *   b_mn_major = 1 : B is stored [K,N] row-major — Paddle's nn.Linear weight layout [in,out]
 *   b_mn_major = 0 : B is stored [N,K] row-major (i.e. B^T)                               — dX = dY * W^T
 *   accumulate != 0: C_new = bf16(fp32(C_old) + acc)   (gradient accumulation, cf. llm/utils/fused_layers.py:36-74)
 *   bias            : optional fp32 [N], added in fp32 before the rounding (Qwen2 q/k/v bias, qwen2/modeling.py:478-480)
 * Reference call sites: llama/modeling.py:933-935,1103 (q/k/v/o), :632-652 (gate/up/down), :1894-1921 (lm_head).
 * Requires M, N, K, lda, ldb, ldc multiples of 8 and 16-byte aligned base pointers.
 */
int b200_gemm_bf16(const void* A, const void* B, void* C, const float* bias, int64_t M, int64_t N, int64_t K,
                   int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int accumulate,
                   cudaStream_t stream);
/* Same with tuning knobs: cta_group 1 (one CTA per 128x256 tile) or 2 (CTA pair per 256x256 tile);
 * max_ctas > 0 limits the persistent grid (used to leave SMs to a concurrent kernel). */
int b200_gemm_bf16_ex(const void* A, const void* B, void* C, const float* bias, int64_t M, int64_t N, int64_t K,
                      int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int accumulate,
                      int cta_group, int max_ctas, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200NLP_H_ */
