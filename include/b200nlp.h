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
/* Same with a fused residual epilogue and tuning knobs.
 *   residual (bf16 [M,N], leading dimension ldr; exclusive with accumulate):
 *       C = bf16( bf16(acc + bias) + residual )  — the Linear-output rounding followed by the decoder layer's residual
 *       add (llama/modeling.py:1212, 1218), i.e. the reference's two rounding points in one kernel.
 *   cta_group 1 (one CTA per 128x256 tile) or 2 (CTA pair per 256x256 tile);
 *   max_ctas > 0 limits the persistent grid (used to leave SMs to a concurrent kernel). */
int b200_gemm_bf16_ex(const void* A, const void* B, void* C, const float* bias, const void* residual, int64_t M,
                      int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr, int a_mn_major,
                      int b_mn_major, int accumulate, int cta_group, int max_ctas, cudaStream_t stream);

/* ---- RMSNorm: replaces fused_ln.fused_rms_norm / fast_ln (apex-derived custom ops) --------------------------
 * fwd : y = bf16( bf16(x * rstd) * w ), rstd[row] = rsqrt(mean(x^2) + eps) in fp32 (saved for the backward).
 * bwd : dx = rstd * (dy*w - xhat * mean(dy*w*xhat)) (+ dres, the gradient arriving through the residual branch);
 *       dw (+)= sum_rows dy * bf16(xhat).   workspace: b200_rmsnorm_bwd_workspace_bytes(rows, h) bytes.
 * Reference: llama/modeling.py:352-386; fusion_ops.py:119-144; legacy/model_zoo/gpt-3/external_ops/fused_ln/
 * layer_norm_cuda.cu:47-66 (fwd), :164-183 (bwd); layer_norm_cuda.h:447-531, 1190-1260.
 */
int b200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t rows, int64_t h, float eps,
                     cudaStream_t stream);
int64_t b200_rmsnorm_bwd_workspace_bytes(int64_t rows, int64_t h);
int b200_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx,
                     void* dw, int accumulate_dw, void* workspace, int64_t rows, int64_t h, cudaStream_t stream);

/* Column sums of a bf16 [rows, n] matrix (leading dimension ld) into a bf16 vector: bias gradients of Qwen2 q/k/v
 * (qwen2/modeling.py:478-480).  workspace: b200_colsum_workspace_bytes(rows, n). */
int64_t b200_colsum_workspace_bytes(int64_t rows, int64_t n);
int b200_colsum_bf16(const void* a, void* out, int accumulate, void* workspace, int64_t rows, int64_t n, int64_t ld,
                     cudaStream_t stream);

/* ---- RoPE (rotate-half), in place on `num_heads` consecutive heads starting at x: replaces Paddle-core
 * fused_rotary_position_embedding(use_neox_rotary_style=False) (fusion_ops.py:57-116; llama/modeling.py:557-577).
 * cos/sin tables: fp32 [max_pos, head_dim/2]; position of token t = position_ids[t] or t %% seq_len.
 * backward != 0 applies the transposed rotation (sin -> -sin). */
int b200_rope_inplace(void* x, const float* cos_table, const float* sin_table, const int32_t* position_ids,
                      int64_t tokens, int64_t seq_len, int64_t ld, int64_t num_heads, int64_t head_dim, int backward,
                      cudaStream_t stream);

/* ---- SwiGLU on a packed [rows, 2*inter] = [gate | up] buffer: replaces Paddle-core swiglu
 * (llama/modeling.py:38-45, 648-650).  bwd writes [dgate | dup] packed the same way. */
int b200_swiglu_fwd(const void* gate_up, void* out, int64_t rows, int64_t inter, cudaStream_t stream);
int b200_swiglu_bwd(const void* gate_up, const void* dout, void* dgate_up, int64_t rows, int64_t inter,
                    cudaStream_t stream);

/* ---- Embedding gather / scatter-add (nn.Embedding, llama/modeling.py:1465-1468, 1634). ids are int64. */
int b200_embedding_fwd(const int64_t* ids, const void* table, void* out, int64_t tokens, int64_t h, int64_t vocab,
                       cudaStream_t stream);
int b200_embedding_bwd(const int64_t* ids, const void* dout, void* dtable, int64_t tokens, int64_t h, int64_t vocab,
                       cudaStream_t stream);

/* ---- Flash attention, causal, GQA, head_dim 128: replaces F.scaled_dot_product_attention(is_causal=True)
 * (fusion_ops.py:147-267; Paddle-vendored FlashAttention-2) and its gradient (csrc/gpu/flash_attn_bwd.cc:22-92).
 * q [B,S,nh,128], k/v [B,S,kvh,128], o [B,S,nh,128]; ld* = token stride in elements (the tensors may be views into a
 * packed QKV projection).  lse [B,nh,S] fp32 (natural log).  Backward workspace: b200_fa_bwd_workspace_bytes(). */
int b200_fa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int64_t B, int64_t S,
                int64_t num_heads, int64_t num_kv_heads, int64_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv,
                int64_t ldo, float softmax_scale, cudaStream_t stream);
int64_t b200_fa_bwd_workspace_bytes(int64_t B, int64_t S, int64_t num_heads, int64_t head_dim);
int b200_fa_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t S, int64_t num_heads,
                int64_t num_kv_heads, int64_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv, float softmax_scale, cudaStream_t stream);

/* ---- Criterion: LlamaPretrainingCriterion (llama/modeling.py:1799-1825) on bf16 logits [tokens, vocab] (ld).
 * fwd : loss_tok[i] = fp32 CE (0 for ignore_index), lse[i]; loss_out[0] = sum(l_i [l_i>0]) / count, loss_out[1] = count.
 * bwd : logits are overwritten by dlogits = (softmax - onehot) * [l_i>0] * grad_scale / count (bf16);
 *       grad_scale_dev (optional device scalar) multiplies grad_scale, so an upstream gradient that lives on the
 *       device (loss / gradient_accumulation_steps, trainer.py:2237-2238) needs no host synchronisation. */
int b200_ce_fwd(const void* logits, const int64_t* labels, float* loss_tok, float* lse, float* loss_out, int64_t tokens,
                int64_t vocab, int64_t ld, int64_t ignore_index, cudaStream_t stream);
int b200_ce_bwd(void* logits_inout, const int64_t* labels, const float* loss_tok, const float* lse,
                const float* loss_out, float grad_scale, const float* grad_scale_dev, int64_t tokens, int64_t vocab,
                int64_t ld, cudaStream_t stream);
/* Greedy token choice: first maximal index of each bf16 row (generation_utils.py:291-363 with top_p = 0). */
int b200_argmax_bf16(const void* logits, int64_t* out, int64_t rows, int64_t vocab, int64_t ld, cudaStream_t stream);

/* ---- Optimizer on the flat parameter buffer: ClipGradByGlobalNorm + AdamW(multi_precision)
 * (trainer.py:1717-1750).  Elements [0, decay_end) receive weight decay.
 * grad_sqnorm: out[0] = || scale * g ||^2 ; adamw: g_eff = g * grad_scale * max_norm / max(||.||, max_norm). */
int64_t b200_grad_sqnorm_workspace_bytes(void);
int b200_grad_sqnorm(const void* grads, float* out, void* workspace, int64_t n, float scale, cudaStream_t stream);
int b200_adamw_step(void* params_bf16, const void* grads_bf16, float* master, float* exp_avg, float* exp_avg_sq,
                    const float* grad_sqnorm, int64_t n, int64_t decay_end, float lr, float beta1, float beta2, float eps,
                    float weight_decay, int64_t step, float grad_scale, float max_grad_norm, cudaStream_t stream);
int b200_bf16_to_f32(const void* src, float* dst, int64_t n, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200NLP_H_ */
