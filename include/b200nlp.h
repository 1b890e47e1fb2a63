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

#include <stdbool.h>
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
/* Programmatic dependent launch for the GEMM kernels (returns the previous setting; NOT an error code): when enabled, a GEMM
 * may start while the previous kernel of the stream is still draining — it prefetches its weight tiles and sets up TMEM /
 * barriers, and only waits (griddepcontrol.wait) before touching activations or outputs.  Used for the decode-step chain. */
int b200_set_pdl(int enable);
/* Which kernel serves the plain-causal b200_fa_fwd (returns the previous setting; NOT an error code): 2 (default) = two
 * 128-row q tiles per CTA with P kept in tensor memory (csrc/fa_fwd2.cu), 1 = one q tile per CTA (csrc/fa_fwd.cu, which
 * also serves every FlashMask call).  Same rounding points; kept switchable for A/B measurements.  The initial value can
 * be set with the environment variable B200_FA_FWD_IMPL. */
int b200_set_fa_fwd_impl(int impl);
/* Same for the plain-causal b200_fa_bwd: 2 (default) = transposed score tiles (kv on the UMMA M dimension), 64-row q steps,
 * S^T double-buffered, P^T kept in tensor memory (csrc/fa_bwd2.cu); 1 = csrc/fa_bwd.cu (also every FlashMask call).
 * Environment override: B200_FA_BWD_IMPL. */
int b200_set_fa_bwd_impl(int impl);
/* Share of the forward softmax exponentials (generation-2 kernel) evaluated by a degree-3 polynomial on the FMA pipe instead of
 * MUFU.EX2 (the MUFU's 16 ex2/clk/SM equals the tensor time of a kv step): 0 = none, 1 (default) = a quarter, 2 = half.
 * Relative error of the polynomial 7.5e-5, below the bf16 rounding P receives.  Environment override: B200_FA_EXP_POLY.
 * Returns the previous setting. */
int b200_set_fa_exp_poly(int mode);
/* Which kernel serves b200_gemm_bf16_splitk for M <= 128 with a row-major A (returns the previous setting; NOT an error
 * code): 1 (default) = the swapped-operand, two-CTA-per-SM weight-streaming kernel (csrc/gemm_skinny.cu), 2 = its stream-K
 * variant (M <= 64), 0 = the persistent 128x256 kernel in split-K mode.  Same results up to fp32 summation order; kept switchable
 * for A/B measurements. */
int b200_set_skinny_gemm(int impl);

/* ---- GEMM: replaces paddle.matmul / nn.Linear (cuBLASLt) --------------------------------------------------
 * C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N]);  bf16 operands, fp32 accumulation in TMEM, ONE rounding to bf16.
 *   a_mn_major = 0 : A is stored [M,K] row-major (lda = row stride in elements)          — activations
 *   a_mn_major = 1 : A is stored [K,M] row-major (i.e. the caller passes A^T)            — dW = X^T * dY
 *   b_mn_major = 1 : B is stored [K,N] row-major — Paddle's nn.Linear weight layout [in,out]
 *   b_mn_major = 0 : B is stored [N,K] row-major (i.e. B^T)                               — dX = dY * W^T
 *   accumulate != 0: C_new = bf16(fp32(C_old) + acc)   (gradient accumulation, cf. llm/utils/fused_layers.py:36-74)
 *   bias            : optional fp32 [N], added in fp32 before the rounding (Qwen2 q/k/v bias, qwen2/modeling.py:478-480)
 * Reference call sites: llama/modeling.py:933-935,1103 (q/k/v/o), :632-652 (gate/up/down), :1894-1921 (lm_head).
 * Requires lda, ldb, ldc multiples of 8 elements and 16-byte aligned base pointers (M, N, K themselves are free: TMA
 * zero-fills / clips partial tiles).
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

/* Weight-streaming GEMM for the decode step (M <= ~128 tokens): same math as b200_gemm_bf16 (C = op(A) op(B) + bias, one
 * rounding), but K is split over CTAs (split_k, 0 = auto) so that every SM streams part of the weight matrix; fp32 partial
 * tiles are summed in L2 by TMA reduce-add into `workspace` (b200_gemm_splitk_workspace_bytes; must be ZERO on entry,
 * is returned zeroed) and rounded once.  C == NULL skips the rounding pass: the consumer (b200_add_rmsnorm_f32,
 * b200_decode_rope_append_f32) reads the fp32 sums, rounds them once to bf16 and re-zeroes the workspace.
 * Replaces the cuBLASLt calls of FusedMultiTransformer's decode step (fused_transformer_layers.py:817-820, 895-896, 967-974). */
int64_t b200_gemm_splitk_workspace_bytes(int64_t M, int64_t N);
int b200_gemm_bf16_splitk(const void* A, const void* B, void* C, const float* bias, void* workspace, int64_t M, int64_t N,
                          int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int split_k,
                          cudaStream_t stream);

/* gate|up projection + SwiGLU in ONE kernel — LlamaMLP.forward with fuse_attention_ffn (llama/modeling.py:632-652, swiglu :38-45):
 *   GU[M, 2I] = bf16(X[M,K] * W[K,2I])  (gate columns [0,I), up columns [I,2I); kept for the backward),
 *   Mout[M, I] = bf16(silu(gate) * up)  with gate/up rounded to bf16 first (the unfused rounding points).
 * A 256-column tcgen05 tile is formed from 128 gate columns and the 128 up columns of the same channels, so no interleaved
 * weight layout is needed; requires I % 128 == 0.  Bit-identical to b200_gemm_bf16 followed by b200_swiglu_fwd. */
int b200_gemm_swiglu_bf16(const void* X, const void* W, void* GU, void* Mout, int64_t M, int64_t inter, int64_t K, int64_t ldx,
                          int64_t ldw, int64_t ldgu, int64_t ldm, int cta_group, cudaStream_t stream);
/* Backward twin: the down-projection dX GEMM with the SwiGLU backward in its epilogue.
 *   d(m) = dY[M,K] * Wdown[I,K]^T (never written);  DGU[M, 2I] = [ d(m) * up * silu'(gate) | d(m) * silu(gate) ],
 * GU = the saved gate|up projection [M, 2I].  Bit-identical to b200_gemm_bf16 (b_mn_major = 0) + b200_swiglu_bwd.  I % 64 == 0. */
int b200_gemm_swiglu_bwd_bf16(const void* dY, const void* Wdown, const void* GU, void* DGU, int64_t M, int64_t inter, int64_t K,
                              int64_t lddy, int64_t ldw, int64_t ldgu, int64_t lddgu, int cta_group, cudaStream_t stream);

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
/* Same, gate|up given as the fp32 split-K workspace [rows, 2*inter] of the producing GEMM (rounded to bf16 here, workspace
 * re-zeroed): the decode step's ffn1 -> fused_bias_act("swiglu") pair (fused_transformer_layers.py:100-168). */
int b200_swiglu_fwd_f32(float* gate_up_f32_ws, void* out, int64_t rows, int64_t inter, cudaStream_t stream);
/* Decode-step ffn1 + SwiGLU in one kernel (M <= 64 token rows): act[M, inter] = bf16(silu(g) * u) with g|u = bf16(X W).
 * W [K, 2*inter] is the reference-layout fused ffn1 weight (gate | up): one 128-feature tile of the swapped-operand
 * weight-streaming kernel is fed by two 64-column TMA boxes, gate channels [64j, 64j+64) and the up columns of the same channels
 * (fused_transformer_layers.py:100-168 fused_bias_act("swiglu") after ffn1).  inter % 64 == 0; ldact = row stride of act. */
int b200_gemm_swiglu_skinny(const void* X, const void* W_gate_up, void* act, int64_t M, int64_t inter, int64_t K,
                            int64_t ldx, int64_t ldw, int64_t ldact, cudaStream_t stream);
int b200_swiglu_bwd(const void* gate_up, const void* dout, void* dgate_up, int64_t rows, int64_t inter,
                    cudaStream_t stream);
/* The GEMM chain of one decode-step layer (M <= 64 token rows) as ONE persistent kernel of two CTAs per SM whose six steps are
 * separated by grid-wide barriers instead of kernel boundaries, the weight stream running ahead across them
 * (fused_transformer_layers.py:895-896 out-linear, :937-949 ffn layernorm, :100-168 ffn1 + swiglu, ffn2, :976-999 residual +
 * next layernorm, :843-856 the NEXT layer's qkv projection):
 *   acc_h += attn @ W_o ; residual += bf16(acc_h), ln = rmsnorm(residual) * w_ffn_ln ; act = swiglu(bf16(ln @ W_ffn1)) ;
 *   acc_h += act @ W_ffn2 ; residual += bf16(acc_h), ln = rmsnorm(residual) * w_next_ln ; acc_qkv += ln @ W_next_qkv^T
 * attn bf16 [M, attn_width]; W_o [attn_width, h], W_ffn1 [h, 2*inter] (gate | up), W_ffn2 [inter, h], W_next_qkv [qkv_n, h] (the
 * reference layouts); residual bf16 [M, h] in/out; ln_buf [M, h] and act_buf [M, inter] bf16 scratch; acc_h fp32 [M, h] and
 * acc_qkv fp32 [M, qkv_n]: zero on entry, acc_h zero again on exit, acc_qkv holds the projection sums (same contract as
 * b200_gemm_bf16_splitk's workspace: the RoPE-append consumer rounds and re-zeroes).  w_next_ln / w_next_qkv NULL (last layer):
 * the chain ends with the residual update.  sync_ws: b200_decode_layer_chain_workspace_bytes() bytes, zero before the first call
 * (handed back zeroed).  Same rounding points and summation structure as the unfused kernels. */
int64_t b200_decode_layer_chain_workspace_bytes(void);
/* Debugging aid (tools/decode_probe.py chain): stamps = device buffer of 2 * 6 * 8 bytes per CTA (2 x SM count CTAs) that the next
 * launches fill with %globaltimer values — [cta][phase][0] = the producer saw the phase's inputs, [1] = the CTA published the
 * phase; NULL switches it off. */
int b200_decode_layer_chain_debug(void* stamps);
int b200_decode_layer_chain(const void* attn, const void* w_o, const void* w_ffn_ln, const void* w_ffn1, const void* w_ffn2,
                            const void* w_next_ln, const void* w_next_qkv, void* residual, void* ln_buf, void* act_buf, float* acc_h,
                            float* acc_qkv, void* sync_ws, int64_t M, int64_t h, int64_t attn_width, int64_t inter, int64_t qkv_n,
                            float eps, cudaStream_t stream);

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
/* FlashMask, causal lower-triangular form: the same two ops with a per-key-column start row
 * (fusion_ops.py:218-231 -> F.flashmask_attention(q, k, v, startend_row_indices=..., causal=True)):
 * mask_start_rows [B, S] int32, query row i sees key column c iff c <= i < mask_start_rows[b, c].  For packed SFT samples
 * ("zero padding", llm/utils/data.py:200-204, paddlenlp/datasets/zero_padding_dataset.py:84-86) it is the end of c's document
 * and must be non-decreasing in c; kv tiles that a q tile cannot see are skipped, not just masked.  NULL = plain causal. */
int b200_fa_fwd_flashmask(const void* q, const void* k, const void* v, void* o, float* lse, const int32_t* mask_start_rows,
                          int64_t B, int64_t S, int64_t num_heads, int64_t num_kv_heads, int64_t head_dim, int64_t ldq,
                          int64_t ldk, int64_t ldv, int64_t ldo, float softmax_scale, cudaStream_t stream);
int b200_fa_bwd_flashmask(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                          const int32_t* mask_start_rows, void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t S,
                          int64_t num_heads, int64_t num_kv_heads, int64_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv,
                          int64_t ldo, int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv, float softmax_scale,
                          cudaStream_t stream);

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

/* ======================================================================================================================
 * Generation path (FusedMultiTransformer / paddlenlp_ops, bf16 non-quantised subset)
 * ====================================================================================================================== */

/* Fused residual add + RMSNorm: r = bf16(x + residual) (residual may be NULL), normed = RMSNorm(r) * w.
 * normed or residual_out may be NULL (last layer: residual add only).  Replaces Paddle-core
 * fused_rms_norm(x, w, ..., residual=) -> (out, residual_out) as called by
 * experimental/transformers/fused_transformer_layers.py:799-805, 937-949, 976-999. */
int b200_add_rmsnorm(const void* x, const void* residual, const void* w, void* normed, void* residual_out, int64_t rows,
                     int64_t h, float eps, cudaStream_t stream);
/* Same, x given as the fp32 split-K workspace [rows, h] of the producing GEMM (rounded to bf16 here, workspace re-zeroed). */
int b200_add_rmsnorm_f32(float* x_f32_ws, const void* residual, const void* w, void* normed, void* residual_out,
                         int64_t rows, int64_t h, float eps, cudaStream_t stream);

/* KV cache tensor: bf16 [2, B, kvh, max_len, d] (K then V), the shape the reference predictor allocates
 * (llm/predict/predictor.py:697-706; experimental/transformers/llama/modeling.py:1768-1794).
 * Prefill: copy rotated K and V rows of the packed [B*S, ld] QKV projection for positions s < seq_lens[b]
 * (seq_lens may be NULL = all S).  Replaces write_cache_kv (csrc/gpu/write_cache_kv.cu:23-99). */
int b200_write_cache_kv(const void* qkv, void* cache, const int32_t* seq_lens, int64_t B, int64_t S, int64_t num_heads,
                        int64_t num_kv_heads, int64_t head_dim, int64_t max_len, int64_t ld, cudaStream_t stream);
/* Decode: rotate-half RoPE of the new token's q,k at position seq_lens[b] (in place in qkv [B, ld]) and append k,v to
 * the cache.  Replaces the RoPE + cache-write half of masked_multihead_attention
 * (fused_transformer_layers.py:884-893) / append_attn/decoder_write_cache_with_rope_kernel.cu:47-390. */
int b200_decode_rope_append(void* qkv, void* cache, const float* cos_table, const float* sin_table, const int32_t* seq_lens,
                            int64_t B, int64_t num_heads, int64_t num_kv_heads, int64_t head_dim, int64_t max_len,
                            int64_t ld, cudaStream_t stream);
/* Same, the QKV projection given as the fp32 split-K workspace [B, (nh+2kvh)*d] (+ optional fp32 bias): rounded to bf16
 * into qkv first, workspace re-zeroed. */
int b200_decode_rope_append_f32(void* qkv, float* acc_f32_ws, const float* bias, void* cache, const float* cos_table,
                                const float* sin_table, const int32_t* seq_lens, int64_t B, int64_t num_heads,
                                int64_t num_kv_heads, int64_t head_dim, int64_t max_len, int64_t ld, cudaStream_t stream);
/* Decode attention of one query token per sequence over cache positions [0, seq_lens[b]] (GQA, head_dim 128);
 * out [B, nh*d].  num_splits > 1 splits each sequence's cache range over that many CTAs (split-KV, merged by a second
 * kernel through `workspace`), as the reference's append_attention does (append_attention_c16_impl.cuh:826-1000).
 * Replaces the attention half of masked_multihead_attention / append_attention decode
 * (csrc/gpu/append_attn/append_attention_c16_impl.cuh:377-744). */
int64_t b200_decode_attention_workspace_bytes(int64_t B, int64_t num_heads, int64_t num_splits);
int b200_decode_attention(const void* qkv, const void* cache, const int32_t* seq_lens, void* out, void* workspace, int64_t B,
                          int64_t num_heads, int64_t num_kv_heads, int64_t head_dim, int64_t max_len, int64_t ld,
                          float softmax_scale, int64_t num_splits, cudaStream_t stream);
/* Same contract on the tensor cores: a persistent tcgen05 kernel streams the cache with TMA through a 192 KB ring per SM
 * (S^T = K_tile Q^T and O^T += V_tile^T P^T, the G query heads of a group padded to N=16); GQA group size 1, 2, 4, 7 or 8.
 * Cache rows past the sequence length must hold finite values (zero-filled allocation, as the reference's paddle.zeros). */
int b200_decode_attention_tc(const void* qkv, const void* cache, const int32_t* seq_lens, void* out, void* workspace, int64_t B,
                             int64_t num_heads, int64_t num_kv_heads, int64_t head_dim, int64_t max_len, int64_t ld,
                             float softmax_scale, int64_t num_splits, cudaStream_t stream);

/* Paged ("block") KV cache of FusedBlockMultiTransformer / append_attention (fused_transformer_layers.py:2192-2354,
 * csrc/gpu/append_attention.cu:428-851): key_cache / value_cache [num_blocks, kvh, block_size, head_dim] bf16,
 * block_tables [B, max_blocks_per_seq] int32 (logical block -> physical block).  Same math as the dense entry points above:
 * prefill cache fill, decode RoPE + append (acc_f32_ws / bias optional as in b200_decode_rope_append_f32), decode attention
 * (tcgen05 kernel; each 128-row tile is gathered page by page with TMA; block_size 32, 64 or 128). */
int b200_write_cache_kv_paged(const void* qkv, void* key_cache, void* value_cache, const int32_t* block_tables,
                              const int32_t* seq_lens, int64_t B, int64_t S, int64_t num_heads, int64_t num_kv_heads,
                              int64_t head_dim, int64_t block_size, int64_t max_blocks_per_seq, int64_t ld, cudaStream_t stream);
int b200_decode_rope_append_paged(void* qkv, float* acc_f32_ws, const float* bias, void* key_cache, void* value_cache,
                                  const int32_t* block_tables, const float* cos_table, const float* sin_table,
                                  const int32_t* seq_lens, int64_t B, int64_t num_heads, int64_t num_kv_heads, int64_t head_dim,
                                  int64_t block_size, int64_t max_blocks_per_seq, int64_t ld, cudaStream_t stream);
int b200_decode_attention_paged(const void* qkv, const void* key_cache, const void* value_cache, const int32_t* block_tables,
                                const int32_t* seq_lens, void* out, void* workspace, int64_t B, int64_t num_heads,
                                int64_t num_kv_heads, int64_t head_dim, int64_t num_blocks, int64_t block_size,
                                int64_t max_blocks_per_seq, int64_t ld, float softmax_scale, int64_t num_splits,
                                cudaStream_t stream);

/* Bookkeeping ops, same semantics as the reference custom ops (file:line beside each). bool = 1-byte flags. */
/* get_padding_offset_v2 (+ remove padding): csrc/gpu/get_padding_offset_v2.cu:17-80 */
int b200_get_padding_offset(const int64_t* input_ids, const int32_t* cum_offsets, const int32_t* seq_lens,
                            int64_t* x_remove_padding, int32_t* padding_offset, int32_t* cum_offsets_out,
                            int32_t* cu_seqlens_q, int32_t* cu_seqlens_k, int64_t bsz, int64_t max_seq_len,
                            cudaStream_t stream);
/* rebuild_padding_v2: csrc/gpu/rebuild_padding_v2.cu:18-69 (one row per sequence = its last valid token) */
int b200_rebuild_padding(const void* tmp_out, const int32_t* cum_offsets, const int32_t* seq_lens_decoder,
                         const int32_t* seq_lens_encoder, void* out, int64_t bsz, int64_t max_len, int64_t dim,
                         cudaStream_t stream);
/* set_value_by_flags_and_idx: csrc/gpu/set_value_by_flags.cu:17-35 ; _v2: csrc/gpu/set_value_by_flags_v2.cu */
int b200_set_value_by_flags_and_idx(const bool* stop_flags, int64_t* pre_ids_all, const int64_t* pre_ids_now,
                                    const int64_t* step_idx, int64_t bs, int64_t length, cudaStream_t stream);
int b200_set_value_by_flags_and_idx_v2(const bool* stop_flags, int64_t* pre_ids_all, const int64_t* input_ids,
                                       const int32_t* seq_lens_encoder, const int32_t* seq_lens_decoder,
                                       const int64_t* step_idx, int64_t bs, int64_t length, int64_t length_input_ids,
                                       cudaStream_t stream);
/* get_token_penalty_multi_scores(_v2): csrc/gpu/token_penalty_multi_scores_v2.cu:19-139 (CPU twin
 * csrc/cpu/src/token_penalty_multi_scores.cc:18-85).  In place on fp32 logits [bs, length]; temperatures / bad_tokens may
 * be NULL (v1 op); workspace = bs*length int32. */
int b200_token_penalty_multi_scores(const int64_t* pre_ids, float* logits, const float* penalty_scores,
                                    const float* frequency_scores, const float* presence_scores, const float* temperatures,
                                    const int64_t* bad_tokens, const int64_t* cur_len, const int64_t* min_len,
                                    const int64_t* eos_token_id, int32_t* workspace, int64_t bs, int64_t length,
                                    int64_t length_id, int64_t bad_len, int64_t eos_len, cudaStream_t stream);
/* set_stop_value_multi_ends: v1 mode 2 csrc/gpu/stop_generation_multi_ends.cu:45-56 ; v2 …_v2.cu:35-59 */
int b200_set_stop_value_multi_ends(bool* stop_flags, int64_t* topk_ids, int64_t* next_tokens, const int64_t* end_ids,
                                   const int32_t* seq_lens, int64_t bs, int64_t end_length, int v2, cudaStream_t stream);
/* fused_get_rotary_embedding(input_ids, position_ids, head_dim_shape_tensor, prompt_num, theta, use_neox):
 * csrc/gpu/fused_get_rope.cu:40-223, called experimental/transformers/llama/modeling.py:799-803.
 * position_ids int64 [bsz, max_position_seq_length]; rope_embedding fp32 [2, bsz, 1, max_seq_length, head_dim] (cos, sin) with
 * angle = position_ids[b, s + prompt_num] * powf(theta, -2j/head_dim); use_neox != 0: value j at columns j and j + head_dim/2
 * (rotate-half, Llama/Qwen2), use_neox == 0: at columns 2j, 2j+1.  max_seq_length is input_ids.shape[1] in the reference. */
int b200_fused_get_rotary_embedding(const int64_t* position_ids, float* rope_embedding, int64_t bsz, int64_t max_seq_length,
                                    int64_t max_position_seq_length, int64_t head_dim, int64_t prompt_num, float theta,
                                    int use_neox, cudaStream_t stream);

/* step_paddle: csrc/gpu/step.cu:19-283 (free_and_dispatch_block + recover_block) — continuous-batching block bookkeeping of the
 * paged KV cache, called once per decode step: finished sequences return their decoder blocks to free_list; running sequences
 * that step into an unallocated block get one (pre-empting the largest holders into step_block_list when the list runs dry);
 * parked sequences are recovered when blocks are available again (lengths / stop flag / input_ids rebuilt from pre_ids).
 * All arguments are updated in place like the reference op (every tensor is an input aliased to an output there).  Sizes:
 * bsz = seq_lens_this_time.shape[0] (<= 1024), block_num_per_seq = block_tables.shape[1], length = input_ids.shape[1],
 * pre_id_length = pre_ids.shape[1]; the reference attribute `encoder_decoder_block_num` is unused by its kernels and omitted.
 * ONE launch, no host synchronisation (the reference copies recover_lens to the host between its two kernels); list order
 * is by sequence index (deterministic; the reference's atomics leave it timing dependent). */
int b200_step_paddle(bool* stop_flags, int32_t* seq_lens_this_time, const int32_t* ori_seq_lens_encoder,
                     int32_t* seq_lens_encoder, int32_t* seq_lens_decoder, int32_t* block_tables, int32_t* encoder_block_lens,
                     bool* is_block_step, int32_t* step_block_list, int32_t* step_lens, int32_t* recover_block_list,
                     int32_t* recover_lens, int32_t* need_block_list, int32_t* need_block_len, int32_t* used_list_len,
                     int32_t* free_list, int32_t* free_list_len, int64_t* input_ids, const int64_t* pre_ids,
                     const int64_t* step_idx, const int64_t* next_tokens, int64_t bsz, int64_t block_size,
                     int64_t block_num_per_seq, int64_t length, int64_t pre_id_length, int64_t first_token_id,
                     cudaStream_t stream);

/* save_output(x, not_need_stop, rank_id) replacement: csrc/gpu/save_with_output_msg.cc:28-52 (producer) / csrc/gpu/get_output.cc
 * (consumer), reader loop paddlenlp/utils/llm_utils.py:753-776.  Writes the step's message {flag, bsz, tokens...} (flag 1 =
 * running, -1 = finished: *stop_count >= bs or step >= last_step) into slot (step % num_slots) of `ring`, an int32 buffer of
 * num_slots x slot_stride in pinned device-mapped host memory, and publishes it by storing step + 1 into the slot's first word
 * last; `step_counter` (device int64) is the running step and is incremented.  No host synchronisation; graph-replayable. */
int b200_save_output_stream(const int64_t* next_tokens, const int32_t* stop_count, int32_t* ring, int64_t slot_stride,
                            int64_t num_slots, int64_t* step_counter, int64_t last_step, int64_t bs, cudaStream_t stream);

/* append_attention(qkv, key_cache, value_cache, seq_lens_encoder, seq_lens_decoder, seq_lens_this_time, padding_offsets,
 * cum_offsets, block_tables, ..., rotary_embs, ...): csrc/gpu/append_attention.cu:428-851, called
 * fused_transformer_layers.py:2215-2262 (FusedBlockMultiTransformer.compute_attn with config.append_attn).  ONE entry point for a
 * mixed batch over the paged KV cache: sequence b contributes seq_lens_this_time[b] rows of the packed projection
 * qkv [token_num, ldq] (rows cu_seqlens_q[b] ..), at absolute positions seq_lens_decoder[b] + i.  RoPE (rotate-half, tables
 * [rope_positions, 64] fp32) is applied to q and k in place, k and v are appended to the pages, and every row attends to cache
 * positions [0, its own]: prompts and prompt CHUNKS on top of a cached prefix (seq_lens_encoder[b] > 0 or more than one row)
 * through the tcgen05 flash kernel with page-gathered K/V, single decode rows through the decode kernel.  out [token_num, ldo].
 * max_q_len >= max(seq_lens_this_time) (host bound for the grid; the reference passes max_enc_len_this_time the same way).
 * cu_seqlens_q replaces padding_offsets / cum_offsets (same information; get_padding_offset produces it).
 * No host synchronisation.  Workspace: b200_append_attention_workspace_bytes(). */
int64_t b200_append_attention_workspace_bytes(int64_t B, int64_t num_heads, int64_t num_kv_heads, int64_t head_dim,
                                              int64_t num_splits);
int b200_append_attention(void* qkv, void* key_cache, void* value_cache, const int32_t* seq_lens_encoder,
                          const int32_t* seq_lens_decoder, const int32_t* seq_lens_this_time, const int32_t* cu_seqlens_q,
                          const int32_t* block_tables, const float* cos_table, const float* sin_table, void* out, void* workspace,
                          int64_t B, int64_t token_num, int64_t max_q_len, int64_t num_heads, int64_t num_kv_heads,
                          int64_t head_dim, int64_t num_blocks, int64_t block_size, int64_t max_blocks_per_seq,
                          int64_t rope_positions, int64_t ldq, int64_t ldo, float softmax_scale, int64_t num_splits,
                          cudaStream_t stream);

/* update_inputs: csrc/gpu/update_inputs.cu:18-82 */
int b200_update_inputs(bool* not_need_stop, int32_t* seq_lens_this_time, int32_t* seq_lens_encoder,
                       int32_t* seq_lens_decoder, int64_t* input_ids, const int64_t* stop_nums, const bool* stop_flags,
                       const bool* is_block_step, const int64_t* next_tokens, int64_t bsz, int64_t max_bsz,
                       int64_t input_ids_stride, cudaStream_t stream);
/* One fused state update per decode step of the dense-cache generate loop
 * (update_model_kwargs_for_generation, experimental/transformers/generation_utils.py:185-260): step_idx, length stop,
 * EOS stop, pre_ids history, seq_len_decoder, next/tgt ids, optional token log (column out_col, or *out_col_dev which is
 * then incremented on the device so that the step is CUDA-graph replayable), stop_count = number of stopped rows. */
int b200_generate_step_update(int64_t* next_tokens, bool* stop_flags, int64_t* step_idx, const int64_t* max_dec_len,
                              int32_t* seq_len_decoder, int64_t* pre_ids, int64_t pre_len, const int64_t* eos_ids,
                              int64_t eos_len, int64_t* out_tokens, int64_t out_stride, int64_t out_col,
                              int64_t* out_col_dev, int32_t* stop_count, int64_t bs, cudaStream_t stream);
int b200_argmax_f32(const float* logits, int64_t* out, int64_t rows, int64_t vocab, int64_t ld, cudaStream_t stream);
int b200_bf16_rows_to_f32(const void* src, float* dst, int64_t rows, int64_t cols, int64_t ld, cudaStream_t stream);
/* In-place fp32 softmax over each row of logits [rows, vocab] (row stride ld): `probs = F.softmax(logits)`,
 * experimental/transformers/generation_utils.py:327. */
int b200_softmax_f32(float* logits, int64_t rows, int64_t vocab, int64_t ld, cudaStream_t stream);
/* top_p_sampling_reject (csrc/gpu/sample_kernels/top_p_sampling_reject.cu:18-60, sampling.cuh:197-376): rejection sampling
 * of one token per row from probs [bs, vocab] restricted to the top-p nucleus; top_p [bs]; uniform [max_rounds, bs] are
 * the U(0,1) draws (the reference draws max_rounds = 32 per row from its generator inside the op); out [bs] int64.
 * top_p == 0 selects the arg max. */
int b200_top_p_sampling_reject(const float* probs, const float* top_p, const float* uniform, int64_t* out, int64_t bs,
                               int64_t vocab, int64_t ld, int64_t max_rounds, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200NLP_H_ */
