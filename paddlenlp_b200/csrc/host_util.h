// Host-side helpers shared by the C-ABI translation units: thread-local error string, CUtensorMap encoding
// through the driver entry point (no link-time dependency on libcuda), device attribute cache.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace b200 {

// Error convention of the C-ABI (include/b200nlp.h): 0 ok, <0 argument error, >0 cudaError_t.
void set_last_error(const char* fmt, ...);
int fail_arg(const char* fmt, ...);   // records message, returns -1
int check_launch(const char* what);  // cudaGetLastError() -> return code

int sm_count();  // multiprocessors of the current device (cached per device)
int skinny_gemm_impl();   // b200_set_skinny_gemm(): 1 = swapped-operand two-CTA/SM kernel (gemm_skinny.cu), 0 = the 128x256 persistent kernel
int fa_fwd_impl();         // b200_set_fa_fwd_impl(): 2 = two-q-tile kernel (fa_fwd2.cu), 1 = fa_fwd.cu
// fa_fwd2.cu: causal forward, optional FlashMask start rows (same argument meaning as b200_fa_fwd_flashmask)
int launch_fa_fwd2(const void* q, const void* k, const void* v, void* o, float* lse, const int32_t* mask_start_rows, int64_t B, int64_t S, int64_t num_heads,
                   int64_t num_kv_heads, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float softmax_scale,
                   cudaStream_t stream);
int l2_prefetch_mb();      // MB of weights a decode-step GEMM requests into L2 before griddepcontrol.wait (B200_L2_PREFETCH_MB, default 8)
int fa_exp_poly();        // b200_set_fa_exp_poly(): 0 / 1 / 2 = none / a quarter / half of the forward exponentials on the FMA pipe
int fa_bwd_impl();         // b200_set_fa_bwd_impl(): 2 = transposed pipelined kernel (fa_bwd2.cu), 1 = fa_bwd.cu
// fa_bwd2.cu: causal backward, optional FlashMask start rows (same argument meaning as b200_fa_bwd_flashmask)
int launch_fa_bwd2(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                   const int32_t* mask_start_rows, void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t S, int64_t num_heads, int64_t num_kv_heads,
                   int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv,
                   float softmax_scale, cudaStream_t stream);
// fa_fwd2.cu, PAGED instantiation: prefill half of b200_append_attention
int launch_fa_prefill_paged(const void* qkv, const void* key_cache, const void* value_cache, void* out, const int32_t* cu_seqlens_q,
                            const int32_t* seq_lens_encoder, const int32_t* seq_lens_decoder, const int32_t* seq_lens_this_time,
                            const int32_t* block_tables, int64_t B, int64_t token_num, int64_t max_q_len, int64_t num_heads,
                            int64_t num_kv_heads, int64_t num_blocks, int64_t block_size, int64_t max_blocks_per_seq, int64_t ldq,
                            int64_t ldo, float softmax_scale, cudaStream_t stream);
bool pdl_enabled();   // b200_set_pdl(): launch GEMMs with programmatic dependent launch (decode-step kernel chains)

// Launch `kern` on `stream`; when PDL is enabled the launch carries the programmatic-stream-serialization attribute, so the
// grid may become resident while its predecessor is still running.  Such kernels MUST call pdl_wait() (common.cuh) before
// touching memory the predecessor writes.  Returns the cudaLaunchKernelEx status (recorded by check_launch()).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// Encode a 2-D or 3-D bf16 (2-byte element) tiled tensor map with 128-byte swizzle.
//   dims[i]    extent of dimension i in elements (dimension 0 is contiguous)
//   strides[i] byte stride of dimension i+1 (i < rank-1); must be multiples of 16
//   box[i]     box extent in elements; box[0]*2 must be <= 128 for SWIZZLE_128B
// Returns 0 on success, <0 on failure (message recorded).
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box);
// bf16, no shared-memory swizzle: the box lands as plain rows of box[0] elements (small staging boxes read row-wise by one thread)
int encode_tmap_bf16_linear(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                            const uint32_t* box);
// Same for fp32 elements (box[0] * 4 must be <= 128); used for TMA reduce-add into fp32 accumulation buffers.
int encode_tmap_f32(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                    const uint32_t* box);

}  // namespace b200

#define B200_CHECK_ARG(cond, ...)                  \
  do {                                             \
    if (!(cond)) return b200::fail_arg(__VA_ARGS__); \
  } while (0)
