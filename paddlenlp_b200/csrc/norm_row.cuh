// One row of the decode step's fused residual-add + RMSNorm, executed by 128 threads (4 warps) of a CTA:
//   res_out[row] = bf16(bf16(x_f32[row]) + res_in[row]) ; normed[row] = bf16(bf16(res_out * rstd) * w) ; x_f32[row] = 0
// Same rounding points and summation order as add_rmsnorm_kernel<.., 4> (generation.cu; fused_transformer_layers.py:937-999): thread
// et owns the 8-element chunks et, et + 128, ...; per-thread partial, warp tree, then the four warp partials in order.
// Shared by the GEMM kernels that finish a split-K projection with the norm that consumes it (gemm_skinny.cu NORM tail,
// decode_chain.cu).  x_f32 is read with ld.global.cg: it was written by other SMs' TMA reduce-adds during this kernel.
#pragma once
#include "common.cuh"

namespace b200 {

// s_part: 4 floats of shared memory; bar_id: a named barrier of the 128 participating threads (et = 0..127, warp-aligned)
__device__ __forceinline__ void add_rmsnorm_row_128(float* __restrict__ x_f32, const bf16* __restrict__ res_in, bf16* __restrict__ res_out,
                                                    const bf16* __restrict__ w, bf16* __restrict__ normed, int row, int h, float eps,
                                                    int et, float* s_part, uint32_t bar_id) {
  const int nchunk = h >> 3;
  float4* xf = reinterpret_cast<float4*>(x_f32 + static_cast<size_t>(row) * h);
  const uint4* rr = res_in ? reinterpret_cast<const uint4*>(res_in + static_cast<size_t>(row) * h) : nullptr;
  uint4* ro = res_out ? reinterpret_cast<uint4*>(res_out + static_cast<size_t>(row) * h) : nullptr;
  float ss = 0.f;
  constexpr int MAXV = 8;                    // h <= 8192
  uint4 v[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = et + 128 * i;
    if (c < nchunk) {
      const float4 a = __ldcg(xf + 2 * c), b = __ldcg(xf + 2 * c + 1);
      v[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
      uint32_t* vi = reinterpret_cast<uint32_t*>(&v[i]);
      if (rr) {
        const uint4 r = __ldcg(rr + c);
        const uint32_t* ri = reinterpret_cast<const uint32_t*>(&r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16x2(vi[j]), y = unpack_bf16x2(ri[j]);
          vi[j] = pack_bf16x2(x.x + y.x, x.y + y.y);
        }
      }
      float part = 0.f;                        // same expression shape as add_rmsnorm_kernel: identical fp32 contraction / order
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 a2 = unpack_bf16x2(vi[j]); part += a2.x * a2.x + a2.y * a2.y; }
      ss += part;
      xf[2 * c] = make_float4(0.f, 0.f, 0.f, 0.f);
      xf[2 * c + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ro) ro[c] = v[i];
    }
  }
  ss = warp_sum(ss);
  if ((et & 31) == 0) s_part[et >> 5] = ss;
  named_bar_sync(bar_id, 128);
  ss = s_part[0] + s_part[1] + s_part[2] + s_part[3];
  if (normed != nullptr) {
    const float rstd = rsqrtf(ss / static_cast<float>(h) + eps);
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    uint4* yr = reinterpret_cast<uint4*>(normed + static_cast<size_t>(row) * h);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int c = et + 128 * i;
      if (c < nchunk) {
        const uint4 wv = __ldg(wr + c);
        uint4 o;
        const uint32_t* xi = reinterpret_cast<const uint32_t*>(&v[i]);
        const uint32_t* wi = reinterpret_cast<const uint32_t*>(&wv);
        uint32_t* oi = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16x2(xi[j]), w2 = unpack_bf16x2(wi[j]);
          oi[j] = pack_bf16x2(bf16_round(x.x * rstd) * w2.x, bf16_round(x.y * rstd) * w2.y);
        }
        yr[c] = o;
      }
    }
  }
  named_bar_sync(bar_id, 128);               // s_part may be rewritten by the next row
}

}  // namespace b200
