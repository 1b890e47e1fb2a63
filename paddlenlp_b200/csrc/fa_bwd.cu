// Causal GQA flash-attention backward on tcgen05 (head_dim 128).
//
//   inputs : q, k, v, o, do, lse            outputs: dq (fp32 accumulation buffer), dk, dv (bf16)
//   P = exp(S*scale - lse) ; dP = dO V^T ; dS = P o (dP - D) * scale, D = rowsum(dO o O)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K
//
// Replaces Paddle-core flash_attn_grad (reference: fusion_ops.py:240-246 backward of scaled_dot_product_attention;
// wrapper shape in csrc/gpu/flash_attn_bwd.cc:22-92).
//
// One CTA = one (batch, q-head, 128-row kv tile); it loops over the q tiles i >= j, accumulating this head's dK/dV in
// TMEM.  Work units are per q-head (not per kv-head) so that the 4096 units of a Llama-3 micro-batch balance over
// 148 SMs (heaviest first); the GQA group's dK/dV partials and the dQ tiles are reduced into fp32 buffers by the TMA
// unit (cp.reduce.async.bulk.tensor .add — no LSU atomics), then converted to bf16 by a finishing kernel.
//   warp 0       TMA producer (K_j, V_j once; Q_i, dO_i per iteration)
//   warp 1       MMA issuer   (5 UMMA GEMMs per iteration, operands K-major or MN-major straight from the same
//                              swizzled tiles: Q and dO are consumed both ways)
//   warps 2..9   two threads per q row (64 columns each; the backward needs no row reduction): P and dS from TMEM
//                S / dP, written as bf16 to swizzled smem; dQ / dK / dV read-out
//   TMEM: S [0,128)  dP|dQ [128,256)  dV [256,384)  dK [384,512).  dQ reuses the dP columns, so S_{i+1} = Q_{i+1} K^T is
//   issued while the dQ_i tile is still being read out, and the Q/dO stage is released before the dQ GEMM is issued.
#include "../../include/b200nlp.h"
#include "common.cuh"
#include "host_util.h"

namespace b200 {
namespace fab {

constexpr int TILE_BYTES = 128 * 128 * 2;
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int NUM_THREADS = 320;   // TMA warp, MMA warp, 8 compute warps
constexpr int STAGE_BYTES = 8 * 4096;                     // per-warp 32x32 fp32 staging for TMA reduce
constexpr int SMEM_BYTES = 6 * TILE_BYTES + STAGE_BYTES + 256 + 1024;   // K, V, Q, dO, P, dS, staging

struct Params {
  int S, B, nh, kvh;
  float scale, scale_log2;
  const float* lse;     // [B, nh, S]  natural log
  const float* delta;   // [B, nh, S]  rowsum(dO o O)
  const int* mask_start;   // FlashMask causal-LT start rows [B, S] (see fa_fwd.cu) or nullptr
};

// TMEM accumulator (this warp's 32 lanes, fp32 columns [32*ch0, 32*(ch0+nch))) -> fp32 staging -> TMA reduce-add of
// 32x32 boxes.  One 4 KB staging buffer per warp: the previous reduce must have finished reading it.
__device__ __forceinline__ void reduce_out_tile(uint32_t tsrc, uint8_t* buf, const CUtensorMap* tm, int lane, int head,
                                                int row0, int batch, int ch0) {
  // both 32-column chunks are fetched from TMEM before the single wait (one tcgen05.ld round trip instead of two)
  uint32_t o[2][32];
  tmem_ld32(tsrc + ch0 * 32, o[0]);
  tmem_ld32(tsrc + (ch0 + 1) * 32, o[1]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
    const uint32_t row_s = smem_u32(buf) + lane * 128;
#pragma unroll
    for (int c = 0; c < 8; ++c)
      st_shared_v4(row_s + ((c ^ (lane & 7)) << 4), make_uint4(o[i][4 * c], o[i][4 * c + 1], o[i][4 * c + 2], o[i][4 * c + 3]));
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      tma_reduce_add_4d(tm, buf, (ch0 + i) * 32, head, row0, batch);
      tma_store_commit();
    }
  }
}

template <bool MASK>     // MASK: FlashMask start rows present (kept out of the plain causal instantiation entirely)
__global__ void __launch_bounds__(NUM_THREADS, 1)
fa_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
              const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdK,
              const __grid_constant__ CUtensorMap tmdV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + TILE_BYTES;
  uint8_t* sQ = smem + 2 * TILE_BYTES;
  uint8_t* sdO = smem + 3 * TILE_BYTES;
  uint8_t* sP = smem + 4 * TILE_BYTES;
  uint8_t* sdS = smem + 5 * TILE_BYTES;
  uint8_t* sStage = smem + 6 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + STAGE_BYTES);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;
  uint64_t* qdo_empty = bars + 2;
  uint64_t* s_full = bars + 3;
  uint64_t* pds_full = bars + 4;
  uint64_t* dq_full = bars + 5;
  uint64_t* dq_empty = bars + 6;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = (p.S + 127) / 128;
  const int jt = static_cast<int>(blockIdx.x);     // kv tile; tile 0 has the most work and is scheduled first
  const int hq = blockIdx.y, batch = blockIdx.z;
  const int kv_head = hq / (p.nh / p.kvh);
  const int kv0 = jt * 128;
  // q tiles jt .. hi-1: with a document mask, q tiles that start at or after the end of the last document of this kv tile
  // see none of its columns (mask_start is non-decreasing; the diagonal tile always remains)
  int hi = num_tiles;
  if constexpr (MASK)
    hi = min(num_tiles, (__ldg(p.mask_start + static_cast<size_t>(batch) * p.S + min(kv0 + 127, p.S - 1)) + 127) / 128);
  const int n_iter = hi - jt;
  __shared__ int s_start[MASK ? 128 : 1];          // mask start row of each column of this kv tile
  if constexpr (MASK) {
    if (threadIdx.x < 128) {
      const int c = kv0 + static_cast<int>(threadIdx.x);
      s_start[threadIdx.x] = c < p.S ? p.mask_start[static_cast<size_t>(batch) * p.S + c] : 0x7fffffff;
    }
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1);
    mbar_init(qdo_full, 1);
    mbar_init(qdo_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(pds_full, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_empty, 256);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_4d(&tmK, kv_full, sK, 0, kv_head, kv0, batch);
      tma_load_4d(&tmK, kv_full, sK + HALF_BYTES, 64, kv_head, kv0, batch);
      tma_load_4d(&tmV, kv_full, sV, 0, kv_head, kv0, batch);
      tma_load_4d(&tmV, kv_full, sV + HALF_BYTES, 64, kv_head, kv0, batch);
      for (int n = 0; n < n_iter; ++n) {
        const int q0 = (jt + n) * 128;
        mbar_wait(qdo_empty, (n & 1) ^ 1u);
        mbar_arrive_expect_tx(qdo_full, 2 * TILE_BYTES);
        tma_load_4d(&tmQ, qdo_full, sQ, 0, hq, q0, batch);
        tma_load_4d(&tmQ, qdo_full, sQ + HALF_BYTES, 64, hq, q0, batch);
        tma_load_4d(&tmdO, qdo_full, sdO, 0, hq, q0, batch);
        tma_load_4d(&tmdO, qdo_full, sdO + HALF_BYTES, 64, hq, q0, batch);
      }
    }
  } else if (warp == 1) {
    // MMA issuer: convergent code, one elected lane issues, descriptors advanced by byte offsets >> 4 (see fa_bwd2.cu)
    {
      const bool leader = elect_one();
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t uS = tb, udP = tb + 128, udV = tb + 256, udK = tb + 384;
      constexpr uint32_t id_kk = umma_idesc_bf16(128, 128, false, false);
      constexpr uint32_t id_mm = umma_idesc_bf16(128, 128, true, true);
      constexpr uint32_t id_km = umma_idesc_bf16(128, 128, false, true);
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ = smem_u32(sQ), adO = smem_u32(sdO), aP = smem_u32(sP),
                     adS = smem_u32(sdS);
      // K-major operand (k-step = 16 of the contiguous dim) / MN-major operand (k-step = 16 rows = 2 KB) base descriptors
      const uint64_t kQ = umma_desc_sw128(aQ, 16, 1024), kK = umma_desc_sw128(aK, 16, 1024), kdO = umma_desc_sw128(adO, 16, 1024),
                     kV = umma_desc_sw128(aV, 16, 1024), kdS = umma_desc_sw128(adS, 16, 1024);
      const uint64_t mP = umma_desc_sw128(aP, HALF_BYTES, 1024), mdO = umma_desc_sw128(adO, HALF_BYTES, 1024),
                     mdS = umma_desc_sw128(adS, HALF_BYTES, 1024), mQ = umma_desc_sw128(aQ, HALF_BYTES, 1024),
                     mK = umma_desc_sw128(aK, HALF_BYTES, 1024);
      auto koff = [](int kk) { return static_cast<uint64_t>(((kk >> 2) * HALF_BYTES + (kk & 3) * 32) >> 4); };
      auto moff = [](int kk) { return static_cast<uint64_t>(kk * 128); };
      mbar_wait(kv_full, 0);
      for (int n = 0; n < n_iter; ++n) {
        mbar_wait(qdo_full, n & 1);
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss<1>(uS, kQ + koff(kk), kK + koff(kk), id_kk, kk > 0);      // S = Q K^T
        }
        mbar_wait(dq_empty, (n & 1) ^ 1u);     // dP|dQ columns drained by the previous iteration's dQ read-out
        tc_fence_after();
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss<1>(udP, kdO + koff(kk), kV + koff(kk), id_kk, kk > 0);    // dP = dO V^T
          umma_commit(s_full);
        }
        mbar_wait(pds_full, n & 1);
        tc_fence_after();
        const uint32_t acc0 = n > 0 ? 1u : 0u;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss<1>(udV, mP + moff(kk), mdO + moff(kk), id_mm, kk > 0 ? 1u : acc0);   // dV += P^T dO
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss<1>(udK, mdS + moff(kk), mQ + moff(kk), id_mm, kk > 0 ? 1u : acc0);   // dK += dS^T Q
          umma_commit(qdo_empty);                // Q / dO are not read by the dQ GEMM: the next tiles can be loaded now
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss<1>(udP, kdS + koff(kk), mK + moff(kk), id_km, kk > 0);    // dQ = dS K
          umma_commit(dq_full);
        }
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;                 // TMEM lane quadrant (warps w and w+4 share the rows of a quadrant)
    const int chalf = (warp - 2) >> 2;         // which 64 of the 128 tile columns this thread handles
    const int r = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t aP = smem_u32(sP), adS = smem_u32(sdS);
    uint8_t* my_stage = sStage + (warp - 2) * 4096;
    // row statistics of the NEXT q tile are fetched one iteration ahead (their global-load latency used to sit at the top of
    // every iteration, in front of the s_full wait)
    const size_t stat_base = (static_cast<size_t>(batch) * p.nh + hq) * p.S;
    auto load_stats = [&](int q0n, float& l, float& d) {
      const bool ok = (q0n + r) < p.S;
      l = ok ? __ldg(p.lse + stat_base + q0n + r) : 0.f;
      d = ok ? __ldg(p.delta + stat_base + q0n + r) : 0.f;
    };
    float lse_next, drow_next;
    load_stats(jt * 128, lse_next, drow_next);
    for (int n = 0; n < n_iter; ++n) {
      const int qt = jt + n;
      const int q0 = qt * 128;
      const bool row_ok = (q0 + r) < p.S;
      const float lse2 = lse_next * 1.4426950408889634f;
      const float drow = drow_next;
      if (n + 1 < n_iter) load_stats(q0 + 128, lse_next, drow_next);
      const bool diag = (qt == jt);
      const bool mtile = MASK && (q0 + 127 >= s_start[0]);   // some column's document ends in / before this q tile
      const int qrow = q0 + r;
      mbar_wait(s_full, n & 1);
      tc_fence_after();
#pragma unroll
      for (int ch = chalf * 2; ch < chalf * 2 + 2; ++ch) {
        uint32_t sv[32], dv[32];
        tmem_ld32(tS + lane_off + ch * 32, sv);
        tmem_ld32(tdP + lane_off + ch * 32, dv);
        tmem_ld_wait();
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int col = ch * 32 + 2 * c;
          float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * c]), p.scale_log2, -lse2));
          float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * c + 1]), p.scale_log2, -lse2));
          if (!row_ok || (diag && col > r)) p0 = 0.f;
          if (!row_ok || (diag && col + 1 > r)) p1 = 0.f;
          if constexpr (MASK) {
            if (mtile && qrow >= s_start[col]) p0 = 0.f;
            if (mtile && qrow >= s_start[col + 1]) p1 = 0.f;
          }
          const float d0 = p0 * (__uint_as_float(dv[2 * c]) - drow) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv[2 * c + 1]) - drow) * p.scale;
          pp[c] = pack_bf16x2(p0, p1);
          dd[c] = pack_bf16x2(d0, d1);
        }
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int c16 = ch * 4 + k4;
          const uint32_t off = (c16 >> 3) * HALF_BYTES + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
          st_shared_v4(aP + off, make_uint4(pp[4 * k4], pp[4 * k4 + 1], pp[4 * k4 + 2], pp[4 * k4 + 3]));
          st_shared_v4(adS + off, make_uint4(dd[4 * k4], dd[4 * k4 + 1], dd[4 * k4 + 2], dd[4 * k4 + 3]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
      // dQ tile read-out: TMEM -> fp32 staging -> TMA reduce-add into the fp32 dQ buffer (rows >= S are clipped)
      mbar_wait(dq_full, n & 1);
      tc_fence_after();
      reduce_out_tile(tdP + lane_off, my_stage, &tmdQ, lane, hq, q0 + quad * 32, batch, chalf * 2);
      tc_fence_before();
      mbar_arrive(dq_empty);
    }
    // epilogue: this head's dK / dV partials -> fp32 reduce-add (the GQA group's heads sum in L2)
    reduce_out_tile(tdK + lane_off, my_stage, &tmdK, lane, kv_head, kv0 + quad * 32, batch, chalf * 2);
    reduce_out_tile(tdV + lane_off, my_stage, &tmdV, lane, kv_head, kv0 + quad * 32, batch, chalf * 2);
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// delta[b, h, s] = sum_d dO[b,s,h,d] * O[b,s,h,d]      (16 lanes per row of 128)
__global__ void fa_bwd_delta_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, float* __restrict__ delta,
                                    int B, int S, int nh, int64_t ldo, int64_t lddo) {
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 4;   // (b, s, h) flattened
  const int sub = threadIdx.x & 15;
  const int64_t total = static_cast<int64_t>(B) * S * nh;
  float acc = 0.f;
  int b = 0, s = 0, h = 0;
  if (row < total) {
    h = static_cast<int>(row % nh);
    const int64_t tok = row / nh;
    s = static_cast<int>(tok % S);
    b = static_cast<int>(tok / S);
    const uint4 ov = ld_nc_v4(reinterpret_cast<const uint4*>(o + tok * ldo + h * 128) + sub);
    const uint4 dv = ld_nc_v4(reinterpret_cast<const uint4*>(dout + tok * lddo + h * 128) + sub);
    const uint32_t* oi = reinterpret_cast<const uint32_t*>(&ov);
    const uint32_t* di = reinterpret_cast<const uint32_t*>(&dv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = unpack_bf16x2(oi[j]), d = unpack_bf16x2(di[j]);
      acc += a.x * d.x + a.y * d.y;
    }
  }
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (row < total && sub == 0) delta[(static_cast<size_t>(b) * nh + h) * S + s] = acc;
}

// out (bf16, token stride ld) = bf16(acc fp32 [tokens, width])
__global__ void fa_bwd_dq_finish_kernel(const float* __restrict__ acc, bf16* __restrict__ dq, int64_t tokens, int width,
                                        int64_t lddq) {
  const int64_t nchunk_row = width >> 3;
  const int64_t total = tokens * nchunk_row;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t t = i / nchunk_row, c = i % nchunk_row;
    const float4* src = reinterpret_cast<const float4*>(acc + t * width) + 2 * c;
    const float4 a = src[0], b = src[1];
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    *(reinterpret_cast<uint4*>(dq + t * lddq) + c) = o;
  }
}

static int make_map(CUtensorMap* tm, const void* base, int64_t B, int64_t S, int64_t heads, int64_t ld) {
  uint64_t dims[4] = {128, static_cast<uint64_t>(heads), static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {128 * 2, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(S) * ld * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return encode_tmap_bf16(tm, base, 4, dims, strides, box);
}
// fp32 accumulation buffer [B, S, heads, 128] contiguous; 32x32 boxes for the per-warp TMA reduce-add.
static int make_acc_map(CUtensorMap* tm, const void* base, int64_t B, int64_t S, int64_t heads) {
  uint64_t dims[4] = {128, static_cast<uint64_t>(heads), static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {128 * 4, static_cast<uint64_t>(heads) * 128 * 4, static_cast<uint64_t>(S) * heads * 128 * 4};
  uint32_t box[4] = {32, 1, 32, 1};
  return encode_tmap_f32(tm, base, 4, dims, strides, box);
}

}  // namespace fab
}  // namespace b200

extern "C" int64_t b200_fa_bwd_workspace_bytes(int64_t B, int64_t S, int64_t num_heads, int64_t head_dim) {
  // fp32 dQ accumulation buffer + fp32 dK/dV accumulation buffers (at most num_heads wide) + per-row statistics; sized for both
  // kernel generations: fa_bwd2.cu pads the sequence to a multiple of 64 and keeps two floats of statistics per row
  const int64_t Spad = (S + 63) / 64 * 64;
  return 3 * B * Spad * num_heads * head_dim * 4 + B * num_heads * Spad * 8;
}

extern "C" int b200_fa_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                           void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t S, int64_t num_heads,
                           int64_t num_kv_heads, int64_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo,
                           int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv, float softmax_scale,
                           cudaStream_t stream) {
  return b200_fa_bwd_flashmask(q, k, v, o, dout, lse, nullptr, dq, dk, dv, workspace, B, S, num_heads, num_kv_heads, head_dim,
                               ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv, softmax_scale, stream);
}

extern "C" int b200_fa_bwd_flashmask(const void* q, const void* k, const void* v, const void* o, const void* dout,
                                     const float* lse, const int32_t* mask_start_rows, void* dq, void* dk, void* dv,
                                     void* workspace, int64_t B, int64_t S, int64_t num_heads, int64_t num_kv_heads,
                                     int64_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo,
                                     int64_t lddq, int64_t lddk, int64_t lddv, float softmax_scale, cudaStream_t stream) {
  using namespace b200;
  using namespace b200::fab;
  B200_CHECK_ARG(q && k && v && o && dout && lse && dq && dk && dv && workspace, "fa_bwd: null pointer");
  B200_CHECK_ARG(head_dim == 128, "fa_bwd: head_dim must be 128 (got %lld)", (long long)head_dim);
  B200_CHECK_ARG(B > 0 && S > 0 && num_heads > 0 && num_kv_heads > 0 && num_heads % num_kv_heads == 0, "fa_bwd: bad shape");
  B200_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && lddo % 8 == 0 && lddq % 8 == 0 &&
                     lddk % 8 == 0 && lddv % 8 == 0,
                 "fa_bwd: token strides must be multiples of 8");
  if (fa_bwd_impl() == 2)    // transposed, software-pipelined kernel (fa_bwd2.cu); plain causal or FlashMask start rows
    return launch_fa_bwd2(q, k, v, o, dout, lse, mask_start_rows, dq, dk, dv, workspace, B, S, num_heads, num_kv_heads, ldq, ldk, ldv, ldo, lddo,
                          lddq, lddk, lddv, softmax_scale, stream);
  float* dq_acc = static_cast<float*>(workspace);
  float* dk_acc = dq_acc + B * S * num_heads * 128;
  float* dv_acc = dk_acc + B * S * num_kv_heads * 128;
  float* delta = dq_acc + 3 * B * S * num_heads * 128;
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(B) * S * (num_heads + 2 * num_kv_heads) * 128 * 4, stream);
  if (e != cudaSuccess) {
    set_last_error("fa_bwd memset: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  {
    const int64_t rows = B * S * num_heads;
    const int64_t threads = rows * 16;
    fa_bwd_delta_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(
        static_cast<const bf16*>(o), static_cast<const bf16*>(dout), delta, (int)B, (int)S, (int)num_heads, ldo, lddo);
    int rc = check_launch("fa_bwd(delta)");
    if (rc) return rc;
  }
  CUtensorMap tmQ, tmK, tmV, tmdO, tmdQ, tmdK, tmdV;
  int rc;
  if ((rc = make_map(&tmQ, q, B, S, num_heads, ldq)) != 0) return rc;
  if ((rc = make_map(&tmK, k, B, S, num_kv_heads, ldk)) != 0) return rc;
  if ((rc = make_map(&tmV, v, B, S, num_kv_heads, ldv)) != 0) return rc;
  if ((rc = make_map(&tmdO, dout, B, S, num_heads, lddo)) != 0) return rc;
  if ((rc = make_acc_map(&tmdQ, dq_acc, B, S, num_heads)) != 0) return rc;
  if ((rc = make_acc_map(&tmdK, dk_acc, B, S, num_kv_heads)) != 0) return rc;
  if ((rc = make_acc_map(&tmdV, dv_acc, B, S, num_kv_heads)) != 0) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(fa_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fa_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("fa_bwd smem attr: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  Params p;
  p.S = (int)S; p.B = (int)B; p.nh = (int)num_heads; p.kvh = (int)num_kv_heads;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.lse = lse; p.delta = delta;
  p.mask_start = mask_start_rows;
  dim3 grid(static_cast<unsigned>((S + 127) / 128), static_cast<unsigned>(num_heads), static_cast<unsigned>(B));
  if (mask_start_rows != nullptr)
    fa_bwd_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmdO, tmdQ, tmdK, tmdV, p);
  else
    fa_bwd_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmdO, tmdQ, tmdK, tmdV, p);
  if ((rc = check_launch("fa_bwd")) != 0) return rc;
  {
    const int64_t tokens = B * S;
    const int width = static_cast<int>(num_heads * 128);
    const int64_t total = tokens * (width / 8);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (blocks > cap) blocks = cap;
    fa_bwd_dq_finish_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(dq_acc, static_cast<bf16*>(dq), tokens,
                                                                              width, lddq);
    if ((rc = check_launch("fa_bwd(dq finish)")) != 0) return rc;
    const int kvw = static_cast<int>(num_kv_heads * 128);
    int64_t kblocks = (tokens * (kvw / 8) + 255) / 256;
    if (kblocks > cap) kblocks = cap;
    fa_bwd_dq_finish_kernel<<<static_cast<unsigned>(kblocks), 256, 0, stream>>>(dk_acc, static_cast<bf16*>(dk), tokens, kvw,
                                                                               lddk);
    if ((rc = check_launch("fa_bwd(dk finish)")) != 0) return rc;
    fa_bwd_dq_finish_kernel<<<static_cast<unsigned>(kblocks), 256, 0, stream>>>(dv_acc, static_cast<bf16*>(dv), tokens, kvw,
                                                                               lddv);
    rc = check_launch("fa_bwd(dv finish)");
  }
  return rc;
}
