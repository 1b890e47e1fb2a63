// Causal GQA flash-attention forward on tcgen05 (head_dim 128).
//
//   O = softmax(Q K^T / sqrt(d) + causal) V ,  LSE saved for the backward.
//   q [b, s, nh, d], k/v [b, s, kvh, d] (arbitrary token stride: they are views into the packed QKV projection),
//   o [b, s, nh, d] contiguous, lse [b, nh, s] fp32 (natural log).
//
// Replaces F.scaled_dot_product_attention(is_causal=True) -> vendored FlashAttention-2 in the reference
// (paddlenlp/transformers/llama/fusion_ops.py:240-246; eager math llama/modeling.py:244-301).
// Rounding points: S and softmax in fp32 (scale applied to S), P rounded to bf16 before P@V, O rounded to bf16.
//
// One CTA = one (batch, q-head, 128-row q tile); kv tiles 0..i (square 128x128 causal tiles).
//   warp 0        TMA producer: Q once, K and V through 2-stage rings (128B swizzle)
//   warp 1        MMA issuer:   S[j&1] = Q K_j^T (UMMA 128x128x16 x8) ; O += P_j V_j (V consumed MN-major)
//   warps 2..9    softmax:      two threads per q row (TMEM lane), 64 score columns each (row max exchanged through
//                               smem); S read with tcgen05.ld, online softmax with lazy rescaling of the TMEM-resident
//                               O accumulator, P written to swizzled smem as bf16
//   TMEM: S0 [0,128) S1 [128,256) O [256,384).  QK_{j+1} is issued before P_j V_j so the tensor pipe works on the
//   next scores while the softmax warps exponentiate the current ones.
#include "../../include/b200nlp.h"
#include "common.cuh"
#include "host_util.h"

namespace b200 {
namespace fa {

constexpr int D = 128;       // head dim
constexpr int BQ = 128;      // q rows per CTA
constexpr int BKV = 128;     // kv rows per tile
constexpr int TILE_BYTES = 128 * 128 * 2;   // 32 KB (two 64-column halves of 16 KB)
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int NUM_THREADS = 320;   // TMA warp, MMA warp, 8 softmax warps
constexpr int SMEM_BYTES = 6 * TILE_BYTES + 256 + 3 * 1024 + 1024;   // Q, K0, K1, V0, V1, P + barriers + row-stat exchange + align slack
constexpr float RESCALE_THRESHOLD = 8.f;                  // log2 units

struct Params {
  int S, B, nh, kvh;
  float scale_log2;   // (1/sqrt(d)) * log2(e)
  float* lse;         // [B, nh, S]
  // FlashMask, causal lower-triangular form (fusion_ops.py:218-231 -> F.flashmask_attention(startend_row_indices, causal=True)):
  // mask_start[b, c] = first query row that may NOT see key column c (the end of c's packed document, llm/utils/data.py:
  // 200-204 + zero_padding_dataset.py:84-86); non-decreasing in c.  nullptr = plain causal.
  const int* mask_start;
};

template <bool MASK>     // MASK: FlashMask start rows present (kept out of the plain causal instantiation entirely)
__global__ void __launch_bounds__(NUM_THREADS, 1)
fa_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
              const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + TILE_BYTES;          // 2 stages
  uint8_t* sV = smem + 3 * TILE_BYTES;      // 2 stages
  uint8_t* sP = smem + 5 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * TILE_BYTES);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2]
  uint64_t* s_empty = bars + 11;    // [2]
  uint64_t* p_full = bars + 13;     // [1]
  uint64_t* pv_done = bars + 14;    // [1]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);
  float* s_stat = reinterpret_cast<float*>(smem + 6 * TILE_BYTES + 256);   // [3][2][128]: two max buffers + row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_q_tiles = (p.S + BQ - 1) / BQ;
  const int qt = num_q_tiles - 1 - static_cast<int>(blockIdx.x);   // heavy tiles first
  const int head = blockIdx.y, batch = blockIdx.z;
  const int kv_head = head / (p.nh / p.kvh);
  const int q0 = qt * BQ;
  // kv tiles j_lo .. qt.  With a document mask the leading tiles whose every column belongs to a document that ended at or
  // before this q tile are skipped (mask_start is non-decreasing, so they form a prefix; the diagonal tile is never empty).
  int j_lo = 0;
  if constexpr (MASK) {
    const int* ms = p.mask_start + static_cast<size_t>(batch) * p.S;
    while (j_lo < qt && __ldg(ms + min(j_lo * BKV + BKV - 1, p.S - 1)) <= q0) ++j_lo;
  }
  const int n_kv = qt + 1 - j_lo;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
    }
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tS0 = tmem_base, tO = tmem_base + 256;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_4d(&tmQ, q_full, sQ, 0, head, q0, batch);
      tma_load_4d(&tmQ, q_full, sQ + HALF_BYTES, 64, head, q0, batch);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_4d(&tmK, &k_full[st], sK + st * TILE_BYTES, 0, kv_head, (j_lo + j) * BKV, batch);
        tma_load_4d(&tmK, &k_full[st], sK + st * TILE_BYTES + HALF_BYTES, 64, kv_head, (j_lo + j) * BKV, batch);
        mbar_wait(&v_empty[st], ph ^ 1u);
        mbar_arrive_expect_tx(&v_full[st], TILE_BYTES);
        tma_load_4d(&tmV, &v_full[st], sV + st * TILE_BYTES, 0, kv_head, (j_lo + j) * BKV, batch);
        tma_load_4d(&tmV, &v_full[st], sV + st * TILE_BYTES + HALF_BYTES, 64, kv_head, (j_lo + j) * BKV, batch);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    // convergent code, one elected lane issues, descriptors advanced by adding byte offsets >> 4 (see fa_fwd2.cu: inside
    // `if (lane == 0)` every UTCHMMA costs ~19 SASS instructions of descriptor rebuilding and R2UR traffic)
    {
      const bool leader = elect_one();
      const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
      const uint32_t uS0 = tb, uO = tb + 256;
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);   // A=Q K-major, B=K K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, false, true);    // A=P K-major, B=V MN-major
      const uint32_t sK_a0 = smem_u32(sK), sV_a0 = smem_u32(sV);
      const uint64_t dQ = umma_desc_sw128(smem_u32(sQ), 16, 1024), dP = umma_desc_sw128(smem_u32(sP), 16, 1024);
      auto koff = [](int kk) { return static_cast<uint64_t>(((kk >> 2) * HALF_BYTES + (kk & 3) * 32) >> 4); };
      auto issue_qk = [&](int j) {
        const int st = j & 1;
        const uint64_t dK = umma_desc_sw128(sK_a0 + st * TILE_BYTES, 16, 1024);
        const uint32_t tS = uS0 + static_cast<uint32_t>((j & 1) * 128);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) umma_ss<1>(tS, dQ + koff(kk), dK + koff(kk), idesc_qk, kk > 0 ? 1u : 0u);
          umma_commit(&k_empty[st]);
          umma_commit(&s_full[j & 1]);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) {
          const int jn = j + 1;
          const uint32_t n = jn >> 1;   // use index of S buffer (jn & 1)
          mbar_wait(&s_empty[jn & 1], (n & 1u) ^ 1u);
          mbar_wait(&k_full[jn & 1], n & 1u);
          tc_fence_after();
          issue_qk(jn);
        }
        const int st = j & 1;
        mbar_wait(&v_full[st], (j >> 1) & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint64_t dV = umma_desc_sw128(sV_a0 + st * TILE_BYTES, HALF_BYTES, 1024);
        const uint32_t acc0 = j > 0 ? 1u : 0u;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk)     // P: kv along K (K-major);  V: 16 kv rows = 2 KB per k-step (MN-major)
            umma_ss<1>(uO, dP + koff(kk), dV + static_cast<uint64_t>(kk * 128), idesc_pv, kk > 0 ? 1u : acc0);
          umma_commit(&v_empty[st]);
          umma_commit(pv_done);
        }
      }
      __syncwarp();
    }
  } else {
    // ------------------------------- softmax / epilogue -------------------------------
    const int quad = warp & 3;                            // TMEM lane quadrant (warps w and w+4 share its rows)
    const int chalf = (warp - 2) >> 2;                    // score / output columns [64*chalf, 64*chalf + 64)
    const int r = quad * 32 + lane;                       // q row within the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    float m_used = -INFINITY, l = 0.f;                    // l: partial row sum over this thread's columns
    const uint32_t sP_a = smem_u32(sP);
    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sv[64];
      {
        uint32_t(*c)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        const uint32_t ta = tS0 + lane_off + static_cast<uint32_t>(sb * 128 + chalf * 64);
        tmem_ld32(ta, c[0]); tmem_ld32(ta + 32, c[1]);
        tmem_ld_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);
      // causal mask on the diagonal tile, row max over this thread's 64 columns (raw scores; the positive softmax
      // scale is folded into the exponent below as one FFMA per element)
      float rowmax = -INFINITY;
      const int jg = j_lo + j;                            // global kv tile index
      const bool diag = (jg == qt);
      if (diag) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if ((chalf * 64 + c) > r) sv[c] = 0xff800000u;   // -inf
      }
      if constexpr (MASK) {
        const int* ms = p.mask_start + static_cast<size_t>(batch) * p.S + jg * BKV;
        if (__ldg(ms) <= q0 + BQ - 1) {                   // some document in this kv tile ends inside / before the q tile
          const int row = q0 + r;
          const int cmax = p.S - jg * BKV - chalf * 64;   // columns of this half that exist
#pragma unroll                                            // full unroll: sv[] must stay in registers
          for (int c = 0; c < 64; ++c) {
            const int start = __ldg(ms + chalf * 64 + min(c, cmax - 1));
            if (c < cmax && row >= start) sv[c] = 0xff800000u;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 64; ++c) rowmax = fmaxf(rowmax, __uint_as_float(sv[c]));
      rowmax *= p.scale_log2;
      // exchange with the thread that owns the other 64 columns of this row
      s_stat[(sb * 2 + chalf) * 128 + r] = rowmax;
      named_bar_sync(2, 256);
      rowmax = fmaxf(rowmax, s_stat[(sb * 2 + (chalf ^ 1)) * 128 + r]);
      bool rescale = false;
      float factor = 1.f;
      if (j == 0) {
        m_used = rowmax;
      } else {
        const bool need = rowmax > m_used + RESCALE_THRESHOLD;
        rescale = __any_sync(0xffffffffu, need);          // identical in both warps of the row pair
        if (need) {
          factor = fast_exp2(m_used - rowmax);
          l *= factor;
          m_used = rowmax;
        }
      }
      uint32_t pk[32];
      float rs = 0.f;
      // with documents a row can be fully masked in its first tiles: exp2(-inf - (-inf)) must not be evaluated
      const float neg_m = (MASK && m_used == -INFINITY) ? 0.f : -m_used;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const float p0 = fast_exp2(fmaf(__uint_as_float(sv[2 * c]), p.scale_log2, neg_m));
        const float p1 = fast_exp2(fmaf(__uint_as_float(sv[2 * c + 1]), p.scale_log2, neg_m));
        rs += p0 + p1;
        pk[c] = pack_bf16x2(p0, p1);
      }
      l += rs;
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);   // P buffer free, O accumulator quiescent
        tc_fence_after();
        if (rescale) {                     // each thread rescales its 64 output columns
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t o[32];
            tmem_ld32(tO + lane_off + chalf * 64 + ch * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * factor);
            tmem_st32(tO + lane_off + chalf * 64 + ch * 32, o);
          }
          tmem_st_wait();
        }
      }
      // P -> swizzled smem (A operand, K-major along kv); this thread's 64 columns are exactly half `chalf`
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t addr = sP_a + chalf * HALF_BYTES + r * 128 + ((k ^ (r & 7)) << 4);
        st_shared_v4(addr, make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]));
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: O / l -> bf16 -> smem (reuse Q tile) -> TMA store ; LSE
    s_stat[(4 + chalf) * 128 + r] = l;
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    named_bar_sync(2, 256);
    l += s_stat[(4 + (chalf ^ 1)) * 128 + r];
    const float inv_l = 1.f / l;
    const uint32_t sO_a = smem_u32(sQ);
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t o[32];
      tmem_ld32(tO + lane_off + chalf * 64 + ch * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        const int k = ch * 4 + c8;     // 16-byte chunk within this thread's 128-byte half row
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[c8 * 8 + 0]) * inv_l, __uint_as_float(o[c8 * 8 + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o[c8 * 8 + 2]) * inv_l, __uint_as_float(o[c8 * 8 + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o[c8 * 8 + 4]) * inv_l, __uint_as_float(o[c8 * 8 + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o[c8 * 8 + 6]) * inv_l, __uint_as_float(o[c8 * 8 + 7]) * inv_l);
        st_shared_v4(sO_a + chalf * HALF_BYTES + r * 128 + ((k ^ (r & 7)) << 4), v);
      }
    }
    if (chalf == 0 && q0 + r < p.S)
      p.lse[(static_cast<size_t>(batch) * p.nh + head) * p.S + q0 + r] = (m_used + log2f(l)) * 0.6931471805599453f;
    fence_proxy_async_smem();
    named_bar_sync(1, 256);
    if (warp == 2 && lane == 0) {
      tma_store_4d(&tmO, sQ, 0, head, q0, batch);
      tma_store_4d(&tmO, sQ + HALF_BYTES, 64, head, q0, batch);
      tma_store_commit();
      tma_store_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// 4-D map over a [B, S, heads, 128] bf16 view with token stride `ld` (elements): dims {128, heads, S, B}.
static int make_map(CUtensorMap* tm, const void* base, int64_t B, int64_t S, int64_t heads, int64_t ld) {
  uint64_t dims[4] = {128, static_cast<uint64_t>(heads), static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {128 * 2, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(S) * ld * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return encode_tmap_bf16(tm, base, 4, dims, strides, box);
}

}  // namespace fa
}  // namespace b200

extern "C" int b200_fa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int64_t B, int64_t S,
                           int64_t num_heads, int64_t num_kv_heads, int64_t head_dim, int64_t ldq, int64_t ldk,
                           int64_t ldv, int64_t ldo, float softmax_scale, cudaStream_t stream) {
  return b200_fa_fwd_flashmask(q, k, v, o, lse, nullptr, B, S, num_heads, num_kv_heads, head_dim, ldq, ldk, ldv, ldo,
                               softmax_scale, stream);
}

extern "C" int b200_fa_fwd_flashmask(const void* q, const void* k, const void* v, void* o, float* lse,
                                     const int32_t* mask_start_rows, int64_t B, int64_t S, int64_t num_heads,
                                     int64_t num_kv_heads, int64_t head_dim, int64_t ldq, int64_t ldk, int64_t ldv,
                                     int64_t ldo, float softmax_scale, cudaStream_t stream) {
  using namespace b200;
  using namespace b200::fa;
  B200_CHECK_ARG(q && k && v && o && lse, "fa_fwd: null pointer");
  B200_CHECK_ARG(head_dim == 128, "fa_fwd: head_dim must be 128 (got %lld)", (long long)head_dim);
  B200_CHECK_ARG(B > 0 && S > 0 && num_heads > 0 && num_kv_heads > 0 && num_heads % num_kv_heads == 0,
                 "fa_fwd: bad shape B=%lld S=%lld nh=%lld kvh=%lld", (long long)B, (long long)S, (long long)num_heads,
                 (long long)num_kv_heads);
  B200_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "fa_fwd: token strides must be multiples of 8");
  if (fa_fwd_impl() == 2)     // two q tiles per CTA, P in TMEM (fa_fwd2.cu); plain causal or FlashMask start rows
    return launch_fa_fwd2(q, k, v, o, lse, mask_start_rows, B, S, num_heads, num_kv_heads, ldq, ldk, ldv, ldo, softmax_scale, stream);
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = make_map(&tmQ, q, B, S, num_heads, ldq)) != 0) return rc;
  if ((rc = make_map(&tmK, k, B, S, num_kv_heads, ldk)) != 0) return rc;
  if ((rc = make_map(&tmV, v, B, S, num_kv_heads, ldv)) != 0) return rc;
  if ((rc = make_map(&tmO, o, B, S, num_heads, ldo)) != 0) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(fa_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("fa_fwd smem attr: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  Params p;
  p.S = static_cast<int>(S); p.B = static_cast<int>(B); p.nh = static_cast<int>(num_heads);
  p.kvh = static_cast<int>(num_kv_heads);
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.lse = lse;
  p.mask_start = mask_start_rows;
  dim3 grid(static_cast<unsigned>((S + BQ - 1) / BQ), static_cast<unsigned>(num_heads), static_cast<unsigned>(B));
  if (mask_start_rows != nullptr) fa_fwd_kernel<true><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  else fa_fwd_kernel<false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  return check_launch("fa_fwd");
}
