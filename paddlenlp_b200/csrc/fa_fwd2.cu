// Causal GQA flash-attention forward, second generation (head_dim 128): two 128-row q tiles per CTA.
//
//   O = softmax(Q K^T / sqrt(d) + causal) V ,  LSE saved for the backward.   Same contract as fa_fwd.cu.
//
// Why a second kernel.  The first one (fa_fwd.cu: one q tile per CTA, two threads per row, P staged through shared
// memory) spends ~3800 cycles per 128x128 kv tile against 1024 cycles of tensor work (ncu: tensor pipe 26.7 %): the 8
// softmax warps run in lock-step (named barrier for the row-max exchange), so the MUFU phase (16 ex2/clk/SM = 1024
// cycles per tile) never overlaps the load / max / pack / store phases, and the MMA pipe idles while they run.
// Here
//   * a CTA owns TWO q tiles A and B (256 q rows of one (batch, head)); each has its own softmax warpgroup, and the
//     MMA warp alternates  PV_A(j), QK_A(j+1), PV_B(j), QK_B(j+1)  so tile A's softmax runs under tile B's MMAs
//     and vice versa;  K/V tiles are fetched once per CTA for both q tiles (half the L2 traffic per q row);
//   * one thread owns one q row (TMEM lane): row max / row sum need no cross-thread exchange and no block barrier;
//   * P never touches shared memory: it is written back to TMEM as packed bf16 over the S columns it came from
//     (tcgen05.st) and consumed as the TMEM A operand of the PV MMA (tcgen05.mma with A in tensor memory);
//   * the scale-and-subtract and the row sum use the packed fp32x2 pipe (fma.rn.f32x2 / add.f32x2).
//
//   warp 0        TMA producer: Q_A, Q_B once; K_j, V_j alternating through ONE 5-stage ring of 32 KB tiles
//   warp 1        MMA issuer
//   warps 2..5    softmax of tile A (rows q0 .. q0+127), one thread per row
//   warps 6..9    softmax of tile B (rows q0+128 .. q0+255)
//   TMEM (512 columns): S_A|P_A [0,128)  S_B|P_B [128,256)  O_A [256,384)  O_B [384,512)
//
// Rounding points are those of fa_fwd.cu (and of the reference's flash path): S and the softmax in
// fp32 with the scale applied to S, P rounded to bf16 before P@V, O accumulated in fp32 and rounded to bf16 once.
// Replaces F.scaled_dot_product_attention(is_causal=True) (paddlenlp/transformers/llama/fusion_ops.py:240-246).
#include <type_traits>

#include "../../include/b200nlp.h"
#include "common.cuh"
#include "host_util.h"

namespace b200 {
namespace fa2 {

constexpr int D = 128;
constexpr int TILE_BYTES = 128 * 128 * 2;   // 32 KB: one 128x128 bf16 tile = two 64-column halves of 16 KB
constexpr int HALF_BYTES = TILE_BYTES / 2;
constexpr int NST = 5;                      // K/V ring stages
constexpr int NUM_THREADS = 320;
constexpr int BAR_BYTES = 512;
constexpr int SMEM_BYTES = (2 + NST) * TILE_BYTES + BAR_BYTES + 1024;   // Q_A, Q_B, ring, barriers, align slack
constexpr float RESCALE_THRESHOLD = 8.f;    // log2 units: O is rescaled only when the row max grew by more than 2^8

struct Params {
  int S, B, nh, kvh;
  float scale_log2;   // (1/sqrt(d)) * log2(e)
  float* lse;         // [B, nh, S]
  // PAGED instantiation (prefill half of append_attention, csrc/gpu/append_attention.cu:428-851): sequence b contributes
  // seq_this[b] new query rows (token rows cu_q[b] .. of the packed projection) at absolute positions seq_dec[b] + i and attends to
  // cache positions [0, seq_dec[b] + i] of its pages; key/value caches [num_blocks, kvh, block_size, 128]
  const int* cu_q;
  const int* seq_dec;
  const int* seq_this;
  const int* seq_enc;
  const int* block_tables;
  int max_blocks, block_size;
  bf16* out;          // [token_num, ldo]
  int64_t ldo;
  int exp_poly;       // 0 / 1 / 2: none / a quarter / half of the exponentials on the FMA pipe (exp2_poly2)
  // MASK instantiation — FlashMask, causal lower-triangular form (fusion_ops.py:218-231 -> F.flashmask_attention(
  // startend_row_indices, causal=True)): mask_start[b, c] = first query row that may NOT see key column c (the end of c's packed
  // document), non-decreasing in c and > c.  Row i sees column c iff c <= i < mask_start[b, c].
  const int* mask_start;
};

// (x0, x1) = (a0, a1) * s + c    on the packed fp32x2 pipe
__device__ __forceinline__ void fma2(float& x0, float& x1, float a0, float a1, float s, float c) {
  asm("{\n\t.reg .b64 ra, rs, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\tmov.b64 rc, {%5, %5};\n\t"
      "fma.rn.f32x2 rd, ra, rs, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(x0), "=f"(x1)
      : "f"(a0), "f"(a1), "f"(s), "f"(c));
}
// (acc0, acc1) += (a0, a1)
__device__ __forceinline__ void add2(float& acc0, float& acc1, float a0, float a1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "+f"(acc0), "+f"(acc1)
      : "f"(a0), "f"(a1));
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// 2^x for a pair of non-positive arguments on the FMA pipe instead of the MUFU (16 ex2/clk/SM: the 32768 exponentials of one kv
// step of the two q tiles occupy it for 2048 cycles, exactly the tensor time of that step, so the softmax of a tile is
// MUFU-paced).  x = xr + xf with xr = round(x) taken from the low mantissa bits of x + 1.5*2^23, 2^xf by a degree-3 minimax
// polynomial on [-0.5, 0.5] (relative error 7.5e-5, far below the bf16 rounding P gets anyway), 2^xr by adding xr to the exponent
// field.  Arguments are clamped at -125 so that the exponent add cannot wrap (2^-125 instead of 0 for masked / far-away scores).
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& p0, float& p1) {
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  uint32_t t0, t1, r0, r1;
  asm("{\n\t.reg .b64 rx, rt, rr, rf, rp, rc;\n\t"
      "mov.b64 rx, {%4, %5};\n\t"
      "mov.b64 rc, {%6, %6};\n\t"
      "add.rn.f32x2 rt, rx, rc;\n\t"          // t = x + 1.5 * 2^23
      "mov.b64 {%2, %3}, rt;\n\t"
      "mov.b64 rc, {%7, %7};\n\t"
      "add.rn.f32x2 rr, rt, rc;\n\t"          // xr = t - 1.5 * 2^23 = round(x)
      "mov.b64 rc, {%8, %8};\n\t"
      "fma.rn.f32x2 rf, rr, rc, rx;\n\t"      // xf = x - xr
      "mov.b64 rp, {%9, %9};\n\t"
      "mov.b64 rc, {%10, %10};\n\t"
      "fma.rn.f32x2 rp, rp, rf, rc;\n\t"      // c3 xf + c2
      "mov.b64 rc, {%11, %11};\n\t"
      "fma.rn.f32x2 rp, rp, rf, rc;\n\t"      // ... xf + c1
      "mov.b64 rc, {%12, %12};\n\t"
      "fma.rn.f32x2 rp, rp, rf, rc;\n\t"      // ... xf + c0
      "mov.b64 {%0, %1}, rp;\n\t}"
      : "=r"(r0), "=r"(r1), "=r"(t0), "=r"(t1)
      : "f"(x0), "f"(x1), "f"(12582912.f), "f"(-12582912.f), "f"(-1.f), "f"(0.055171649903059006f), "f"(0.2426111251115799f),
        "f"(0.6932609677314758f), "f"(0.9999280571937561f));
  p0 = __uint_as_float(r0 + (t0 << 23));
  p1 = __uint_as_float(r1 + (t1 << 23));
}

template <bool PAGED, bool MASK>
// 10 warps = 3 on one SM sub-partition (16 K registers each): 16384 / (3 * 32) = 170 registers per thread is the hardware limit for this
// block shape (a 200-register build fails to launch), which is what __launch_bounds__(320, 1) makes ptxas target
__global__ void __launch_bounds__(NUM_THREADS, 1)
fa_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                          // [2] tiles: A, B (reused as the O staging of the same tile)
  uint8_t* sKV = smem + 2 * TILE_BYTES;        // [NST] ring: item 2j = K_j, item 2j+1 = V_j
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (2 + NST) * TILE_BYTES);
  uint64_t* q_full = bars;                     // [2]
  uint64_t* kv_full = bars + 2;                // [NST]
  uint64_t* kv_empty = bars + 2 + NST;         // [NST]
  uint64_t* s_full = bars + 2 + 2 * NST;       // [2]  MMA -> softmax: S_t(j) complete (and every earlier MMA)
  uint64_t* p_full = s_full + 2;               // [2][2] softmax -> MMA: half h (64 kv columns) of P_t(j) in TMEM, O_t rescaled
  uint64_t* o_full = p_full + 4;               // [2]  MMA -> softmax: last PV of tile t complete
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int kv_head = head / (p.nh / p.kvh);
  int q0, nA, nB, pos0 = 0, n_rows = p.S, tok0 = 0;
  bool b_active;
  if constexpr (PAGED) {
    n_rows = p.seq_this[batch];
    // decode rows (one new token on top of a cache, no prompt) belong to the decode kernel; idle slots to nobody
    if (n_rows <= 0 || (n_rows == 1 && p.seq_enc[batch] <= 0)) return;
    const int qb = (n_rows + 255) / 256 - 1 - static_cast<int>(blockIdx.x);
    if (qb < 0) return;
    pos0 = p.seq_dec[batch];
    tok0 = p.cu_q[batch];
    q0 = qb * 256;
    b_active = (q0 + 128) < n_rows;
    nA = (pos0 + min(q0 + 127, n_rows - 1)) / 128 + 1;          // kv tiles that tile A's last row can see
    nB = b_active ? (pos0 + min(q0 + 255, n_rows - 1)) / 128 + 1 : 0;
  } else {
    const int num_qb = (p.S + 255) / 256;
    const int qb = num_qb - 1 - static_cast<int>(blockIdx.x);   // heavy blocks first
    q0 = qb * 256;
    b_active = (q0 + 128) < p.S;
    nA = 2 * qb + 1;                   // kv tiles 0 .. 2qb     (diagonal = last)
    nB = b_active ? 2 * qb + 2 : 0;    // kv tiles 0 .. 2qb+1   (diagonal = last)
  }
  // With a document mask the leading kv tiles whose every column belongs to a document that ended at or before a q tile's first
  // row are skipped (mask_start is non-decreasing, so they form a prefix; the diagonal tile is never empty).  Tile B starts
  // 128 rows later than tile A and may skip more: the CTA's kv stream starts at tile A's first tile `lo`, tile B joins at
  // relative tile loB.  From here on kv tile indices are RELATIVE to lo (ring items 2j = K_{lo+j}, 2j+1 = V_{lo+j}).
  int lo = 0, loB = 0;
  if constexpr (MASK) {
    const int* ms = p.mask_start + static_cast<size_t>(batch) * p.S;
    while (lo < nA - 1 && __ldg(ms + min(lo * 128 + 127, p.S - 1)) <= q0) ++lo;
    if (b_active) {
      loB = lo;
      while (loB < nB - 1 && __ldg(ms + min(loB * 128 + 127, p.S - 1)) <= q0 + 128) ++loB;
      loB -= lo;
    }
    nA -= lo;
    if (b_active) nB -= lo;
  }
  const int n_kv = b_active ? nB : nA;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[2 * i], 128);
      mbar_init(&p_full[2 * i + 1], 128);
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < NST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, 512);
  if constexpr (PAGED) {
    // pages past the end of a sequence are not fetched: the ring must hold finite values there (P = 0 for those columns, but
    // 0 * NaN = NaN in the PV accumulation)
    if (warp >= 2) {
      for (int i = threadIdx.x - 64; i < NST * TILE_BYTES / 16; i += 256) reinterpret_cast<uint4*>(sKV)[i] = make_uint4(0, 0, 0, 0);
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      if constexpr (PAGED) {
        mbar_arrive_expect_tx(&q_full[0], TILE_BYTES);
        tma_load_3d(&tmQ, &q_full[0], sQ, 0, head, tok0 + q0);
        tma_load_3d(&tmQ, &q_full[0], sQ + HALF_BYTES, 64, head, tok0 + q0);
        if (b_active) {
          mbar_arrive_expect_tx(&q_full[1], TILE_BYTES);
          tma_load_3d(&tmQ, &q_full[1], sQ + TILE_BYTES, 0, head, tok0 + q0 + 128);
          tma_load_3d(&tmQ, &q_full[1], sQ + TILE_BYTES + HALF_BYTES, 64, head, tok0 + q0 + 128);
        }
        const int kv_total = pos0 + n_rows;                 // cache positions that exist for this sequence after the append
        const int ppt = 128 / p.block_size;
        const uint32_t page_bytes = static_cast<uint32_t>(p.block_size) * D * 2;
        for (int it = 0; it < 2 * n_kv; ++it) {
          const int st = it % NST;
          const uint32_t use = static_cast<uint32_t>(it / NST);
          const int t0 = (it >> 1) * 128;
          int phys[4];
          int npages = 0;
          for (int pg = 0; pg < ppt; ++pg) {
            const int tpos = t0 + pg * p.block_size;
            if (tpos < kv_total) phys[npages++] = __ldg(p.block_tables + static_cast<size_t>(batch) * p.max_blocks + tpos / p.block_size);
          }
          mbar_wait(&kv_empty[st], (use & 1u) ^ 1u);
          mbar_arrive_expect_tx(&kv_full[st], page_bytes * npages);
          uint8_t* dst = sKV + st * TILE_BYTES;
          const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
          for (int pg = 0; pg < npages; ++pg) {
            tma_load_4d(tm, &kv_full[st], dst + pg * (page_bytes / 2), 0, 0, kv_head, phys[pg]);
            tma_load_4d(tm, &kv_full[st], dst + pg * (page_bytes / 2) + HALF_BYTES, 64, 0, kv_head, phys[pg]);
          }
        }
      } else {
        mbar_arrive_expect_tx(&q_full[0], TILE_BYTES);
        tma_load_4d(&tmQ, &q_full[0], sQ, 0, head, q0, batch);
        tma_load_4d(&tmQ, &q_full[0], sQ + HALF_BYTES, 64, head, q0, batch);
        if (b_active) {
          mbar_arrive_expect_tx(&q_full[1], TILE_BYTES);
          tma_load_4d(&tmQ, &q_full[1], sQ + TILE_BYTES, 0, head, q0 + 128, batch);
          tma_load_4d(&tmQ, &q_full[1], sQ + TILE_BYTES + HALF_BYTES, 64, head, q0 + 128, batch);
        }
        for (int it = 0; it < 2 * n_kv; ++it) {
          const int st = it % NST;
          const uint32_t use = static_cast<uint32_t>(it / NST);
          mbar_wait(&kv_empty[st], (use & 1u) ^ 1u);
          mbar_arrive_expect_tx(&kv_full[st], TILE_BYTES);
          uint8_t* dst = sKV + st * TILE_BYTES;
          const CUtensorMap* tm = (it & 1) ? &tmV : &tmK;
          const int row = (lo + (it >> 1)) * 128;
          tma_load_4d(tm, &kv_full[st], dst, 0, kv_head, row, batch);
          tma_load_4d(tm, &kv_full[st], dst + HALF_BYTES, 64, kv_head, row, batch);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    // Convergent code: the whole warp runs the control flow, the barrier waits and the descriptor arithmetic, one elected lane
    // issues.  (Inside `if (lane == 0)` every descriptor is rebuilt in vector registers and moved to the uniform registers
    // UTCHMMA reads through ELECT / R2UR.BROADCAST loops: ~19 SASS instructions, ~120 cycles per MMA against 64 tensor cycles —
    // the first version of this kernel was bound by that, tensor pipe 50 %.)  Descriptors are built once per tile and advanced
    // by adding the k-step's byte offset >> 4.
    {
      const bool leader = elect_one();
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, false, false);   // A = Q (smem, K-major), B = K (K-major)
      constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, false, true);    // A = P (TMEM), B = V (MN-major)
      const uint32_t sQ_a = smem_u32(sQ), sKV_a = smem_u32(sKV);
      auto wait_kv = [&](int it) { mbar_wait(&kv_full[it % NST], static_cast<uint32_t>(it / NST) & 1u); };
      auto free_kv = [&](int it) { if (leader) umma_commit(&kv_empty[it % NST]); };
      auto koff = [](int kk) { return static_cast<uint64_t>(((kk >> 2) * HALF_BYTES + (kk & 3) * 32) >> 4); };
      auto issue_qk = [&](int t, int j) {          // S_t = Q_t K_j^T
        const uint64_t dQ = umma_desc_sw128(sQ_a + t * TILE_BYTES, 16, 1024);
        const uint64_t dK = umma_desc_sw128(sKV_a + ((2 * j) % NST) * TILE_BYTES, 16, 1024);
        const uint32_t tS = tbase + static_cast<uint32_t>(t * 128);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < D / 16; ++kk) umma_ss<1>(tS, dQ + koff(kk), dK + koff(kk), idesc_qk, kk > 0 ? 1u : 0u);
          umma_commit(&s_full[t]);
        }
      };
      // O_t (+)= P_t V_j ; P_t = packed bf16 in the first 64 columns of S_t.  The softmax warps publish P in two halves of 64 kv
      // columns: the first four k-steps run while the second half is still being exponentiated.
      auto issue_pv = [&](int t, int j, int jl) {   // jl = j - (first kv tile of q tile t): phase and accumulate flag
        const uint64_t dV = umma_desc_sw128(sKV_a + ((2 * j + 1) % NST) * TILE_BYTES, HALF_BYTES, 1024);
        const uint32_t tP = tbase + static_cast<uint32_t>(t * 128);
        const uint32_t tO = tbase + 256u + static_cast<uint32_t>(t * 128);
        const uint32_t acc0 = jl > 0 ? 1u : 0u;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mbar_wait(&p_full[2 * t + h], jl & 1);
          if (h == 0) wait_kv(2 * j + 1);
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int kk = 4 * h; kk < 4 * h + 4; ++kk)
              umma_ts(tO, tP + kk * 8, dV + static_cast<uint64_t>(kk * 128), idesc_pv, kk > 0 ? 1u : acc0);
          }
        }
      };
      mbar_wait(&q_full[0], 0);
      wait_kv(0);
      tc_fence_after();
      issue_qk(0, 0);
      if (b_active && loB == 0) {
        mbar_wait(&q_full[1], 0);
        tc_fence_after();
        issue_qk(1, 0);
      }
      free_kv(0);
      for (int j = 0; j < n_kv; ++j) {
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int nt = t == 0 ? nA : nB;
          const int lt = t == 0 ? 0 : loB;
          if (j >= lt && j < nt) {
            issue_pv(t, j, j - lt);
            if (j + 1 < nt) {
              wait_kv(2 * j + 2);
              tc_fence_after();
              issue_qk(t, j + 1);
            } else if (leader) {
              umma_commit(&o_full[t]);
            }
          } else if (MASK && t == 1 && b_active && j + 1 == lt) {   // tile B joins the stream at its first visible kv tile
            mbar_wait(&q_full[1], 0);
            wait_kv(2 * j + 2);
            tc_fence_after();
            issue_qk(1, j + 1);
          }
        }
        // both tiles have issued everything that reads V_j and K_{j+1}: hand the stages back when those MMAs retire
        free_kv(2 * j + 1);
        if (j + 1 < n_kv) free_kv(2 * j + 2);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------- softmax / epilogue: one thread per q row -------------------------------
    const int t = (warp - 2) >> 2;                        // 0: tile A, 1: tile B
    const int nt = t == 0 ? nA : nB;
    const int lt = t == 0 ? 0 : loB;
    if (nt > 0) {
      const int quad = warp & 3;                          // TMEM lane quadrant this warp may touch
      const int r = quad * 32 + lane;                     // row within the tile == TMEM lane
      const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
      const uint32_t tS = tmem_base + lane_off + static_cast<uint32_t>(t * 128);
      const uint32_t tO = tmem_base + lane_off + 256u + static_cast<uint32_t>(t * 128);
      const int q0t = q0 + t * 128;
      float m_used = -INFINITY, l0 = 0.f, l1 = 0.f;
      for (int j = lt; j < nt; ++j) {
        mbar_wait(&s_full[t], (j - lt) & 1);
        tc_fence_after();
        uint32_t sv[128];
        {
          uint32_t(*c)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
          tmem_ld32(tS, c[0]); tmem_ld32(tS + 32, c[1]); tmem_ld32(tS + 64, c[2]); tmem_ld32(tS + 96, c[3]);
          tmem_ld_wait();
        }
        if constexpr (PAGED) {
          // row r sits at absolute position pos0 + q0t + r and sees cache positions <= that: column c of kv tile j is position
          // 128 j + c.  With a cached prefix the diagonal band is not tile aligned: up to two kv tiles per q tile need the mask.
          if (128 * j + 127 > pos0 + q0t) {
            const int lim = pos0 + q0t + r - 128 * j;
#pragma unroll
            for (int c = 0; c < 128; ++c)
              if (c > lim) sv[c] = 0xff800000u;           // -inf
          }
        } else if (j == nt - 1) {                         // diagonal tile: columns beyond the row are masked
#pragma unroll
          for (int c = 0; c < 128; ++c)
            if (c > r) sv[c] = 0xff800000u;               // -inf
        }
        if constexpr (MASK) {
          // columns whose document ended at or before this row: start rows are non-decreasing, so they are the first `hid`
          // columns of the tile, hid = #{c : mask_start[c] <= row} by bisection (7 L1-resident loads, only in tiles where some
          // document ends at or before the q tile's last row)
          const int col0 = (lo + j) * 128;
          const int* ms = p.mask_start + static_cast<size_t>(batch) * p.S + col0;
          if (__ldg(ms) <= q0t + 127) {
            const int row = q0t + r;
            int a = 0, b = min(128, p.S - col0);
            while (a < b) {
              const int mid = (a + b) >> 1;
              if (__ldg(ms + mid) <= row) a = mid + 1; else b = mid;
            }
#pragma unroll
            for (int c = 0; c < 128; ++c)
              if (c < a) sv[c] = 0xff800000u;             // -inf
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 128; c += 8) {
          mx0 = max3(mx0, __uint_as_float(sv[c]), __uint_as_float(sv[c + 1]));
          mx1 = max3(mx1, __uint_as_float(sv[c + 2]), __uint_as_float(sv[c + 3]));
          mx2 = max3(mx2, __uint_as_float(sv[c + 4]), __uint_as_float(sv[c + 5]));
          mx3 = max3(mx3, __uint_as_float(sv[c + 6]), __uint_as_float(sv[c + 7]));
        }
        const float rowmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * p.scale_log2;   // scale > 0
        bool need = false;
        float factor = 1.f;
        if (j == lt) {
          m_used = rowmax;
        } else if (rowmax > m_used + RESCALE_THRESHOLD) {
          need = true;
          factor = fast_exp2(m_used - rowmax);
          l0 *= factor; l1 *= factor;
          m_used = rowmax;
        }
        const bool rescale = __any_sync(0xffffffffu, need);
        // with documents a row can be fully masked in its first tiles: (-inf) * scale - (-inf) must not be evaluated
        const float neg_m = (MASK && m_used == -INFINITY) ? 0.f : -m_used;
        // P over the first 64 columns of this row's S (every S value of the row is already in registers), published in two halves
        // of 64 kv columns so that the first four PV k-steps run under the second half's exponentials
        auto half = [&](auto mode_tag, auto half_tag, bool publish) {
          constexpr int MODE = decltype(mode_tag)::value;   // share of the exponentials evaluated on the FMA pipe: 0, 1/4, 1/2
          constexpr int H = decltype(half_tag)::value;
          uint32_t pk[32];
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float x0, x1, p0, p1;
            fma2(x0, x1, __uint_as_float(sv[64 * H + 2 * c]), __uint_as_float(sv[64 * H + 2 * c + 1]), p.scale_log2, neg_m);
            if ((MODE == 1 && (c & 3) == 3) || (MODE == 2 && (c & 1) == 1)) {
              exp2_poly2(x0, x1, p0, p1);
            } else {
              p0 = fast_exp2(x0); p1 = fast_exp2(x1);
            }
            add2(l0, l1, p0, p1);
            pk[c] = pack_bf16x2(p0, p1);
          }
          tmem_st32(tS + 32 * H, pk);
          if (publish) {
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[2 * t + H]);
          }
        };
        auto both = [&](auto mode_tag) {
          if (!rescale) {
            half(mode_tag, std::integral_constant<int, 0>{}, true);
            half(mode_tag, std::integral_constant<int, 1>{}, true);
          } else {
            // (rare: the row max of some row grew by more than 2^8)  O must be rescaled before ANY PV k-step of this kv tile:
            // both halves of P first (the scores leave the registers), then O, then both halves are published together.
            // s_full(j) tracks every MMA issued before QK_t(j), PV_t(j-1) included: the O accumulator of this tile is quiescent
            half(mode_tag, std::integral_constant<int, 0>{}, false);
            half(mode_tag, std::integral_constant<int, 1>{}, false);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              uint32_t o[32];
              tmem_ld32(tO + ch * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * factor);
              tmem_st32(tO + ch * 32, o);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[2 * t]);
            mbar_arrive(&p_full[2 * t + 1]);
          }
        };
        if (p.exp_poly == 2) both(std::integral_constant<int, 2>{});
        else if (p.exp_poly == 1) both(std::integral_constant<int, 1>{});
        else both(std::integral_constant<int, 0>{});
      }
      // epilogue: O / l -> bf16 -> swizzled smem (this tile's Q buffer) -> TMA store ; LSE
      const float l = l0 + l1;
      const float inv_l = 1.f / l;
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      if constexpr (PAGED) {
        // packed token rows: a tile may end inside the batch's next sequence, so rows are stored one by one (256 bytes each)
        const bool valid = q0t + r < n_rows;
        bf16* orow = p.out + static_cast<size_t>(tok0 + q0t + r) * p.ldo + head * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t o[32];
          tmem_ld32(tO + ch * 32, o);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              uint4 v;
              v.x = pack_bf16x2(__uint_as_float(o[c8 * 8 + 0]) * inv_l, __uint_as_float(o[c8 * 8 + 1]) * inv_l);
              v.y = pack_bf16x2(__uint_as_float(o[c8 * 8 + 2]) * inv_l, __uint_as_float(o[c8 * 8 + 3]) * inv_l);
              v.z = pack_bf16x2(__uint_as_float(o[c8 * 8 + 4]) * inv_l, __uint_as_float(o[c8 * 8 + 5]) * inv_l);
              v.w = pack_bf16x2(__uint_as_float(o[c8 * 8 + 6]) * inv_l, __uint_as_float(o[c8 * 8 + 7]) * inv_l);
              *reinterpret_cast<uint4*>(orow + ch * 32 + c8 * 8) = v;
            }
          }
        }
      } else {
        const uint32_t sO_a = smem_u32(sQ + t * TILE_BYTES);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t o[32];
          tmem_ld32(tO + ch * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int c8 = 0; c8 < 4; ++c8) {
            const int k = (ch & 1) * 4 + c8;     // 16-byte chunk within the 128-byte half row; half = ch >> 1
            uint4 v;
            v.x = pack_bf16x2(__uint_as_float(o[c8 * 8 + 0]) * inv_l, __uint_as_float(o[c8 * 8 + 1]) * inv_l);
            v.y = pack_bf16x2(__uint_as_float(o[c8 * 8 + 2]) * inv_l, __uint_as_float(o[c8 * 8 + 3]) * inv_l);
            v.z = pack_bf16x2(__uint_as_float(o[c8 * 8 + 4]) * inv_l, __uint_as_float(o[c8 * 8 + 5]) * inv_l);
            v.w = pack_bf16x2(__uint_as_float(o[c8 * 8 + 6]) * inv_l, __uint_as_float(o[c8 * 8 + 7]) * inv_l);
            st_shared_v4(sO_a + (ch >> 1) * HALF_BYTES + r * 128 + ((k ^ (r & 7)) << 4), v);
          }
        }
        if (q0t + r < p.S)
          p.lse[(static_cast<size_t>(batch) * p.nh + head) * p.S + q0t + r] = (m_used + log2f(l)) * 0.6931471805599453f;
        fence_proxy_async_smem();
        named_bar_sync(1 + t, 128);
        if ((warp == 2 || warp == 6) && lane == 0) {
          tma_store_4d(&tmO, sQ + t * TILE_BYTES, 0, head, q0t, batch);
          tma_store_4d(&tmO, sQ + t * TILE_BYTES + HALF_BYTES, 64, head, q0t, batch);
          tma_store_commit();
          tma_store_wait<0>();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// 4-D map over a [B, S, heads, 128] bf16 view with token stride `ld` (elements): dims {128, heads, S, B}, 64x128 boxes.
static int make_map(CUtensorMap* tm, const void* base, int64_t B, int64_t S, int64_t heads, int64_t ld) {
  uint64_t dims[4] = {128, static_cast<uint64_t>(heads), static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {128 * 2, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(S) * ld * 2};
  uint32_t box[4] = {64, 1, 128, 1};
  return encode_tmap_bf16(tm, base, 4, dims, strides, box);
}

}  // namespace fa2

// Causal forward (optionally with FlashMask start rows) through the two-tile kernel; called by b200_fa_fwd_flashmask (fa_fwd.cu).
int launch_fa_fwd2(const void* q, const void* k, const void* v, void* o, float* lse, const int32_t* mask_start_rows, int64_t B, int64_t S, int64_t num_heads,
                   int64_t num_kv_heads, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, float softmax_scale,
                   cudaStream_t stream) {
  using namespace fa2;
  CUtensorMap tmQ, tmK, tmV, tmO;
  int rc;
  if ((rc = make_map(&tmQ, q, B, S, num_heads, ldq)) != 0) return rc;
  if ((rc = make_map(&tmK, k, B, S, num_kv_heads, ldk)) != 0) return rc;
  if ((rc = make_map(&tmV, v, B, S, num_kv_heads, ldv)) != 0) return rc;
  if ((rc = make_map(&tmO, o, B, S, num_heads, ldo)) != 0) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fa_fwd2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("fa_fwd2 smem attr: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  Params p;
  p.S = static_cast<int>(S); p.B = static_cast<int>(B); p.nh = static_cast<int>(num_heads);
  p.kvh = static_cast<int>(num_kv_heads);
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.lse = lse;
  dim3 grid(static_cast<unsigned>((S + 255) / 256), static_cast<unsigned>(num_heads), static_cast<unsigned>(B));
  p.cu_q = p.seq_dec = p.seq_this = p.seq_enc = p.block_tables = nullptr;
  p.max_blocks = p.block_size = 0; p.out = nullptr; p.ldo = 0;
  p.exp_poly = fa_exp_poly();
  p.mask_start = mask_start_rows;
  if (mask_start_rows != nullptr) fa_fwd2_kernel<false, true><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  else fa_fwd2_kernel<false, false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmO, p);
  return check_launch("fa_fwd2");
}

// Prefill half of append_attention: causal attention of the NEW token rows of every prompt / prompt-chunk sequence over its paged
// cache (cached prefix + the rows themselves, already appended).  qkv: packed projection [token_num, ldq] (q heads first, rotated);
// key / value caches [num_blocks, kvh, block_size, 128]; out [token_num, ldo].  max_q_len bounds seq_lens_this_time (grid size).
int launch_fa_prefill_paged(const void* qkv, const void* key_cache, const void* value_cache, void* out, const int32_t* cu_seqlens_q,
                            const int32_t* seq_lens_encoder, const int32_t* seq_lens_decoder, const int32_t* seq_lens_this_time,
                            const int32_t* block_tables, int64_t B, int64_t token_num, int64_t max_q_len, int64_t num_heads,
                            int64_t num_kv_heads, int64_t num_blocks, int64_t block_size, int64_t max_blocks_per_seq, int64_t ldq,
                            int64_t ldo, float softmax_scale, cudaStream_t stream) {
  using namespace fa2;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  {
    uint64_t dims[3] = {128, static_cast<uint64_t>(num_heads), static_cast<uint64_t>(token_num)};
    uint64_t strides[2] = {128 * 2, static_cast<uint64_t>(ldq) * 2};
    uint32_t box[3] = {64, 1, 128};
    if ((rc = encode_tmap_bf16(&tmQ, qkv, 3, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[4] = {128, static_cast<uint64_t>(block_size), static_cast<uint64_t>(num_kv_heads), static_cast<uint64_t>(num_blocks)};
    uint64_t strides[3] = {128 * 2, static_cast<uint64_t>(block_size) * 128 * 2,
                           static_cast<uint64_t>(num_kv_heads) * block_size * 128 * 2};
    uint32_t box[4] = {64, static_cast<uint32_t>(block_size), 1, 1};
    if ((rc = encode_tmap_bf16(&tmK, key_cache, 4, dims, strides, box)) != 0) return rc;
    if ((rc = encode_tmap_bf16(&tmV, value_cache, 4, dims, strides, box)) != 0) return rc;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("fa_prefill_paged smem attr: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  Params p;
  p.S = static_cast<int>(max_q_len); p.B = static_cast<int>(B); p.nh = static_cast<int>(num_heads);
  p.kvh = static_cast<int>(num_kv_heads);
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.lse = nullptr;
  p.cu_q = cu_seqlens_q; p.seq_dec = seq_lens_decoder; p.seq_this = seq_lens_this_time; p.seq_enc = seq_lens_encoder;
  p.block_tables = block_tables;
  p.max_blocks = static_cast<int>(max_blocks_per_seq); p.block_size = static_cast<int>(block_size);
  p.out = static_cast<bf16*>(out); p.ldo = ldo;
  p.exp_poly = fa_exp_poly();
  p.mask_start = nullptr;
  dim3 grid(static_cast<unsigned>((max_q_len + 255) / 256), static_cast<unsigned>(num_heads), static_cast<unsigned>(B));
  fa_fwd2_kernel<true, false><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmQ, p);
  return check_launch("fa_prefill_paged");
}

}  // namespace b200
