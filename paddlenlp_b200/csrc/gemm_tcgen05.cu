// bf16 GEMM on 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA).
//
//   C[M,N] (+)= op(A)[M,K] * op(B)[K,N] (+ bias[N])      bf16 in, fp32 accumulate, one rounding to bf16.
//
// Replaces the cuBLAS(Lt) calls under paddle `nn.Linear` on the Llama/Qwen2 hot path
// (reference: paddlenlp/transformers/llama/modeling.py:771-799 q/k/v/o, :627-630 gate/up/down, :1894-1921 lm_head;
//  weights are stored [in,out] = "B is [K,N] row-major" = MN-major B operand) and the dX / dW GEMMs of their backward.
//
// Design (one persistent kernel, warp-specialised, 192 threads):
//   warp 0      TMA producer   : global -> 128B-swizzled smem ring, STAGES deep, BK = 64
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (UMMA M=128*CG, N=256, K=16); tcgen05.commit
//                                releases smem stages and publishes the finished accumulator
//   warps 2..5  epilogue       : tcgen05.ld (TMEM -> regs), (+bias, +C_old | +residual), round to bf16, swizzled smem staging,
//                                per-warp TMA store of 32x64 boxes
//   TMEM        2 accumulator stages x 256 fp32 columns = all 512 columns, so the epilogue of tile i overlaps the
//               main loop of tile i+1.
//   CG = 2      CTA pair (cluster of 2, cta_group::2): each CTA loads its 128 rows of A and its 128 columns of B;
//               the pair computes a 256x256 tile, halving per-SM shared-memory operand traffic.
//   Operand majors: both K-major (contraction dim contiguous) and MN-major operands are fed straight from their
//   row-major global layout through TMA; no transposes are materialised.
#include "../../include/b200nlp.h"
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "host_util.h"

namespace b200 {
namespace gemm {

constexpr int BM = 128;   // rows of the tile owned by one CTA
constexpr int BN = 256;   // tile columns (UMMA N)
constexpr int BK = 64;    // K per pipeline stage (= one 128-byte swizzle row of bf16)
constexpr int UK = 16;    // UMMA K for 16-bit inputs
constexpr int NUM_THREADS = 192;
constexpr int EPI_WARPS = 4;
constexpr int EPI_BOX_ROWS = 32;
constexpr int EPI_BOX_COLS = 64;
constexpr int EPI_BUF_BYTES = EPI_BOX_ROWS * EPI_BOX_COLS * 2;  // 4 KB
constexpr int TMEM_COLS = 512;

template <int CG>
struct Cfg {
  static constexpr int B_COLS = BN / CG;                       // B columns staged per CTA
  static constexpr int A_BYTES = BM * BK * 2;                  // 16 KB
  static constexpr int B_BYTES = B_COLS * BK * 2;              // 32 KB (CG=1) / 16 KB (CG=2)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (CG == 2) ? 6 : 4;
  static constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;  // 32 KB
  static constexpr int BAR_BYTES = 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
};

struct Params {
  int M, N, K;
  int num_m_tiles, num_n_tiles;
  int epi_mode;            // 0: C = acc ; 1: C = bf16(C_old + acc) ; 2: C = bf16(bf16(acc) + R) ; 3: F32ws += acc (split-K)
                           // 4: gate|up GEMM + SwiGLU: a 256-column tile = 128 gate columns | the 128 up columns of the same
                           //    channels; C gets gate and up (bf16, their [M, 2I] positions), tmR's tensor gets
                           //    m = bf16(silu(gate) * up) [M, I]
                           // 5: down-proj dX GEMM + SwiGLU backward: acc = d(m) tile; tmR's tensor = saved gate|up [M, 2I];
                           //    C = [d(gate) | d(up)] [M, 2I]
  int swiglu_inter;        // modes 4, 5: I (the up half starts at column I)
  int gm;                  // m-tiles per raster group (tile_coords)
  int split_k;             // work items per output tile (K is cut into split_k ranges of kb_per_split k-blocks)
  int kb_per_split;
  int b_prefetch;          // PDL: issue the first stages' B (weight) loads before griddepcontrol.wait
  int l2_prefetch_kb;      // PDL: ... and L2-prefetch this many further B k-blocks of the CTA's first work item
  const float* bias;       // [N] fp32 or nullptr
  const void* aux = nullptr;   // mode 5: the saved gate|up projection [M, 2I] bf16 (read with plain loads, one row per thread)
  int64_t ld_aux = 0;      //         its leading dimension in elements
};

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& m_blk, int& n_blk, int GM) {
  // Grouped ordering: GM consecutive m-tiles share each n-tile column so that concurrently running CTAs reuse
  // A and B tiles out of L2.  GM is chosen per shape by the host (Params::gm).
  const int per_group = GM * num_n;
  const int g = t / per_group;
  const int first_m = g * GM;
  const int gsize = min(GM, num_m - first_m);
  const int r = t - g * per_group;
  m_blk = first_m + (r % gsize);
  n_blk = r / gsize;
}

template <int CG, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const Params p) {
  using C = Cfg<CG>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + C::EPI_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;          // [STAGES]
  uint64_t* tmem_full = bars + 2 * C::STAGES;      // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint64_t* epi_bar = tmem_empty + 2;              // [EPI_WARPS]
  uint64_t* epi_ld_bar = epi_bar + EPI_WARPS;      // [EPI_WARPS][4]  mode 5: gate|up slab loads of each epilogue warp
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(epi_ld_bar + 4 * EPI_WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int pair_id = blockIdx.x / CG;
  const int num_pairs = gridDim.x / CG;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;   // work items
  const int num_kb_total = (p.K + BK - 1) / BK;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    tma_prefetch_desc(&tmR);
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], CG * EPI_WARPS);
    }
    for (int i = 0; i < EPI_WARPS; ++i) mbar_init(&epi_bar[i], 1);
    for (int i = 0; i < 4 * EPI_WARPS; ++i) mbar_init(&epi_ld_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<CG>(tmem_ptr_smem, TMEM_COLS);
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // first column of this CTA's c-th 64-column chunk of the B tile (MN-major B).  Mode 4: chunk g of the 256-column tile is
      // g = 0,1: gate channels [128 n_blk, +128), g = 2,3: the up columns of the same channels (at column I + ...), so the stored
      // [h, 2I] weight needs no interleaved copy
      auto b_col64 = [&](int n_blk, int c) {
        if (p.epi_mode == 4) {
          const int g = static_cast<int>(cta_rank) * (C::B_COLS / 64) + c;
          return n_blk * 128 + (g & 1) * 64 + (g >> 1) * p.swiglu_inter;
        }
        return n_blk * BN + static_cast<int>(cta_rank) * C::B_COLS + c * 64;
      };
      // With PDL (p.b_prefetch): the B operand (weights) of the first stages does not depend on the previous kernel, so
      // its TMA loads are issued before griddepcontrol.wait; the A loads (activations) follow after the wait.
      int prefetched = 0;
      if (p.b_prefetch && pair_id < num_tiles) {
        int m_blk, n_blk;
        tile_coords(pair_id / p.split_k, p.num_m_tiles, p.num_n_tiles, m_blk, n_blk, p.gm);
        const int kb0 = (pair_id % p.split_k) * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        const int n0 = n_blk * BN + static_cast<int>(cta_rank) * C::B_COLS;
        prefetched = min(C::STAGES, kb1 - kb0);
        for (int s = 0; s < prefetched; ++s) {
          uint8_t* sb = smem + s * C::STAGE_BYTES + C::A_BYTES;
          const int k0 = (kb0 + s) * BK;
          if (is_leader) mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES * CG);
          auto load = [&](const CUtensorMap* tm, void* dst, int c0, int c1) {
            if constexpr (CG == 2) tma_load_2d_pair(tm, &full_bar[s], dst, c0, c1);
            else tma_load_2d(tm, &full_bar[s], dst, c0, c1);
          };
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < C::B_COLS / 64; ++c) load(&tmB, sb + c * (64 * BK * 2), b_col64(n_blk, c), k0);
          } else {
#pragma unroll
            for (int c = 0; c < C::B_COLS / 128; ++c) load(&tmB, sb + c * (128 * BK * 2), k0, n0 + c * 128);
          }
        }
        // Beyond the smem ring: warm L2 with the next B k-blocks while the predecessor kernel is still running, so the
        // weight stream does not idle during the (activation-only) kernels between two GEMMs of the decode step.
        const int l2_end = min(kb1 - kb0, prefetched + p.l2_prefetch_kb);
        for (int s = prefetched; s < l2_end; ++s) {
          const int k0 = (kb0 + s) * BK;
          if constexpr (B_MN) {
#pragma unroll
            for (int c = 0; c < C::B_COLS / 64; ++c) tma_prefetch_l2_2d(&tmB, b_col64(n_blk, c), k0);
          } else {
#pragma unroll
            for (int c = 0; c < C::B_COLS / 128; ++c) tma_prefetch_l2_2d(&tmB, k0, n0 + c * 128);
          }
        }
      }
      pdl_wait();
      for (int t = pair_id; t < num_tiles; t += num_pairs) {
        int m_blk, n_blk;
        tile_coords(t / p.split_k, p.num_m_tiles, p.num_n_tiles, m_blk, n_blk, p.gm);
        const int kb0 = (t % p.split_k) * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        const int m0 = m_blk * (BM * CG) + static_cast<int>(cta_rank) * BM;   // this CTA's A rows
        const int n0 = n_blk * BN + static_cast<int>(cta_rank) * C::B_COLS;   // this CTA's B columns
        for (int kb = kb0; kb < kb1; ++kb) {
          const bool b_done = (t == pair_id) && (kb - kb0 < prefetched);     // B tile (and expect_tx) already issued
          if (!b_done) mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          const int k0 = kb * BK;
          if (is_leader && !b_done) mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES * CG);
          auto load = [&](const CUtensorMap* tm, void* dst, int c0, int c1) {
            if constexpr (CG == 2) tma_load_2d_pair(tm, &full_bar[stage], dst, c0, c1);
            else tma_load_2d(tm, &full_bar[stage], dst, c0, c1);
          };
          if constexpr (A_MN) {
            // A stored [K, M] row-major: boxes of 64 (M) x 64 (K), one per 64 rows of the tile.
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) load(&tmA, sa + c * (64 * BK * 2), m0 + c * 64, k0);
          } else {
            load(&tmA, sa, k0, m0);  // A stored [M, K] row-major: one 64 (K) x 128 (M) box
          }
          if (!b_done) {
            if constexpr (B_MN) {
#pragma unroll
              for (int c = 0; c < C::B_COLS / 64; ++c) load(&tmB, sb + c * (64 * BK * 2), b_col64(n_blk, c), k0);
            } else {
#pragma unroll
              for (int c = 0; c < C::B_COLS / 128; ++c) load(&tmB, sb + c * (128 * BK * 2), k0, n0 + c * 128);
            }
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    // The leader CTA's whole warp runs the (warp-uniform) loop, waits and descriptor arithmetic; one elected lane issues.  Inside
    // `if (lane == 0)` the compiler rebuilds every descriptor in vector registers and moves it to the uniform registers UTCHMMA
    // reads through ELECT / R2UR.BROADCAST sequences (~19 SASS instructions per MMA, measured on the attention kernels); here a
    // stage's descriptors are built once and advanced by adding the k-step's byte offset >> 4.
    if (is_leader) {
      const bool leader = elect_one();
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);
      constexpr uint32_t idesc = umma_idesc_bf16(BM * CG, BN, A_MN, B_MN);
      // K-major  : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused.       K advance = 32 B
      // MN-major : 64-element chunks along M/N 8 KB apart (LBO), 8-k groups 1024 B.  K advance = 16 rows = 2 KB
      constexpr uint32_t a_lbo = A_MN ? 64 * BK * 2 : 16, a_adv = A_MN ? UK * 128 : UK * 2;
      constexpr uint32_t b_lbo = B_MN ? 64 * BK * 2 : 16, b_adv = B_MN ? UK * 128 : UK * 2;
      const uint32_t smem_a0 = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = pair_id; t < num_tiles; t += num_pairs) {
        mbar_wait(&tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tbase + static_cast<uint32_t>(as * BN);
        const int kb0 = (t % p.split_k) * p.kb_per_split, kb1 = min(num_kb_total, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_a0 + stage * C::STAGE_BYTES;
          const uint64_t da0 = umma_desc_sw128(sa, a_lbo, 1024);
          const uint64_t db0 = umma_desc_sw128(sa + C::A_BYTES, b_lbo, 1024);
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_ss<CG>(tmem_d, da0 + static_cast<uint64_t>((k * a_adv) >> 4), db0 + static_cast<uint64_t>((k * b_adv) >> 4), idesc,
                          (kb > kb0 || k > 0) ? 1u : 0u);
            if constexpr (CG == 2) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        if (leader) {
          if constexpr (CG == 2) umma_commit_pair(&tmem_full[as]); else umma_commit(&tmem_full[as]);
        }
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
      __syncwarp();
    }
  } else {
    // ===================================== epilogue =====================================
    pdl_wait();                                    // outputs / residual / workspace belong to the previous kernel until now
    const int q = warp & 3;                        // TMEM lane quadrant this warp may access
    uint8_t* my_buf = epi_smem + q * 2 * EPI_BUF_BYTES;
    const uint32_t my_buf_s = smem_u32(my_buf);
    uint32_t ephase = 0;
    int as = 0;
    uint32_t aphase = 0;
    int buf = 0;
    if (p.epi_mode == 5) {
      // ---- down-proj dX GEMM + SwiGLU backward (llama/modeling.py:632-652) ----
      // acc = d(m) tile;  d(gate) = d(m) * up * silu'(gate),  d(up) = d(m) * silu(gate)   (swiglu_bwd_pair, common.cuh)
      // One thread = one token row; a slab = 16 channels.  The saved gate / up values of a slab come in by TMA (boxes of 32 rows
      // x 32 bytes, no swizzle) into one of FOUR 2 KB buffer pairs of this warp, two slabs ahead of their use and across tile
      // boundaries; the results overwrite them in place and leave by TMA store.  Nothing in this loop waits for a memory round
      // trip: the load of slab g+2 is issued when the store of slab g-2 has released its pair (one slab old).  (Earlier versions —
      // TMA load and use in the same slab, then per-thread 128-byte global loads one slab ahead — left the epilogue slower than the
      // 256x256x4096 mainloop it has to hide under: profiles/r02_swiglu_bwd_epilogue.md.)
      constexpr int SL = 16, NSL = BN / SL, SLAB_BYTES = EPI_BOX_ROWS * SL * 2;     // 1 KB per operand and slab
      uint64_t* ld_bar = epi_ld_bar + q * 4;
      int it_t = pair_id, it_s = 0;                  // next slab to request
      uint32_t n_issued = 0, n_done = 0;
      auto issue_next = [&]() {
        int m2 = 0, n2 = 0, c0 = 0;
        while (it_t < num_tiles) {
          tile_coords(it_t / p.split_k, p.num_m_tiles, p.num_n_tiles, m2, n2, p.gm);
          c0 = n2 * BN + it_s * SL;
          if (c0 < p.N) break;
          it_s = 0;                                  // N % 64 == 0: the dead slabs of a tile are its tail
          it_t += num_pairs;
        }
        if (it_t >= num_tiles) return;
        const uint32_t pr = n_issued & 3u;
        if (lane == 0) {
          const int r0 = m2 * (BM * CG) + static_cast<int>(cta_rank) * BM + q * 32;   // fully out-of-range boxes read zeros
          mbar_arrive_expect_tx(&ld_bar[pr], 2 * SLAB_BYTES);
          tma_load_2d(&tmR, &ld_bar[pr], my_buf + pr * 2 * SLAB_BYTES, c0, r0);
          tma_load_2d(&tmR, &ld_bar[pr], my_buf + pr * 2 * SLAB_BYTES + SLAB_BYTES, c0 + p.swiglu_inter, r0);
        }
        ++n_issued;
        if (++it_s == NSL) { it_s = 0; it_t += num_pairs; }
      };
      issue_next();
      issue_next();
      for (int t = pair_id; t < num_tiles; t += num_pairs) {
        int m_blk, n_blk;
        tile_coords(t / p.split_k, p.num_m_tiles, p.num_n_tiles, m_blk, n_blk, p.gm);
        const int row0 = m_blk * (BM * CG) + static_cast<int>(cta_rank) * BM + q * 32;
        const int col0 = n_blk * BN;
        const int n_live = min(NSL, (p.N - col0) / SL);
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
#pragma unroll 1
        for (int sl = 0; sl < n_live; ++sl) {
          const int c0 = col0 + sl * SL;
          if (lane == 0) tma_store_wait_read<1>();   // every store but the previous slab's has read its pair: slab n_done-2's is free
          __syncwarp();
          issue_next();
          uint32_t v[16];
          tmem_ld16(taddr + sl * SL, v);
          tmem_ld_wait();
          if (sl == n_live - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
          }
          const uint32_t pr = n_done & 3u;
          mbar_wait(&ld_bar[pr], (n_done >> 2) & 1u);
          const uint32_t ga = my_buf_s + pr * 2 * SLAB_BYTES + lane * (SL * 2), ua = ga + SLAB_BYTES;
          uint4 gv[2], uv[2], og[2], ou[2];
          gv[0] = ld_shared_v4(ga); gv[1] = ld_shared_v4(ga + 16);
          uv[0] = ld_shared_v4(ua); uv[1] = ld_shared_v4(ua + 16);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            // d(m) with the GEMM's own bf16 output rounding (one packed convert + unpack per pair)
            const float2 dr = unpack_bf16x2(pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])));
            swiglu_bwd_pair(reinterpret_cast<const uint32_t*>(&gv[j >> 2])[j & 3], reinterpret_cast<const uint32_t*>(&uv[j >> 2])[j & 3],
                            dr.x, dr.y, reinterpret_cast<uint32_t*>(&og[j >> 2])[j & 3], reinterpret_cast<uint32_t*>(&ou[j >> 2])[j & 3]);
          }
          st_shared_v4(ga, og[0]); st_shared_v4(ga + 16, og[1]);
          st_shared_v4(ua, ou[0]); st_shared_v4(ua + 16, ou[1]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, my_buf + pr * 2 * SLAB_BYTES, c0, row0);
            tma_store_2d(&tmC, my_buf + pr * 2 * SLAB_BYTES + SLAB_BYTES, c0 + p.swiglu_inter, row0);
            tma_store_commit();
          }
          ++n_done;
        }
        if (++as == 2) { as = 0; aphase ^= 1u; }
      }
    } else
    for (int t = pair_id; t < num_tiles; t += num_pairs) {
      int m_blk, n_blk;
      tile_coords(t / p.split_k, p.num_m_tiles, p.num_n_tiles, m_blk, n_blk, p.gm);
      const int row0 = m_blk * (BM * CG) + static_cast<int>(cta_rank) * BM + q * 32;
      const int col0 = n_blk * BN;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * BN);
      if (p.epi_mode == 4) {
        // gate|up + SwiGLU (llama/modeling.py:38-45, 632-652): accumulator columns [0,128) = gate, [128,256) = up of channels
        // [128 n_blk, +128).  Rounding points of the unfused path: gate and up each rounded to bf16 (the Linear outputs, kept for
        // the backward), then silu(g) * u in fp32 and one rounding.
#pragma unroll 1
        for (int sp = 0; sp < 2; ++sp) {
          uint32_t g0[32], g1[32], u0[32], u1[32];
          tmem_ld32(taddr + sp * 64, g0);
          tmem_ld32(taddr + sp * 64 + 32, g1);
          tmem_ld32(taddr + 128 + sp * 64, u0);
          tmem_ld32(taddr + 128 + sp * 64 + 32, u1);
          tmem_ld_wait();
          if (sp == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
          }
          const int cg = n_blk * 128 + sp * 64;            // gate column in C == channel column in the m tensor
          if (row0 < p.M) {
#pragma unroll
            for (int which = 0; which < 3; ++which) {       // 0: gate -> C, 1: up -> C (+I), 2: m -> tmR's tensor
              if (which < 2 && p.aux == nullptr) continue;  // inference: only m is wanted
              if (lane == 0) tma_store_wait_read<1>();
              __syncwarp();
              const uint32_t row_s = my_buf_s + buf * EPI_BUF_BYTES + lane * 128;
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                uint32_t o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int idx = ch * 8 + 2 * j;
                  const uint32_t gp = idx < 32 ? pack_bf16x2(__uint_as_float(g0[idx]), __uint_as_float(g0[idx + 1]))
                                               : pack_bf16x2(__uint_as_float(g1[idx - 32]), __uint_as_float(g1[idx - 31]));
                  const uint32_t up = idx < 32 ? pack_bf16x2(__uint_as_float(u0[idx]), __uint_as_float(u0[idx + 1]))
                                               : pack_bf16x2(__uint_as_float(u1[idx - 32]), __uint_as_float(u1[idx - 31]));
                  o[j] = which == 0 ? gp : (which == 1 ? up : swiglu_fwd_pair(gp, up));
                }
                st_shared_v4(row_s + ((ch ^ (lane & 7)) << 4), make_uint4(o[0], o[1], o[2], o[3]));
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (which == 2) tma_store_2d(&tmR, my_buf + buf * EPI_BUF_BYTES, cg, row0);
                else tma_store_2d(&tmC, my_buf + buf * EPI_BUF_BYTES, cg + which * p.swiglu_inter, row0);
                tma_store_commit();
              }
              buf ^= 1;
            }
          }
        }
        if (++as == 2) { as = 0; aphase ^= 1u; }
        continue;
      }
#pragma unroll 1
      for (int slab = 0; slab < BN / EPI_BOX_COLS; ++slab) {
        const int c0 = col0 + slab * EPI_BOX_COLS;
        const uint32_t sbuf = my_buf_s + buf * EPI_BUF_BYTES;
        // the TMA store issued two slabs ago from this buffer must have finished reading it
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        const bool live = (row0 < p.M) && (c0 < p.N);
        if (p.epi_mode && p.epi_mode != 3 && live && lane == 0) {
          mbar_arrive_expect_tx(&epi_bar[q], EPI_BUF_BYTES);
          tma_load_2d(&tmR, &epi_bar[q], my_buf + buf * EPI_BUF_BYTES, c0, row0);
        }
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr + slab * EPI_BOX_COLS, v0);
        tmem_ld32(taddr + slab * EPI_BOX_COLS + 32, v1);
        tmem_ld_wait();
        if (slab == BN / EPI_BOX_COLS - 1) {
          // accumulator fully read: hand the TMEM stage back to the MMA warp of the leader CTA
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(&tmem_empty[as]);
        }
        if (live && p.epi_mode == 3) {
          // split-K: fp32 partial tile -> swizzled staging -> TMA reduce-add into the fp32 workspace (two 32-col boxes)
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int hb = buf ^ half;
            if (half == 1) {
              if (lane == 0) tma_store_wait_read<1>();
              __syncwarp();
            }
            const uint32_t row_s = my_buf_s + hb * EPI_BUF_BYTES + lane * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const uint32_t* v = half ? v1 : v0;
              st_shared_v4(row_s + ((c ^ (lane & 7)) << 4), make_uint4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && c0 + half * 32 < p.N) {
              tma_reduce_add_2d(&tmR, my_buf + hb * EPI_BUF_BYTES, c0 + half * 32, row0);
              tma_store_commit();
            }
          }
        } else if (live) {
          if (p.epi_mode) { mbar_wait(&epi_bar[q], ephase); }
          const uint32_t row_s = sbuf + lane * 128;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int idx = ch * 8 + j;
              f[j] = __uint_as_float(idx < 32 ? v0[idx] : v1[idx - 32]);
            }
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int col = c0 + ch * 8 + j;
                f[j] += (col < p.N) ? __ldg(p.bias + col) : 0.f;
              }
            }
            const uint32_t addr = row_s + ((ch ^ (lane & 7)) << 4);
            if (p.epi_mode) {
              if (p.epi_mode == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = bf16_round(f[j]);
              }
              const uint4 old = ld_shared_v4(addr);
              const float2 o0 = unpack_bf16x2(old.x), o1 = unpack_bf16x2(old.y), o2 = unpack_bf16x2(old.z),
                           o3 = unpack_bf16x2(old.w);
              f[0] += o0.x; f[1] += o0.y; f[2] += o1.x; f[3] += o1.y;
              f[4] += o2.x; f[5] += o2.y; f[6] += o3.x; f[7] += o3.y;
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            st_shared_v4(addr, o);
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, my_buf + buf * EPI_BUF_BYTES, c0, row0);
            tma_store_commit();
          }
          if (p.epi_mode) ephase ^= 1u;
        }
        if (p.epi_mode != 3) buf ^= 1;
      }
      if (++as == 2) { as = 0; aphase ^= 1u; }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  // ===================================== teardown =====================================
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  }
}

template <int CG, bool A_MN, bool B_MN>
static int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
                  const Params& p, int max_ctas, cudaStream_t stream) {
  using C = Cfg<CG>;
  auto kern = gemm_bf16_kernel<CG, A_MN, B_MN>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm smem=%d): %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.split_k;
  int sms = sm_count();
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
  int pairs = sms / CG;
  if (pairs > num_tiles) pairs = num_tiles;
  if (pairs < 1) pairs = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(pairs * CG);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  Params pp = p;
  // Raster group: the tiles of GM consecutive m-tiles are walked n-column by n-column, so the group's A panels (GM x BM*CG x K
  // elements) are re-used out of L2 by every wave while the B panels stream through once per group: DRAM traffic ~ A + B * (m-tiles /
  // GM).  GM = 16 halves the B re-reads as long as the group's A panels stay resident (<= ~36 MB of the 126 MB L2, i.e. K <= 4608:
  // the forward projections and the K = 4096 dX GEMMs); the K-long shapes keep 8, where a wave of 74 CTA pairs is closest to square.
  // Full-step sweep of a FIXED value: 8 -> 1460-1467 ms, 16 -> 1455, 4 -> 1512, 32 -> 1522; per-shape choice against fixed 8 on
  // another box: 1381.9 / 1374.6 ms against 1387.5 / 1385.7 (profiles/r02_bench_gemm_raster_gm_sweep.log).
  static const int gm_env = []() { const char* e = getenv("B200_GEMM_GM"); return e ? atoi(e) : 0; }();
  if (gm_env > 0) pp.gm = gm_env;
  else pp.gm = (16ll * BM * CG * p.K * 2 <= (36ll << 20) || p.N > 65536 || p.K > 65536) ? 16 : 8;
  // (thresholds of 52 / 72 / 120 MB measured within noise of 36; the vocabulary-sized lm_head GEMMs take 16 by measurement:
  //  per-shape DRAM bytes and times for GM = 4 ... 32 in profiles/r02_gemm_traffic_gm_sweep.txt)
  pp.b_prefetch = 0;
  pp.l2_prefetch_kb = 0;
  if (pdl_enabled()) {
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
    pp.b_prefetch = 1;
    // at most ~64 MB of weights resident in L2 ahead of the loads (the L2 is 126 MB and also holds the activations)
    const long long per_kb = static_cast<long long>(C::B_COLS) * BK * 2 * pairs * CG;
    long long kbs = (static_cast<long long>(l2_prefetch_mb()) << 20) / per_kb;
    pp.l2_prefetch_kb = static_cast<int>(kbs > 64 ? 64 : kbs);
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, pp);
  if (e != cudaSuccess) {
    set_last_error("gemm launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return 0;
}

}  // namespace gemm
}  // namespace b200

extern "C" int b200_gemm_bf16_ex(const void* A, const void* B, void* C, const float* bias, const void* residual,
                                 int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int64_t ldr,
                                 int a_mn_major, int b_mn_major, int accumulate, int cta_group, int max_ctas,
                                 cudaStream_t stream) {
  using namespace b200;
  using namespace b200::gemm;
  B200_CHECK_ARG(A && B && C, "gemm: null pointer");
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm: non-positive dimension M=%lld N=%lld K=%lld", (long long)M,
                 (long long)N, (long long)K);
  B200_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0, "gemm: leading dimensions must be multiples of 8");
  B200_CHECK_ARG(cta_group == 1 || cta_group == 2, "gemm: cta_group must be 1 or 2");
  B200_CHECK_ARG(!(residual && accumulate), "gemm: residual and accumulate are mutually exclusive");
  B200_CHECK_ARG(!residual || ldr % 8 == 0, "gemm: ldr must be a multiple of 8");
  B200_CHECK_ARG(M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31), "gemm: dimension too large");

  CUtensorMap tmA, tmB, tmC, tmR;
  int rc;
  {
    // A: K-major  -> stored [M, K], dims {K, M}, box {64, 128}
    //    MN-major -> stored [K, M], dims {M, K}, box {64, 64}
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (a_mn_major) { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = BK; }
    else            { dims[0] = K; dims[1] = M; box[0] = BK; box[1] = 128; }
    strides[0] = static_cast<uint64_t>(lda) * 2;
    if ((rc = encode_tmap_bf16(&tmA, A, 2, dims, strides, box)) != 0) return rc;
  }
  {
    // B: K-major  -> stored [N, K], dims {K, N}, box {64, 128}
    //    MN-major -> stored [K, N], dims {N, K}, box {64, 64}
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (b_mn_major) { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = BK; }
    else            { dims[0] = K; dims[1] = N; box[0] = BK; box[1] = 128; }
    strides[0] = static_cast<uint64_t>(ldb) * 2;
    if ((rc = encode_tmap_bf16(&tmB, B, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldc) * 2};
    uint32_t box[2] = {EPI_BOX_COLS, EPI_BOX_ROWS};
    if ((rc = encode_tmap_bf16(&tmC, C, 2, dims, strides, box)) != 0) return rc;
    if (residual) {
      strides[0] = static_cast<uint64_t>(ldr) * 2;
      if ((rc = encode_tmap_bf16(&tmR, residual, 2, dims, strides, box)) != 0) return rc;
    } else {
      tmR = tmC;
    }
  }
  Params p;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.num_m_tiles = static_cast<int>((M + BM * cta_group - 1) / (BM * cta_group));
  p.num_n_tiles = static_cast<int>((N + BN - 1) / BN);
  p.epi_mode = residual ? 2 : (accumulate ? 1 : 0);
  p.split_k = 1;
  p.kb_per_split = static_cast<int>((K + BK - 1) / BK);
  p.bias = bias;
  p.swiglu_inter = 0;

#define B200_GEMM_DISPATCH(CG)                                                                          \
  do {                                                                                                  \
    if (a_mn_major && b_mn_major) return launch<CG, true, true>(tmA, tmB, tmC, tmR, p, max_ctas, stream);    \
    if (a_mn_major) return launch<CG, true, false>(tmA, tmB, tmC, tmR, p, max_ctas, stream);                 \
    if (b_mn_major) return launch<CG, false, true>(tmA, tmB, tmC, tmR, p, max_ctas, stream);                 \
    return launch<CG, false, false>(tmA, tmB, tmC, tmR, p, max_ctas, stream);                                \
  } while (0)
  if (cta_group == 2) B200_GEMM_DISPATCH(2);
  B200_GEMM_DISPATCH(1);
#undef B200_GEMM_DISPATCH
}

// gate|up projection + SwiGLU in one kernel (training forward of LlamaMLP, llama/modeling.py:632-652 with fuse_attention_ffn):
//   GU[M, 2I] = bf16(X[M, K] * W[K, 2I])   (gate columns [0, I), up columns [I, 2I): kept for the backward)
//   Mout[M, I] = bf16( silu(GU[:, c]) * GU[:, I + c] )
// The 256-column tile of the tcgen05 GEMM is formed from 128 gate columns and the 128 up columns of the same channels (two TMA
// boxes at different column coordinates of the SAME row-major weight), so the epilogue holds both halves of every channel.
extern "C" int b200_gemm_swiglu_bf16(const void* X, const void* W, void* GU, void* Mout, int64_t M, int64_t inter, int64_t K,
                                     int64_t ldx, int64_t ldw, int64_t ldgu, int64_t ldm, int cta_group, cudaStream_t stream) {
  using namespace b200;
  using namespace b200::gemm;
  B200_CHECK_ARG(X && W && Mout, "gemm_swiglu: null pointer");     // GU may be null: gate|up are then not written (inference)
  B200_CHECK_ARG(M > 0 && inter > 0 && K > 0 && inter % 128 == 0, "gemm_swiglu: intermediate size must be a multiple of 128 (got %lld)",
                 (long long)inter);
  B200_CHECK_ARG(ldx % 8 == 0 && ldw % 8 == 0 && ldgu % 8 == 0 && ldm % 8 == 0, "gemm_swiglu: leading dimensions must be multiples of 8");
  B200_CHECK_ARG(cta_group == 1 || cta_group == 2, "gemm_swiglu: cta_group must be 1 or 2");
  B200_CHECK_ARG(M < (1ll << 31) && inter < (1ll << 30) && K < (1ll << 31), "gemm_swiglu: dimension too large");
  const int64_t N = 2 * inter;
  CUtensorMap tmA, tmB, tmC, tmM;
  int rc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)}, strides[1] = {static_cast<uint64_t>(ldx) * 2};
    uint32_t box[2] = {BK, 128};
    if ((rc = encode_tmap_bf16(&tmA, X, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(K)}, strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {64, BK};
    if ((rc = encode_tmap_bf16(&tmB, W, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)}, strides[1] = {static_cast<uint64_t>(ldgu) * 2};
    uint32_t box[2] = {EPI_BOX_COLS, EPI_BOX_ROWS};
    if (GU) {
      if ((rc = encode_tmap_bf16(&tmC, GU, 2, dims, strides, box)) != 0) return rc;
    } else {
      memset(&tmC, 0, sizeof(tmC));
    }
    dims[0] = static_cast<uint64_t>(inter);
    strides[0] = static_cast<uint64_t>(ldm) * 2;
    if ((rc = encode_tmap_bf16(&tmM, Mout, 2, dims, strides, box)) != 0) return rc;
  }
  Params p;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(N);
  p.K = static_cast<int>(K);
  p.num_m_tiles = static_cast<int>((M + BM * cta_group - 1) / (BM * cta_group));
  p.num_n_tiles = static_cast<int>(inter / 128);
  p.epi_mode = 4;
  p.aux = GU;
  p.split_k = 1;
  p.kb_per_split = static_cast<int>((K + BK - 1) / BK);
  p.bias = nullptr;
  p.swiglu_inter = static_cast<int>(inter);
  if (cta_group == 2) return launch<2, false, true>(tmA, tmB, tmC, tmM, p, 0, stream);
  return launch<1, false, true>(tmA, tmB, tmC, tmM, p, 0, stream);
}

// down-projection dX GEMM + SwiGLU backward in one kernel (backward of LlamaMLP, llama/modeling.py:632-652):
//   d(m)[M, I] = dY[M, h] * W_down[I, h]^T   (never written),   DGU[M, 2I] = [ d(m) * up * silu'(gate) | d(m) * silu(gate) ]
// GU is the saved gate|up projection [M, 2I].  Bit-identical to b200_gemm_bf16 (dX) followed by b200_swiglu_bwd.  I % 64 == 0.
extern "C" int b200_gemm_swiglu_bwd_bf16(const void* dY, const void* Wdown, const void* GU, void* DGU, int64_t M, int64_t inter,
                                         int64_t K, int64_t lddy, int64_t ldw, int64_t ldgu, int64_t lddgu, int cta_group,
                                         cudaStream_t stream) {
  using namespace b200;
  using namespace b200::gemm;
  B200_CHECK_ARG(dY && Wdown && GU && DGU, "gemm_swiglu_bwd: null pointer");
  B200_CHECK_ARG(M > 0 && inter > 0 && K > 0 && inter % 64 == 0, "gemm_swiglu_bwd: intermediate size must be a multiple of 64 (got %lld)",
                 (long long)inter);
  B200_CHECK_ARG(lddy % 8 == 0 && ldw % 8 == 0 && ldgu % 8 == 0 && lddgu % 8 == 0, "gemm_swiglu_bwd: leading dimensions must be multiples of 8");
  B200_CHECK_ARG(cta_group == 1 || cta_group == 2, "gemm_swiglu_bwd: cta_group must be 1 or 2");
  B200_CHECK_ARG(M < (1ll << 31) && inter < (1ll << 30) && K < (1ll << 31), "gemm_swiglu_bwd: dimension too large");
  CUtensorMap tmA, tmB, tmC, tmG;
  int rc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)}, strides[1] = {static_cast<uint64_t>(lddy) * 2};
    uint32_t box[2] = {BK, 128};
    if ((rc = encode_tmap_bf16(&tmA, dY, 2, dims, strides, box)) != 0) return rc;
  }
  {   // W_down stored [I, h] = [N, K] row-major: K-major B operand
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(inter)}, strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {BK, 128};
    if ((rc = encode_tmap_bf16(&tmB, Wdown, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(2 * inter), static_cast<uint64_t>(M)}, strides[1] = {static_cast<uint64_t>(lddgu) * 2};
    uint32_t box[2] = {16, EPI_BOX_ROWS};            // the mode-5 epilogue works on slabs of 16 channels: 32 rows x 32 bytes, unswizzled
    if ((rc = encode_tmap_bf16_linear(&tmC, DGU, 2, dims, strides, box)) != 0) return rc;
    strides[0] = static_cast<uint64_t>(ldgu) * 2;
    if ((rc = encode_tmap_bf16_linear(&tmG, GU, 2, dims, strides, box)) != 0) return rc;
  }
  Params p;
  p.M = static_cast<int>(M);
  p.N = static_cast<int>(inter);
  p.K = static_cast<int>(K);
  p.num_m_tiles = static_cast<int>((M + BM * cta_group - 1) / (BM * cta_group));
  p.num_n_tiles = static_cast<int>((inter + BN - 1) / BN);
  p.epi_mode = 5;
  p.aux = GU;
  p.ld_aux = ldgu;
  p.split_k = 1;
  p.kb_per_split = static_cast<int>((K + BK - 1) / BK);
  p.bias = nullptr;
  p.swiglu_inter = static_cast<int>(inter);
  if (cta_group == 2) return launch<2, false, false>(tmA, tmB, tmC, tmG, p, 0, stream);
  return launch<1, false, false>(tmA, tmB, tmC, tmG, p, 0, stream);
}

namespace b200 {
namespace skinny {
int gemm_skinny_f32(const void* X, const void* W, void* workspace, int64_t M, int64_t N, int64_t K, int64_t ldx, int64_t ldw,
                    bool w_kmajor, int split_k, cudaStream_t stream);   // gemm_skinny.cu
}
namespace gemm {
// out[m, n] = bf16(ws[m, n] + bias[n]) ; ws is re-zeroed for the next split-K GEMM that uses it
__global__ void splitk_finish_kernel(float* __restrict__ ws, const float* __restrict__ bias, bf16* __restrict__ out,
                                     int64_t M, int64_t N, int64_t ldc) {
  const int64_t nch = N >> 3;
  const int64_t total = M * nch;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / nch, c = i % nch;
    float4* src = reinterpret_cast<float4*>(ws + r * N) + 2 * c;
    float4 a = src[0], b = src[1];
    src[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    src[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias != nullptr) {
      const float4* bp = reinterpret_cast<const float4*>(bias) + 2 * c;
      const float4 b0 = __ldg(bp), b1 = __ldg(bp + 1);
      a.x += b0.x; a.y += b0.y; a.z += b0.z; a.w += b0.w;
      b.x += b1.x; b.y += b1.y; b.z += b1.z; b.w += b1.w;
    }
    uint4 o;
    o.x = pack_bf16x2(a.x, a.y); o.y = pack_bf16x2(a.z, a.w);
    o.z = pack_bf16x2(b.x, b.y); o.w = pack_bf16x2(b.z, b.w);
    *(reinterpret_cast<uint4*>(out + r * ldc) + c) = o;
  }
}
}  // namespace gemm
}  // namespace b200

extern "C" int64_t b200_gemm_splitk_workspace_bytes(int64_t M, int64_t N) { return M * N * 4; }

// Weight-streaming ("skinny") GEMM for the decode step: M <= 128 tokens, the weight matrix dominates the traffic, so K is
// split across CTAs until the persistent grid covers every SM; fp32 partial tiles are reduced in L2 by the TMA unit.
extern "C" int b200_gemm_bf16_splitk(const void* A, const void* B, void* C, const float* bias, void* workspace, int64_t M,
                                     int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major,
                                     int b_mn_major, int split_k, cudaStream_t stream) {
  using namespace b200;
  using namespace b200::gemm;
  B200_CHECK_ARG(A && B && workspace, "gemm_splitk: null pointer");
  B200_CHECK_ARG(M > 0 && N > 0 && K > 0 && N % 8 == 0, "gemm_splitk: bad dimensions (N must be a multiple of 8)");
  B200_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && ldc % 8 == 0, "gemm_splitk: leading dimensions must be multiples of 8");
  auto finish = [&]() -> int {
    if (C == nullptr) return 0;   // C == NULL: the consumer kernel reads (and re-zeroes) the fp32 workspace itself
    const int64_t total = M * (N / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > sm_count() * 8) blocks = sm_count() * 8;
    splitk_finish_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<float*>(workspace), bias,
                                                                           static_cast<bf16*>(C), M, N, ldc);
    return check_launch("gemm_splitk(finish)");
  };
  if (!a_mn_major && M <= 128 && skinny_gemm_impl() >= 1) {
    // decode-step shapes: swapped-operand weight-streaming kernel, two CTAs per SM (gemm_skinny.cu)
    int r = skinny::gemm_skinny_f32(A, B, workspace, M, N, K, lda, ldb, /*w_kmajor=*/!b_mn_major, split_k, stream);
    return r ? r : finish();
  }
  CUtensorMap tmA, tmB, tmC, tmF;
  int rc;
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (a_mn_major) { dims[0] = M; dims[1] = K; box[0] = 64; box[1] = BK; }
    else            { dims[0] = K; dims[1] = M; box[0] = BK; box[1] = 128; }
    strides[0] = static_cast<uint64_t>(lda) * 2;
    if ((rc = encode_tmap_bf16(&tmA, A, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2], strides[1];
    uint32_t box[2];
    if (b_mn_major) { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = BK; }
    else            { dims[0] = K; dims[1] = N; box[0] = BK; box[1] = 128; }
    strides[0] = static_cast<uint64_t>(ldb) * 2;
    if ((rc = encode_tmap_bf16(&tmB, B, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldc) * 2};
    uint64_t fstrides[1] = {static_cast<uint64_t>(N) * 4};
    uint32_t fbox[2] = {32, 32};
    if ((rc = encode_tmap_f32(&tmF, workspace, 2, dims, fstrides, fbox)) != 0) return rc;
    tmC = tmF;   // the bf16 output map is unused by epilogue mode 3
    (void)strides;
  }
  Params p;
  p.swiglu_inter = 0;
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  p.num_m_tiles = static_cast<int>((M + BM - 1) / BM);
  p.num_n_tiles = static_cast<int>((N + BN - 1) / BN);
  const int num_kb = static_cast<int>((K + BK - 1) / BK);
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  if (split_k <= 0) {
    // smallest split that gives every SM at least one work item, capped so each item keeps >= 4 k-blocks
    const int sms = sm_count();
    split_k = (sms + tiles - 1) / tiles;
    if (split_k > num_kb / 4) split_k = num_kb / 4;
    if (split_k < 1) split_k = 1;
  }
  p.kb_per_split = (num_kb + split_k - 1) / split_k;
  p.split_k = (num_kb + p.kb_per_split - 1) / p.kb_per_split;   // no empty ranges
  p.epi_mode = 3;
  p.bias = nullptr;
  // `workspace` must be all-zero on entry; the finish kernel leaves it zeroed again (no memset per GEMM).
  if (a_mn_major && b_mn_major) rc = launch<1, true, true>(tmA, tmB, tmC, tmF, p, 0, stream);
  else if (a_mn_major) rc = launch<1, true, false>(tmA, tmB, tmC, tmF, p, 0, stream);
  else if (b_mn_major) rc = launch<1, false, true>(tmA, tmB, tmC, tmF, p, 0, stream);
  else rc = launch<1, false, false>(tmA, tmB, tmC, tmF, p, 0, stream);
  return rc ? rc : finish();
}

extern "C" int b200_gemm_bf16(const void* A, const void* B, void* C, const float* bias, int64_t M, int64_t N, int64_t K,
                              int64_t lda, int64_t ldb, int64_t ldc, int a_mn_major, int b_mn_major, int accumulate,
                              cudaStream_t stream) {
  return b200_gemm_bf16_ex(A, B, C, bias, nullptr, M, N, K, lda, ldb, ldc, 0, a_mn_major, b_mn_major, accumulate,
                           /*cta_group=*/2, /*max_ctas=*/0, stream);
}
