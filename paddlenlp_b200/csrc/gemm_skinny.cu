// Weight-streaming GEMM for the decode step (M <= 128 token rows): swapped operands, two CTAs per SM.
//
//   ws[M, N] (fp32, zero on entry) += X[M, K] * W[K, N]        X bf16 [tokens, K];  W bf16 [K, N] or [N, K] ("trans_b")
//
// Same contract as epilogue mode 3 of gemm_tcgen05.cu (b200_gemm_bf16_splitk): the consumer kernel rounds the fp32 sums once.
// Why a second kernel: with 64 token rows the 128x256 tile of the training GEMM wastes half of every UMMA and needs ~200 KB
// of shared memory per CTA, so consecutive GEMMs of the decode chain can never overlap (one CTA per SM, ~7 us of prologue /
// ramp / drain per 12-30 us kernel).  Here the WEIGHT tile is the 128-row M operand and the tokens are the N operand:
//   S^T-style tile  acc[128 features, NT tokens] = W_tile^T (128 x 64k)  x  X_tile^T (64k x NT),  NT = 64 or 128
//   stage = 16 KB of weights + NT*128 B of activations -> 96 KB ring, 64/128 TMEM columns: two CTAs per SM, so under
//   programmatic dependent launch the next GEMM is resident, has its weight tiles in flight and the rest of its slice
//   prefetched to L2 while the previous kernels of the chain are still draining;
//   K is split over CTAs (work item = feature tile x K range), partial tiles leave through per-warp transposed fp32
//   staging (reusing the drained ring) and TMA reduce-add into the [tokens, N] workspace.
#include "../../include/b200nlp.h"
#include "common.cuh"
#include "host_util.h"
#include <stdlib.h>

namespace b200 {
namespace skinny {

constexpr int BF = 128;    // output features per work item (UMMA M)
constexpr int BK = 64;     // k per stage (one 128-byte swizzle row)
constexpr int UK = 16;
constexpr int NUM_THREADS = 192;
constexpr int W_BYTES = BF * BK * 2;   // 16 KB

template <int NT>
struct Cfg {
  static constexpr int X_BYTES = NT * BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
  static constexpr int STAGES = (NT == 64) ? 4 : 3;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;              // 96 KB
  static constexpr int SMEM_BYTES = RING_BYTES + 256 + 1024;
  static_assert(4 * NT * 32 * 4 <= RING_BYTES, "epilogue staging must fit in the drained ring");
};

struct Params {
  int M, N, K;
  int split_k, kb_per_split;
  int f_tiles;          // feature tiles (work item = feature tile x K range)
  int split_major;      // item order: 1 = consecutive CTAs take consecutive feature tiles of the same K range
  int w_prefetch;       // PDL: weight tiles of the first stages are loaded before griddepcontrol.wait
  int l2_prefetch_kb;   // PDL: further weight k-blocks prefetched into L2 before the wait
  bf16* act;            // SWIGLU epilogue: output [M, inter] bf16
  int inter;
};

// SWIGLU (ffn1 of the decode step, no K split): feature tile j pairs the gate columns [64j, 64j+64) with the up columns
// [I + 64j, I + 64j + 64) of the reference-layout weight [K, 2I] (two 64-column TMA boxes per stage, no re-laid-out copy), so
// accumulator lanes 0..63 are gate channels and lanes 64..127 the matching up channels.  The epilogue rounds both to bf16 (the Linear
// output rounding), the gate warps and the up warps swap one half of their tokens through the drained ring so that all four warps
// evaluate silu(g) * u (swiglu_fwd_pair, common.cuh), the [tokens x 64 channels] result is staged row-major in shared memory and
// leaves as ONE TMA store into act[M, inter]: no fp32 workspace, no separate activation kernel, no 2-byte scattered stores.
template <int NT, bool W_KMAJOR, bool SWIGLU = false>
__global__ void __launch_bounds__(NUM_THREADS, 2)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                   const __grid_constant__ CUtensorMap tmF, const Params p) {
  using C = Cfg<NT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::RING_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;    // [STAGES]
  uint64_t* acc_full = bars + 2 * C::STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int item = blockIdx.x;
  int f_tile, split;
  if (p.split_major) { split = item / p.f_tiles; f_tile = item - split * p.f_tiles; }
  else { f_tile = item / p.split_k; split = item - f_tile * p.split_k; }
  const int f0 = f_tile * BF;
  const int num_kb_total = (p.K + BK - 1) / BK;
  const int kb0 = split * p.kb_per_split;
  const int nkb = min(num_kb_total, kb0 + p.kb_per_split) - kb0;      // >= 1 by construction
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmF);
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      auto load_w = [&](int s, int kb) {
        uint8_t* sw = smem + s * C::STAGE_BYTES;
        const int k0 = kb * BK;
        if constexpr (W_KMAJOR) {
          tma_load_2d(&tmW, &full_bar[s], sw, k0, f0);                      // [128 features x 64 k], k contiguous
        } else if constexpr (SWIGLU) {
          tma_load_2d(&tmW, &full_bar[s], sw, f_tile * 64, k0);                         // 64 gate columns
          tma_load_2d(&tmW, &full_bar[s], sw + 64 * BK * 2, p.inter + f_tile * 64, k0);  // the up columns of the same channels
        } else {
          tma_load_2d(&tmW, &full_bar[s], sw, f0, k0);                      // two [64 k x 64 features] boxes
          tma_load_2d(&tmW, &full_bar[s], sw + 64 * BK * 2, f0 + 64, k0);
        }
      };
      int pre = 0;
      if (p.w_prefetch) {
        // the weights do not depend on the previous kernel: put them in flight (and the rest of the slice into L2) first
        pre = min(C::STAGES, nkb);
        for (int s = 0; s < pre; ++s) {
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          load_w(s, kb0 + s);
        }
        const int l2_end = min(nkb, pre + p.l2_prefetch_kb);
        for (int s = pre; s < l2_end; ++s) {
          const int k0 = (kb0 + s) * BK;
          if constexpr (W_KMAJOR) {
            tma_prefetch_l2_2d(&tmW, k0, f0);
          } else if constexpr (SWIGLU) {
            tma_prefetch_l2_2d(&tmW, f_tile * 64, k0);
            tma_prefetch_l2_2d(&tmW, p.inter + f_tile * 64, k0);
          } else {
            tma_prefetch_l2_2d(&tmW, f0, k0);
            tma_prefetch_l2_2d(&tmW, f0 + 64, k0);
          }
        }
      }
      pdl_wait();
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        if (i >= pre) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
          load_w(stage, kb0 + i);
        }
        tma_load_2d(&tmX, &full_bar[stage], smem + stage * C::STAGE_BYTES + W_BYTES, (kb0 + i) * BK, 0);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BF, NT, !W_KMAJOR, false);
      constexpr uint32_t w_lbo = W_KMAJOR ? 16 : 64 * BK * 2, w_adv = W_KMAJOR ? UK * 2 : UK * 128;
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sw = smem_u32(smem + stage * C::STAGE_BYTES);
        const uint32_t sx = sw + W_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UK; ++k)
          umma_ss<1>(tmem_base, umma_desc_sw128(sw + k * w_adv, w_lbo, 1024), umma_desc_sw128(sx + k * UK * 2, 16, 1024),
                     idesc, (i > 0 || k > 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
        if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
      }
      umma_commit(acc_full);
    }
  } else {
    // ===================================== epilogue =====================================
    pdl_wait();                                   // the workspace belongs to the previous kernels until now
    const int q = warp & 3;                       // TMEM lane quadrant: features f0 + 32q .. +31
    mbar_wait(acc_full, 0);                       // every MMA has completed: all TMA loads landed, the ring is free
    tc_fence_after();
    uint8_t* stage_buf = smem + q * (NT * 32 * 4);
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if constexpr (SWIGLU) {
      static_assert(!SWIGLU || NT == 64, "SWIGLU epilogue: 64 token columns");
      uint32_t* s_x = reinterpret_cast<uint32_t*>(smem);        // [32 token pairs][64 channels] bf16x2: the swapped halves (8 KB)
      bf16* s_m = reinterpret_cast<bf16*>(smem + 8192);         // [64 tokens][64 channels] bf16, row-major = the TMA store box (8 KB)
      uint32_t pk[32];                                          // this lane's channel, token pairs (2i, 2i+1), rounded to bf16
      {
        uint32_t v[2][32];
        tmem_ld32(taddr, v[0]);
        tmem_ld32(taddr + 32, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          pk[i] = pack_bf16x2(__uint_as_float(v[i >> 4][(2 * i) & 31]), __uint_as_float(v[i >> 4][(2 * i + 1) & 31]));
      }
      const bool is_gate = q < 2;
      const int chl = (q & 1) * 32 + lane;                      // channel within the tile
      // gate warps keep tokens 0..31 and hand their gates of tokens 32..63 to the up warps; the up warps do the opposite
#pragma unroll
      for (int i = 0; i < 16; ++i) s_x[(is_gate ? 16 + i : i) * 64 + chl] = pk[is_gate ? 16 + i : i];
      named_bar_sync(1, 128);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int tp = is_gate ? i : 16 + i;                    // token pair
        const uint32_t other = s_x[tp * 64 + chl];
        const uint32_t m2 = is_gate ? swiglu_fwd_pair(pk[tp], other) : swiglu_fwd_pair(other, pk[tp]);
        reinterpret_cast<unsigned short*>(s_m)[(2 * tp) * 64 + chl] = static_cast<unsigned short>(m2 & 0xffffu);
        reinterpret_cast<unsigned short*>(s_m)[(2 * tp + 1) * 64 + chl] = static_cast<unsigned short>(m2 >> 16);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (threadIdx.x == 64) {                                  // rows >= M are clipped by the tensor map
        tma_store_2d(&tmF, s_m, f_tile * 64, 0);
        tma_store_commit();
        tma_store_wait<0>();
      }
    } else if (f0 + q * 32 < p.N) {
#pragma unroll
      for (int ch = 0; ch < NT / 32; ++ch) {
        if (ch * 32 >= p.M) break;                // warp-uniform
        uint32_t v[32];
        tmem_ld32(taddr + ch * 32, v);
        tmem_ld_wait();
        // transpose through smem: staging row = token, 32 features (128 B) per row, 128B-swizzled like the fp32 tensor map
        const uint32_t base = smem_u32(stage_buf + ch * 4096) + (lane & 3) * 4;
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          const uint32_t addr = base + t * 128 + ((((lane >> 2) ^ (t & 7))) << 4);
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v[t]) : "memory");
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_reduce_add_2d(&tmF, stage_buf + ch * 4096, f0 + q * 32, ch * 32);
          tma_store_commit();
        }
      }
      if (lane == 0) tma_store_wait<0>();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, NT);
  }
}

// ------------------------------------------------------------------------------------------------
// Stream-K variant (NT = 64): when feature tiles x K ranges do not fill one wave of 2 CTAs/SM evenly (o-proj: 256 of 296 slots;
// ffn1 / lm_head: hundreds of tiles), the (feature tile, k-block) units are cut into gridDim.x equal contiguous runs instead.  A
// CTA's run may cross feature-tile boundaries, so it walks 1-3 SEGMENTS (tile, k range): the TMA ring runs on across segments,
// the accumulator ping-pongs between two TMEM stages, every segment leaves through the same transposed TMA reduce-add.
// ------------------------------------------------------------------------------------------------
struct SKCfg {
  static constexpr int NT = 64;
  static constexpr int X_BYTES = NT * BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;                // 24 KB
  static constexpr int STAGES = 3;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES;              // 72 KB
  static constexpr int EPI_BYTES = 4 * 4096;                           // one 32x32 fp32 staging buffer per epilogue warp
  static constexpr int SMEM_BYTES = RING_BYTES + EPI_BYTES + 256 + 1024;
};

struct Seg { int f0, kb0, nkb; };
__device__ __forceinline__ Seg next_seg(long long& u, long long u1, int num_kb) {
  Seg sg;
  const int f_tile = static_cast<int>(u / num_kb);
  sg.kb0 = static_cast<int>(u - static_cast<long long>(f_tile) * num_kb);
  sg.nkb = static_cast<int>(min(static_cast<long long>(num_kb - sg.kb0), u1 - u));
  sg.f0 = f_tile * BF;
  u += sg.nkb;
  return sg;
}

template <bool W_KMAJOR>
__global__ void __launch_bounds__(NUM_THREADS, 2)
gemm_skinny_streamk_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX,
                           const __grid_constant__ CUtensorMap tmF, const Params p) {
  using C = SKCfg;
  constexpr int NT = C::NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi = smem + C::RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + C::EPI_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + C::STAGES;    // [STAGES]
  uint64_t* acc_full = bars + 2 * C::STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;            // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = (p.K + BK - 1) / BK;
  const long long total = static_cast<long long>(p.f_tiles) * num_kb;
  const long long u_begin = total * blockIdx.x / gridDim.x, u_end = total * (blockIdx.x + 1) / gridDim.x;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmF);
    for (int i = 0; i < C::STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, 2 * NT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      auto load_w = [&](int s, int f0, int kb) {
        uint8_t* sw = smem + s * C::STAGE_BYTES;
        const int k0 = kb * BK;
        if constexpr (W_KMAJOR) {
          tma_load_2d(&tmW, &full_bar[s], sw, k0, f0);
        } else {
          tma_load_2d(&tmW, &full_bar[s], sw, f0, k0);
          tma_load_2d(&tmW, &full_bar[s], sw + 64 * BK * 2, f0 + 64, k0);
        }
      };
      // weights first (they do not depend on the previous kernel): the first ring stages, then L2 prefetch of the run's next units
      int pre = 0;
      if (p.w_prefetch) {
        long long u = u_begin;
        int issued_l2 = 0;
        while (u < u_end && (pre < C::STAGES || issued_l2 < p.l2_prefetch_kb)) {
          const Seg sg = next_seg(u, u_end, num_kb);
          for (int i = 0; i < sg.nkb; ++i) {
            if (pre < C::STAGES) {
              mbar_arrive_expect_tx(&full_bar[pre], C::STAGE_BYTES);
              load_w(pre, sg.f0, sg.kb0 + i);
              ++pre;
            } else if (issued_l2 < p.l2_prefetch_kb) {
              const int k0 = (sg.kb0 + i) * BK;
              if constexpr (W_KMAJOR) {
                tma_prefetch_l2_2d(&tmW, k0, sg.f0);
              } else {
                tma_prefetch_l2_2d(&tmW, sg.f0, k0);
                tma_prefetch_l2_2d(&tmW, sg.f0 + 64, k0);
              }
              ++issued_l2;
            }
          }
        }
      }
      pdl_wait();
      int stage = 0, n = 0;
      uint32_t phase = 0;
      for (long long u = u_begin; u < u_end;) {
        const Seg sg = next_seg(u, u_end, num_kb);
        for (int i = 0; i < sg.nkb; ++i, ++n) {
          if (n >= pre) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            load_w(stage, sg.f0, sg.kb0 + i);
          }
          tma_load_2d(&tmX, &full_bar[stage], smem + stage * C::STAGE_BYTES + W_BYTES, (sg.kb0 + i) * BK, 0);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BF, NT, !W_KMAJOR, false);
      constexpr uint32_t w_lbo = W_KMAJOR ? 16 : 64 * BK * 2, w_adv = W_KMAJOR ? UK * 2 : UK * 128;
      int stage = 0, seg = 0;
      uint32_t phase = 0;
      for (long long u = u_begin; u < u_end; ++seg) {
        const Seg sg = next_seg(u, u_end, num_kb);
        const int as = seg & 1;
        mbar_wait(&acc_empty[as], ((seg >> 1) & 1) ^ 1u);        // the epilogue drained this accumulator stage
        tc_fence_after();
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(as * NT);
        for (int i = 0; i < sg.nkb; ++i) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sw = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sx = sw + W_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UK; ++k)
            umma_ss<1>(tacc, umma_desc_sw128(sw + k * w_adv, w_lbo, 1024), umma_desc_sw128(sx + k * UK * 2, 16, 1024), idesc,
                       (i > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&acc_full[as]);
      }
    }
  } else {
    // ===================================== epilogue =====================================
    pdl_wait();
    const int q = warp & 3;
    uint8_t* stage_buf = epi + q * 4096;
    const uint32_t base = smem_u32(stage_buf) + (lane & 3) * 4;
    int seg = 0;
    for (long long u = u_begin; u < u_end; ++seg) {
      const Seg sg = next_seg(u, u_end, num_kb);
      const int as = seg & 1;
      mbar_wait(&acc_full[as], (seg >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * NT);
      uint32_t v[2][32];
      tmem_ld32(taddr, v[0]);
      tmem_ld32(taddr + 32, v[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);               // accumulator stage free for the segment after next
      if (sg.f0 + q * 32 < p.N) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          if (ch * 32 >= p.M) break;
          if (lane == 0) tma_store_wait_read<0>();               // the previous reduce has finished reading the buffer
          __syncwarp();
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            const uint32_t addr = base + t * 128 + ((((lane >> 2) ^ (t & 7))) << 4);
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v[ch][t]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_2d(&tmF, stage_buf, sg.f0 + q * 32, ch * 32);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 2 * NT);
  }
}

template <bool W_KMAJOR>
static int launch_streamk(const CUtensorMap& tmW, const CUtensorMap& tmX, const CUtensorMap& tmF, Params p, int grid,
                          cudaStream_t stream) {
  auto kern = gemm_skinny_streamk_kernel<W_KMAJOR>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SKCfg::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm_skinny_streamk smem=%d): %s", SKCfg::SMEM_BYTES, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  p.w_prefetch = pdl_enabled() ? 1 : 0;
  const long long per_kb = static_cast<long long>(W_BYTES) * grid;
  const long long kbs = (static_cast<long long>(l2_prefetch_mb()) << 20) / per_kb;
  p.l2_prefetch_kb = p.w_prefetch ? static_cast<int>(kbs > 64 ? 64 : kbs) : 0;
  cudaError_t e = launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(NUM_THREADS), SKCfg::SMEM_BYTES, stream, tmW, tmX, tmF, p);
  if (e != cudaSuccess) {
    set_last_error("gemm_skinny_streamk launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return check_launch("gemm_skinny_streamk");
}

template <int NT, bool W_KMAJOR, bool SWIGLU = false>
static int launch(const CUtensorMap& tmW, const CUtensorMap& tmX, const CUtensorMap& tmF, Params p, int items,
                  cudaStream_t stream) {
  using C = Cfg<NT>;
  auto kern = gemm_skinny_kernel<NT, W_KMAJOR, SWIGLU>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(gemm_skinny smem=%d): %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  p.w_prefetch = pdl_enabled() ? 1 : 0;
  // at most ~64 MB of weights parked in L2 ahead of the loads
  const long long per_kb = static_cast<long long>(W_BYTES) * items;
  const long long kbs = (static_cast<long long>(l2_prefetch_mb()) << 20) / per_kb;
  p.l2_prefetch_kb = p.w_prefetch ? static_cast<int>(kbs > 64 ? 64 : kbs) : 0;
  cudaError_t e = launch_pdl(kern, dim3(static_cast<unsigned>(items)), dim3(NUM_THREADS), C::SMEM_BYTES, stream, tmW, tmX, tmF, p);
  if (e != cudaSuccess) {
    set_last_error("gemm_skinny launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return check_launch("gemm_skinny");
}

// ws[M, N] += X[M, K] W ; returns 0 or an error code.  Called by b200_gemm_bf16_splitk for the decode-step shapes.
int gemm_skinny_f32(const void* X, const void* W, void* workspace, int64_t M, int64_t N, int64_t K, int64_t ldx, int64_t ldw,
                    bool w_kmajor, int split_k, cudaStream_t stream) {
  if (!(M > 0 && M <= 128 && N > 0 && K > 0 && N % 8 == 0 && ldx % 8 == 0 && ldw % 8 == 0))
    return fail_arg("gemm_skinny: need 0 < M <= 128, N %% 8 == 0, leading dimensions %% 8 == 0");
  const int NT = M <= 64 ? 64 : 128;
  CUtensorMap tmW, tmX, tmF;
  int rc;
  {
    uint64_t dims[2], strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2];
    if (w_kmajor) { dims[0] = K; dims[1] = N; box[0] = BK; box[1] = BF; }     // W [N, K]
    else          { dims[0] = N; dims[1] = K; box[0] = 64; box[1] = BK; }     // W [K, N]
    if ((rc = encode_tmap_bf16(&tmW, W, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldx) * 2};
    uint32_t box[2] = {BK, static_cast<uint32_t>(NT)};
    if ((rc = encode_tmap_bf16(&tmX, X, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(N) * 4};
    uint32_t box[2] = {32, 32};
    if ((rc = encode_tmap_f32(&tmF, workspace, 2, dims, strides, box)) != 0) return rc;
  }
  Params p = {};
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  const int f_tiles = static_cast<int>((N + BF - 1) / BF);
  const int num_kb = static_cast<int>((K + BK - 1) / BK);
  if (split_k <= 0) {
    // as many work items as fit in one wave of two CTAs per SM, each keeping >= 4 k-blocks
    split_k = (2 * sm_count()) / f_tiles;
    if (split_k > num_kb / 4) split_k = num_kb / 4;
    if (split_k < 1) split_k = 1;
  }
  if (split_k > num_kb) split_k = num_kb;
  p.kb_per_split = (num_kb + split_k - 1) / split_k;
  p.split_k = (num_kb + p.kb_per_split - 1) / p.kb_per_split;   // no empty ranges
  const int items = f_tiles * p.split_k;
  p.f_tiles = f_tiles;
  // Stream-K (equal contiguous runs of (tile, k-block) units per CTA) is available but OFF by default: measured in the decode
  // chain it loses to the plain (tile, K-range) items even where those fill only 256 of 296 slots (o-proj 9.0 vs 8.0 us, ffn1
  // 41.3 vs 40.3 us; profiles/r01_decode_ablation_streamk.log) — the 3-stage ring and the per-segment accumulator hand-over cost
  // more than the idle slots.  B200_SKINNY_STREAMK=1 forces it (NT = 64 only).
  {
    const int slots = 2 * sm_count();
    const long long total = static_cast<long long>(f_tiles) * num_kb;
    const long long waves = (items + slots - 1) / slots;
    const double eff = static_cast<double>(total) / (static_cast<double>(waves) * p.kb_per_split * slots);
    static const int force = []() { const char* e = getenv("B200_SKINNY_STREAMK"); return e ? atoi(e) : -1; }();
    const bool want = force > 0 || skinny_gemm_impl() == 2;
    (void)eff;
    if (NT == 64 && want && total >= 2ll * slots) {
      return w_kmajor ? launch_streamk<true>(tmW, tmX, tmF, p, slots, stream) : launch_streamk<false>(tmW, tmX, tmF, p, slots, stream);
    }
  }
  // measured (tools/decode_ablation.py): with W [K, N] neighbouring CTAs should read neighbouring 256-byte column segments of the
  // same rows (split-major: o-proj -6 %, ffn2 -2 %); with W [N, K] neighbouring K ranges of the same rows are better
  p.split_major = w_kmajor ? 0 : 1;
  if (NT == 64) return w_kmajor ? launch<64, true>(tmW, tmX, tmF, p, items, stream) : launch<64, false>(tmW, tmX, tmF, p, items, stream);
  return w_kmajor ? launch<128, true>(tmW, tmX, tmF, p, items, stream) : launch<128, false>(tmW, tmX, tmF, p, items, stream);
}

// act[M, inter] = bf16(silu(g) * u), g|u = bf16(X W) with W the reference-layout fused weight [K, 2*inter] (gate | up).
int gemm_swiglu_skinny(const void* X, const void* W, void* act, int64_t M, int64_t inter, int64_t K, int64_t ldx, int64_t ldw,
                       int64_t ldact, cudaStream_t stream) {
  if (!(M > 0 && M <= 64 && inter > 0 && inter % 64 == 0 && K > 0 && ldx % 8 == 0 && ldw % 8 == 0 && ldact % 8 == 0))
    return fail_arg("gemm_swiglu_skinny: need 0 < M <= 64, inter %% 64 == 0, leading dimensions %% 8 == 0");
  const int64_t N = 2 * inter;
  CUtensorMap tmW, tmX, tmAct;
  int rc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(K)}, strides[1] = {static_cast<uint64_t>(ldw) * 2};
    uint32_t box[2] = {64, BK};
    if ((rc = encode_tmap_bf16(&tmW, W, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(inter), static_cast<uint64_t>(M)}, strides[1] = {static_cast<uint64_t>(ldact) * 2};
    uint32_t box[2] = {64, 64};
    if ((rc = encode_tmap_bf16_linear(&tmAct, act, 2, dims, strides, box)) != 0) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
    uint64_t strides[1] = {static_cast<uint64_t>(ldx) * 2};
    uint32_t box[2] = {BK, 64};
    if ((rc = encode_tmap_bf16(&tmX, X, 2, dims, strides, box)) != 0) return rc;
  }
  Params p = {};
  p.M = static_cast<int>(M); p.N = static_cast<int>(N); p.K = static_cast<int>(K);
  const int num_kb = static_cast<int>((K + BK - 1) / BK);
  p.split_k = 1;
  p.kb_per_split = num_kb;
  p.f_tiles = static_cast<int>(N / BF);
  p.split_major = 0;
  p.act = static_cast<bf16*>(act);
  p.inter = static_cast<int>(inter);
  return launch<64, false, true>(tmW, tmX, tmAct, p, p.f_tiles, stream);  // tmF = the activation map
}

}  // namespace skinny
}  // namespace b200

extern "C" int b200_gemm_swiglu_skinny(const void* X, const void* W_gate_up, void* act, int64_t M, int64_t inter, int64_t K,
                                       int64_t ldx, int64_t ldw, int64_t ldact, cudaStream_t stream) {
  using namespace b200;
  B200_CHECK_ARG(X && W_gate_up && act, "gemm_swiglu_skinny: null pointer");
  return skinny::gemm_swiglu_skinny(X, W_gate_up, act, M, inter, K, ldx, ldw, ldact, stream);
}
