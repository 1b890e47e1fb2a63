// The GEMM chain of one decode-step layer as ONE persistent kernel (M <= 64 token rows).
//
//   P0  acc_h  += attn @ W_o                                        o-proj (fused_transformer_layers.py:895-896), split-K
//   P1  residual += bf16(acc_h) ; ln = rmsnorm(residual) * w_ffn_ln ; acc_h = 0        compute_ffn_layernorm (:937-949)
//   P2  act = bf16(silu(g) * u), g|u = bf16(ln @ W_ffn1)            ffn1 + fused_bias_act("swiglu") (:100-168)
//   P3  acc_h  += act @ W_ffn2                                      ffn2, split-K
//   P4  residual += bf16(acc_h) ; ln = rmsnorm(residual) * w_next_ln ; acc_h = 0       compute_bias_residual_layernorm (:976-999)
//   P5  acc_qkv += ln @ W_qkv(next layer)^T                         compute_qkv of the NEXT layer (:843-856), split-K
//
// Why.  As six launches chained by programmatic dependent launch the step already prefetches each GEMM's weights while its
// predecessor drains, but every boundary still idles HBM for 4-8 us (CUPTI timeline, profiles/r02_decode_trace_timeline_b.json:
// kernel ends 7.1 / 7.9 / 38.0 / 17.4 / 4.6 / 7.9 us apart for 4.8 / 0 / 33.3 / 16.7 / 0 / 7.1 us of weight streaming at 7 TB/s): a
// dependent grid may only proceed when the WHOLE previous grid has retired and its memory is flushed, and the small norm kernels sit
// on the critical path with a launch latency of their own.  Here the six steps are PHASES of one grid of 2 CTAs per SM:
//   * a phase boundary is a grid-wide counter in global memory (release / acquire), not a kernel boundary;
//   * the TMA producer of every CTA runs AHEAD across phases: the weights do not depend on the activations, so while the CTA's
//     epilogue warps finish phase p and the grid barrier is pending, the ring already fills with the first weight tiles of phase
//     p+1 (and further tiles are requested into L2); only the small activation tiles wait for the barrier;
//   * the two norm phases run on the epilogue warps of the first M CTAs (one row each) between two barriers.
// GEMM phases are the swapped-operand tiles of gemm_skinny.cu (weights = 128-row M operand, tokens = N operand, fp32 partial tiles
// leave through TMA reduce-add, ffn1 tiles pair 64 gate columns with their 64 up columns and leave as one TMA store), same rounding
// points as the unfused path.
//
// Co-residency: a CTA waits for the whole grid at every barrier, so all CTAs must be resident: grid = 2 x SMs, 2 x (SMEM_BYTES +
// 1 KB) <= 228 KB (static_assert), 192 threads.  Every wait is bounded (trap instead of hang).
#include "../../include/b200nlp.h"
#include <cstring>

#include "common.cuh"
#include "host_util.h"
#include "norm_row.cuh"

namespace b200 {
namespace chain {

constexpr int BF = 128;    // output features per work item (UMMA M)
constexpr int BK = 64;     // k per stage
constexpr int UK = 16;
constexpr int NT = 64;     // token columns (UMMA N)
constexpr int NUM_THREADS = 192;
constexpr int W_BYTES = BF * BK * 2;          // 16 KB
constexpr int X_BYTES = NT * BK * 2;          // 8 KB
constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
constexpr int STAGES = 4;
constexpr int RING_BYTES = STAGES * STAGE_BYTES;       // 96 KB
constexpr int EPI_BYTES = 4 * 4096;                    // one 32x32 fp32 staging buffer per epilogue warp (= 16 KB for the SwiGLU tile)
constexpr int SMEM_BYTES = RING_BYTES + EPI_BYTES + 256;    // no alignment slack: the dynamic window is declared 1024-aligned
static_assert(2 * (SMEM_BYTES + 1024) <= 233472, "two CTAs per SM");
constexpr int MAX_PHASES = 6;

struct Phase {
  int kind;              // 0 = GEMM (fp32 reduce-add), 1 = GEMM + SwiGLU, 2 = add + RMSNorm
  // GEMM
  int N, K, f_tiles, split_k, kb_per_split, w_kmajor, inter, items;
  int w_map, x_map, o_map;
  // norm: residual += bf16(x_f32) ; normed = rmsnorm(residual) * w ; x_f32 = 0
  float* x_f32;
  bf16* res;
  const bf16* w;
  bf16* normed;          // nullptr: residual update only
  int h;
  float eps;
};

struct Params {
  int nph, M;
  Phase ph[MAX_PHASES];
  unsigned* counter;       // grid barriers: counter[ph * 32] += 1 per CTA when its share of phase ph is published (one 128-byte
                           // line per phase; a CTA without work in a phase publishes it at once, so the phases cannot share a count)
  unsigned* exit_counter;  // the last CTA to leave resets all of them
  int l2_kb;               // weight k-blocks requested into L2 beyond the ring while a barrier is pending
  unsigned long long* dbg; // optional [gridDim.x][MAX_PHASES][2] globaltimer stamps: inputs seen ready (producer) / phase published
};

struct Maps {
  CUtensorMap w[4];        // W_o, W_ffn1, W_ffn2, W_qkv
  CUtensorMap x[3];        // attn, ln, act   (bf16 [M, K], box 64 x 64)
  CUtensorMap o[3];        // acc_h (fp32 32x32), act (bf16 linear 64x64), acc_qkv (fp32 32x32)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* ptr, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// wait until `target` CTAs have published a phase; bounded
__device__ __forceinline__ void grid_wait(const unsigned* counter, unsigned target) {
  if (ld_acquire_u32(counter) >= target) return;
  const long long t0 = clock64();
  while (ld_acquire_u32(counter) < target) {
    __nanosleep(40);
    if (clock64() - t0 > 8000000000LL) {
      printf("[b200 watchdog] decode chain: grid barrier timed out: block %d target %u have %u\n", (int)blockIdx.x, target,
             ld_acquire_u32(counter));
      __trap();
    }
  }
}

struct Item { int f_tile, f0, kb0, nkb; };
__device__ __forceinline__ Item get_item(const Phase& P, int item) {
  Item it;
  int split;
  if (P.w_kmajor) { it.f_tile = item / P.split_k; split = item - it.f_tile * P.split_k; }
  else { split = item / P.f_tiles; it.f_tile = item - split * P.f_tiles; }      // split-major for [K, N] weights
  it.f0 = it.f_tile * BF;
  const int num_kb = (P.K + BK - 1) / BK;
  it.kb0 = split * P.kb_per_split;
  it.nkb = min(num_kb, it.kb0 + P.kb_per_split) - it.kb0;
  return it;
}

__global__ void __launch_bounds__(NUM_THREADS, 2)
decode_chain_kernel(const __grid_constant__ Maps maps, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];       // 128-byte-swizzled tiles need 1024-byte aligned bases
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* epi = smem + RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + EPI_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;    // [1]
  uint64_t* acc_empty = acc_full + 1;        // [1]  4 arrivals (one per epilogue warp)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 1);
  float* s_part = reinterpret_cast<float*>(tmem_ptr_smem + 2);      // [4] norm partial sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned G = gridDim.x;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) tma_prefetch_desc(&maps.w[i]);
    for (int i = 0; i < 3; ++i) { tma_prefetch_desc(&maps.x[i]); tma_prefetch_desc(&maps.o[i]); }
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
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
      int stage = 0;
      uint32_t par = 0;
      bool waited_prev = false;                // griddepcontrol.wait done (nothing but weights may be touched before)
      for (int ph = 0; ph < p.nph; ++ph) {
        const Phase& P = p.ph[ph];
        if (P.kind == 2) continue;
        const CUtensorMap* tmW = &maps.w[P.w_map];
        const CUtensorMap* tmX = &maps.x[P.x_map];
        bool ready = false;                    // this phase's activations are published
        int pend_stage[STAGES], pend_k0[STAGES], npend = 0;
        auto load_w = [&](int s, const Item& it, int kb) {
          uint8_t* sw = smem + s * STAGE_BYTES;
          const int k0 = kb * BK;
          if (P.w_kmajor) {
            tma_load_2d(tmW, &full_bar[s], sw, k0, it.f0);                       // [128 features x 64 k], k contiguous
          } else if (P.kind == 1) {
            tma_load_2d(tmW, &full_bar[s], sw, it.f_tile * 64, k0);              // 64 gate columns
            tma_load_2d(tmW, &full_bar[s], sw + 64 * BK * 2, P.inter + it.f_tile * 64, k0);   // their up columns
          } else {
            tma_load_2d(tmW, &full_bar[s], sw, it.f0, k0);                       // two [64 k x 64 features] boxes
            tma_load_2d(tmW, &full_bar[s], sw + 64 * BK * 2, it.f0 + 64, k0);
          }
        };
        auto l2_w = [&](const Item& it, int kb) {
          const int k0 = kb * BK;
          if (P.w_kmajor) {
            tma_prefetch_l2_2d(tmW, k0, it.f0);
          } else if (P.kind == 1) {
            tma_prefetch_l2_2d(tmW, it.f_tile * 64, k0);
            tma_prefetch_l2_2d(tmW, P.inter + it.f_tile * 64, k0);
          } else {
            tma_prefetch_l2_2d(tmW, it.f0, k0);
            tma_prefetch_l2_2d(tmW, it.f0 + 64, k0);
          }
        };
        // the inputs of phase ph exist once every CTA has published phases 0 .. ph-1 (phase 0: once the previous kernel is done)
        auto become_ready = [&](const Item& it, int next_i) {
          for (int i = next_i; i < min(it.nkb, next_i + p.l2_kb); ++i) l2_w(it, it.kb0 + i);   // keep HBM busy while waiting
          // (the counter itself belongs to the previous launch of this kernel until the predecessor chain has retired)
          if (!waited_prev) { pdl_wait(); waited_prev = true; }
          if (ph > 0) grid_wait(p.counter + (ph - 1) * 32, G);
          if (p.dbg) p.dbg[(blockIdx.x * MAX_PHASES + ph) * 2] = globaltimer();
          fence_proxy_async_all();             // generic-proxy writes of other CTAs (ln) -> our async-proxy (TMA) reads
          for (int q = 0; q < npend; ++q)
            tma_load_2d(tmX, &full_bar[pend_stage[q]], smem + pend_stage[q] * STAGE_BYTES + W_BYTES, pend_k0[q], 0);
          npend = 0;
          ready = true;
        };
        Item last = {};
        int last_i = 0;
        for (int item = blockIdx.x; item < P.items; item += G) {
          const Item it = get_item(P, item);
          for (int i = 0; i < it.nkb; ++i) {
            if (!ready && npend == STAGES) become_ready(it, i);
            mbar_wait(&empty_bar[stage], par ^ 1u);
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            load_w(stage, it, it.kb0 + i);
            if (ready) {
              tma_load_2d(tmX, &full_bar[stage], smem + stage * STAGE_BYTES + W_BYTES, (it.kb0 + i) * BK, 0);
            } else {
              pend_stage[npend] = stage; pend_k0[npend] = (it.kb0 + i) * BK; ++npend;
            }
            if (++stage == STAGES) { stage = 0; par ^= 1u; }
          }
          last = it; last_i = it.nkb;
        }
        if (!ready && npend > 0) become_ready(last, last_i);
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      constexpr uint32_t idesc_mn = umma_idesc_bf16(BF, NT, true, false);    // W [K, N]: MN-major A
      constexpr uint32_t idesc_k = umma_idesc_bf16(BF, NT, false, false);    // W [N, K]: K-major A
      int stage = 0;
      uint32_t par = 0, acc_n = 0;
      for (int ph = 0; ph < p.nph; ++ph) {
        const Phase& P = p.ph[ph];
        if (P.kind == 2) continue;
        const uint32_t idesc = P.w_kmajor ? idesc_k : idesc_mn;
        const uint32_t w_lbo = P.w_kmajor ? 16 : 64 * BK * 2, w_adv = P.w_kmajor ? UK * 2 : UK * 128;
        for (int item = blockIdx.x; item < P.items; item += G) {
          const Item it = get_item(P, item);
          if (acc_n > 0) {
            mbar_wait(acc_empty, (acc_n - 1) & 1u);       // the epilogue has read the previous item's accumulator
            tc_fence_after();
          }
          for (int i = 0; i < it.nkb; ++i) {
            mbar_wait(&full_bar[stage], par);
            tc_fence_after();
            const uint32_t sw = smem_u32(smem + stage * STAGE_BYTES);
            const uint32_t sx = sw + W_BYTES;
#pragma unroll
            for (int k = 0; k < BK / UK; ++k)
              umma_ss<1>(tmem_base, umma_desc_sw128(sw + k * w_adv, w_lbo, 1024), umma_desc_sw128(sx + k * UK * 2, 16, 1024), idesc,
                         (i > 0 || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (++stage == STAGES) { stage = 0; par ^= 1u; }
          }
          umma_commit(acc_full);
          ++acc_n;
        }
      }
    }
  } else {
    // ===================================== epilogue / norm warps =====================================
    const int q = warp & 3;                       // TMEM lane quadrant
    const int et = threadIdx.x - 64;              // 0..127
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    uint8_t* my_buf = epi + q * 4096;
    uint32_t acc_n = 0;
    // the fp32 workspaces AND the barrier counter (reset by the last CTA of the previous launch of this kernel) belong to the
    // predecessor kernels until now: nothing is published before
    pdl_wait();
    for (int ph = 0; ph < p.nph; ++ph) {
      const Phase& P = p.ph[ph];
      if (P.kind != 2) {
        for (int item = blockIdx.x; item < P.items; item += G) {
          const Item it = get_item(P, item);
          mbar_wait(acc_full, acc_n & 1u);
          tc_fence_after();
          if (P.kind == 1) {
            // lanes [0,64) = gate channels 64 f_tile .. +63, lanes [64,128) = the matching up channels; columns = tokens
            uint32_t* s_x = reinterpret_cast<uint32_t*>(epi);            // [32 token pairs][64 channels] bf16x2 (8 KB)
            bf16* s_m = reinterpret_cast<bf16*>(epi + 8192);             // [64 tokens][64 channels] bf16 = the TMA store box
            uint32_t pk[32];
            {
              uint32_t v[2][32];
              tmem_ld32(taddr, v[0]);
              tmem_ld32(taddr + 32, v[1]);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i)
                pk[i] = pack_bf16x2(__uint_as_float(v[i >> 4][(2 * i) & 31]), __uint_as_float(v[i >> 4][(2 * i + 1) & 31]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty);
            const bool is_gate = q < 2;
            const int chl = (q & 1) * 32 + lane;
#pragma unroll
            for (int i = 0; i < 16; ++i) s_x[(is_gate ? 16 + i : i) * 64 + chl] = pk[is_gate ? 16 + i : i];
            named_bar_sync(1, 128);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int tp = is_gate ? i : 16 + i;
              const uint32_t other = s_x[tp * 64 + chl];
              const uint32_t m2 = is_gate ? swiglu_fwd_pair(pk[tp], other) : swiglu_fwd_pair(other, pk[tp]);
              reinterpret_cast<unsigned short*>(s_m)[(2 * tp) * 64 + chl] = static_cast<unsigned short>(m2 & 0xffffu);
              reinterpret_cast<unsigned short*>(s_m)[(2 * tp + 1) * 64 + chl] = static_cast<unsigned short>(m2 >> 16);
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (et == 0) {
              tma_store_2d(&maps.o[P.o_map], s_m, it.f_tile * 64, 0);
              tma_store_commit();
              tma_store_wait<0>();                // also frees s_x / s_m for the next item
            }
            named_bar_sync(1, 128);
          } else {
            const CUtensorMap* tmF = &maps.o[P.o_map];
#pragma unroll
            for (int ch = 0; ch < NT / 32; ++ch) {
              uint32_t v[32];
              const bool live = (it.f0 + q * 32 < P.N) && (ch * 32 < p.M);     // warp-uniform
              if (live) {
                tmem_ld32(taddr + ch * 32, v);
                tmem_ld_wait();
              }
              if (ch == NT / 32 - 1) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(acc_empty);
              }
              if (!live) continue;
              if (lane == 0) tma_store_wait_read<0>();     // the previous box of this warp has been read out of the staging buffer
              __syncwarp();
              // transpose through smem: staging row = token, 32 features (128 B) per row, 128B-swizzled like the fp32 tensor map
              const uint32_t base = smem_u32(my_buf) + (lane & 3) * 4;
#pragma unroll
              for (int t = 0; t < 32; ++t) {
                const uint32_t addr = base + t * 128 + ((((lane >> 2) ^ (t & 7))) << 4);
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v[t]) : "memory");
              }
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_reduce_add_2d(tmF, my_buf, it.f0 + q * 32, ch * 32);
                tma_store_commit();
              }
            }
          }
          ++acc_n;
        }
        if (lane == 0) tma_store_wait<0>();        // this warp's reduce-adds / stores of the phase are complete
        __threadfence();
      } else {
        // ---- residual += bf16(x_f32) ; normed = rmsnorm(residual) * w ; x_f32 = 0 : one row per CTA, 128 threads ----
        if (static_cast<int>(blockIdx.x) < p.M) {
          if (et == 0) grid_wait(p.counter + (ph - 1) * 32, G);
          named_bar_sync(1, 128);
          add_rmsnorm_row_128(P.x_f32, P.res, P.res, P.w, P.normed, static_cast<int>(blockIdx.x), P.h, P.eps, et, s_part, 1);
          fence_proxy_async_all();                 // ln / the zeroed workspace are read / added to by other CTAs' TMA
          __threadfence();
        }
      }
      // publish this CTA's share of phase ph
      named_bar_sync(1, 128);
      if (et == 0) {
        red_release_add_u32(p.counter + ph * 32, 1u);
        if (p.dbg) p.dbg[(blockIdx.x * MAX_PHASES + ph) * 2 + 1] = globaltimer();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, NT);
  }
  if (threadIdx.x == 0) {
    // the last CTA out hands the counters back zeroed (every CTA has passed every barrier by then)
    __threadfence();
    const unsigned old = atomicAdd(p.exit_counter, 1u);
    if (old == G - 1) {
      for (int ph = 0; ph < MAX_PHASES; ++ph) p.counter[ph * 32] = 0u;
      *p.exit_counter = 0u;
      __threadfence();
    }
  }
}

}  // namespace chain
}  // namespace b200

static unsigned long long* g_chain_dbg = nullptr;
/* debugging aid of tools/decode_probe.py: per-CTA, per-phase globaltimer stamps of the next launches (NULL = off) */
extern "C" int b200_decode_layer_chain_debug(void* stamps) {
  g_chain_dbg = static_cast<unsigned long long*>(stamps);
  return 0;
}

extern "C" int64_t b200_decode_layer_chain_workspace_bytes(void) { return (b200::chain::MAX_PHASES + 1) * 128; }

extern "C" int b200_decode_layer_chain(const void* attn, const void* w_o, const void* w_ffn_ln, const void* w_ffn1,
                                       const void* w_ffn2, const void* w_next_ln, const void* w_next_qkv, void* residual,
                                       void* ln_buf, void* act_buf, float* acc_h, float* acc_qkv, void* sync_ws, int64_t M, int64_t h,
                                       int64_t attn_width, int64_t inter, int64_t qkv_n, float eps, cudaStream_t stream) {
  using namespace b200;
  using namespace b200::chain;
  B200_CHECK_ARG(attn && w_o && w_ffn_ln && w_ffn1 && w_ffn2 && residual && ln_buf && act_buf && acc_h && sync_ws,
                 "decode_layer_chain: null pointer");
  B200_CHECK_ARG(M > 0 && M <= 64 && h % 128 == 0 && h <= 8192 && attn_width % 64 == 0 && inter % 64 == 0,
                 "decode_layer_chain: need 0 < M <= 64, h %% 128 == 0, h <= 8192, inter %% 64 == 0");
  const bool has_next = w_next_qkv != nullptr;
  B200_CHECK_ARG(!has_next || (w_next_ln && acc_qkv && qkv_n % 128 == 0), "decode_layer_chain: next-layer arguments");
  Maps maps;
  memset(&maps, 0, sizeof(maps));
  int rc;
  auto w_kn = [&](CUtensorMap* tm, const void* W, int64_t K, int64_t N) {          // W [K, N], boxes of 64 features x 64 k
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(K)}, strides[1] = {static_cast<uint64_t>(N) * 2};
    uint32_t box[2] = {64, BK};
    return encode_tmap_bf16(tm, W, 2, dims, strides, box);
  };
  auto x_map = [&](CUtensorMap* tm, const void* X, int64_t K) {                    // X [M, K]
    uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)}, strides[1] = {static_cast<uint64_t>(K) * 2};
    uint32_t box[2] = {BK, NT};
    return encode_tmap_bf16(tm, X, 2, dims, strides, box);
  };
  auto f_map = [&](CUtensorMap* tm, float* F, int64_t N) {                         // fp32 [M, N], 32 x 32 boxes
    uint64_t dims[2] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M)}, strides[1] = {static_cast<uint64_t>(N) * 4};
    uint32_t box[2] = {32, 32};
    return encode_tmap_f32(tm, F, 2, dims, strides, box);
  };
  if ((rc = w_kn(&maps.w[0], w_o, attn_width, h)) != 0) return rc;
  if ((rc = w_kn(&maps.w[1], w_ffn1, h, 2 * inter)) != 0) return rc;
  if ((rc = w_kn(&maps.w[2], w_ffn2, inter, h)) != 0) return rc;
  if (has_next) {                                                                   // W_qkv [qkv_n, h]: k contiguous
    uint64_t dims[2] = {static_cast<uint64_t>(h), static_cast<uint64_t>(qkv_n)}, strides[1] = {static_cast<uint64_t>(h) * 2};
    uint32_t box[2] = {BK, BF};
    if ((rc = encode_tmap_bf16(&maps.w[3], w_next_qkv, 2, dims, strides, box)) != 0) return rc;
  } else {
    maps.w[3] = maps.w[0];
  }
  if ((rc = x_map(&maps.x[0], attn, attn_width)) != 0) return rc;
  if ((rc = x_map(&maps.x[1], ln_buf, h)) != 0) return rc;
  if ((rc = x_map(&maps.x[2], act_buf, inter)) != 0) return rc;
  if ((rc = f_map(&maps.o[0], acc_h, h)) != 0) return rc;
  {
    uint64_t dims[2] = {static_cast<uint64_t>(inter), static_cast<uint64_t>(M)}, strides[1] = {static_cast<uint64_t>(inter) * 2};
    uint32_t box[2] = {64, 64};
    if ((rc = encode_tmap_bf16_linear(&maps.o[1], act_buf, 2, dims, strides, box)) != 0) return rc;
  }
  if (has_next) { if ((rc = f_map(&maps.o[2], acc_qkv, qkv_n)) != 0) return rc; } else { maps.o[2] = maps.o[0]; }

  const int slots = 2 * sm_count();
  Params p = {};
  p.M = static_cast<int>(M);
  auto gemm = [&](int kind, int64_t N, int64_t K, int w_kmajor, int w_map, int xm, int om, int64_t inter_) {
    Phase P = {};
    P.kind = kind; P.N = static_cast<int>(N); P.K = static_cast<int>(K); P.w_kmajor = w_kmajor; P.inter = static_cast<int>(inter_);
    P.w_map = w_map; P.x_map = xm; P.o_map = om;
    P.f_tiles = static_cast<int>((N + BF - 1) / BF);
    const int num_kb = static_cast<int>((K + BK - 1) / BK);
    int split_k = 1;
    if (kind == 0) {                              // as many items as fit one wave of two CTAs per SM, each keeping >= 4 k-blocks
      split_k = slots / P.f_tiles;
      if (split_k > num_kb / 4) split_k = num_kb / 4;
      if (split_k < 1) split_k = 1;
    }
    P.kb_per_split = (num_kb + split_k - 1) / split_k;
    P.split_k = (num_kb + P.kb_per_split - 1) / P.kb_per_split;
    P.items = P.f_tiles * P.split_k;
    return P;
  };
  auto norm = [&](const void* w, bool want_normed) {
    Phase P = {};
    P.kind = 2; P.x_f32 = acc_h; P.res = static_cast<bf16*>(residual); P.w = static_cast<const bf16*>(w);
    P.normed = want_normed ? static_cast<bf16*>(ln_buf) : nullptr; P.h = static_cast<int>(h); P.eps = eps;
    return P;
  };
  int n = 0;
  p.ph[n++] = gemm(0, h, attn_width, 0, 0, 0, 0, 0);                 // P0 o-proj
  p.ph[n++] = norm(w_ffn_ln, true);                                  // P1
  p.ph[n++] = gemm(1, 2 * inter, h, 0, 1, 1, 1, inter);              // P2 ffn1 + SwiGLU: tile = 64 gate + 64 up columns
  p.ph[n - 1].f_tiles = static_cast<int>(inter / 64);
  p.ph[n - 1].items = p.ph[n - 1].f_tiles;
  p.ph[n++] = gemm(0, h, inter, 0, 2, 2, 0, 0);                      // P3 ffn2
  p.ph[n++] = norm(has_next ? w_next_ln : w_ffn_ln, has_next);       // P4
  if (has_next) p.ph[n++] = gemm(0, qkv_n, h, 1, 3, 1, 2, 0);        // P5 next layer's qkv
  p.nph = n;
  p.counter = static_cast<unsigned*>(sync_ws);
  p.exit_counter = p.counter + MAX_PHASES * 32;                      // its own 128-byte line
  p.l2_kb = static_cast<int>((static_cast<long long>(l2_prefetch_mb()) << 20) / (static_cast<long long>(W_BYTES) * slots));
  p.dbg = g_chain_dbg;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(decode_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_last_error("decode_layer_chain smem attr: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  cudaError_t e = launch_pdl(decode_chain_kernel, dim3(static_cast<unsigned>(slots)), dim3(NUM_THREADS), SMEM_BYTES, stream, maps, p);
  if (e != cudaSuccess) {
    set_last_error("decode_layer_chain launch: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return check_launch("decode_layer_chain");
}
