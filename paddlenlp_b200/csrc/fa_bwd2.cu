// Causal GQA flash-attention backward, second generation (head_dim 128): transposed tiles, software-pipelined.
//
//   P = exp(S*scale - lse) ; dP = dO V^T ; dS = P o (dP - D) * scale, D = rowsum(dO o O)
//   dV = P^T dO ; dK = dS^T Q ; dQ = dS K                                  (same contract as fa_bwd.cu)
//
// Why a second kernel.  fa_bwd.cu runs each (kv tile, q tile) step as a serial chain — 2 MMAs, element-wise, 3 MMAs, dQ read-out —
// because S, dP|dQ, dV, dK fill all 512 TMEM columns and nothing can be double-buffered (ncu: tensor pipe 27.5 %, 20 % of
// the warp samples waiting for dQ).  Here the score tiles are computed TRANSPOSED, kv on the UMMA M dimension:
//     S^T = K Q^T ,  dP^T = V dO^T                       M = 128 kv rows, N = 64 q rows
// which (a) lets a step cover only 64 q rows at full tensor-core efficiency (N = 64 halves the cycles, M stays 128), so that
// S^T can be double-buffered inside 512 columns, (b) puts P^T and dS^T where the next MMAs need them: P^T stays in tensor
// memory as the A operand of dV += P^T dO (tcgen05.mma with A in TMEM), dS^T goes to shared memory once and serves dK += dS^T Q
// (K-major A) and dQ^T = K^T dS^T (MN-major B), and (c) gives every compute thread one kv ROW (TMEM lane): no row exchange.
//   TMEM (512 columns): S^T[0] [0,64)  S^T[1] [64,128)  dP^T [128,192)  dQ^T [192,256)  dV [256,384)  dK [384,512)
//   (P^T(n) overwrites the first 32 columns of S^T[n&1] as packed bf16)
//   warp 0      TMA producer: K, V once; Q_n, dO_n (64 rows each) + the 64 row statistics through a 3-stage ring
//   warp 1      MMA issuer, per step n:  S^T(n+1) | dP^T(n+1) | dV += P^T(n) dO(n) | dK += dS^T(n) Q(n) | dQ^T(n) = K^T dS^T(n)
//   warps 2..5  compute group 0 (even n), warps 6..9 compute group 1 (odd n): P^T / dS^T of step n, then the dQ^T(n) read-out
//   While group b exponentiates step n the tensor pipe runs S^T / dP^T of step n+1 and the three accumulating MMAs of step
//   n-1; dP^T and dQ^T are single-buffered and handed over by mbarriers (dp_free after the loads, dq_empty after the read-out).
// dQ^T tiles (d on lanes) are reduce-added by TMA into a TRANSPOSED fp32 buffer [B, nh, 128, Spad]; a finishing kernel
// transposes and rounds it to the caller's dq.  dK / dV partials of the GQA group are reduce-added like in fa_bwd.cu.
//
// Replaces Paddle-core flash_attn_grad (reference: fusion_ops.py:240-246 backward; csrc/gpu/flash_attn_bwd.cc:22-92).
#include "../../include/b200nlp.h"
#include "common.cuh"
#include <type_traits>

#include "host_util.h"

namespace b200 {
namespace fab2 {

constexpr int KV_TILE_BYTES = 128 * 128 * 2;   // 32 KB, two 64-column halves of 16 KB
constexpr int KV_HALF = KV_TILE_BYTES / 2;
constexpr int Q_TILE_BYTES = 64 * 128 * 2;     // 16 KB, two halves of 8 KB  (64 q rows)
constexpr int Q_HALF = Q_TILE_BYTES / 2;
constexpr int DS_BYTES = 128 * 64 * 2;         // 16 KB: dS^T [128 kv][64 q], one 128-byte swizzle row per kv row
constexpr int STAT_BYTES = 64 * 8;             // 64 x (-lse*log2e, -delta*scale)
constexpr int QST = 3;                         // Q / dO / stats ring stages
constexpr int STAGE_BYTES = 8 * 4096;          // per-warp 32x32 fp32 staging for the TMA reduce-adds
constexpr int NUM_THREADS = 320;
constexpr int OFF_K = 0, OFF_V = KV_TILE_BYTES, OFF_Q = 2 * KV_TILE_BYTES, OFF_DO = OFF_Q + QST * Q_TILE_BYTES,
              OFF_DS = OFF_DO + QST * Q_TILE_BYTES, OFF_STAGE = OFF_DS + 2 * DS_BYTES, OFF_STAT = OFF_STAGE + STAGE_BYTES,
              OFF_BAR = OFF_STAT + QST * STAT_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;

struct Params {
  int S, Spad, B, nh, kvh;
  float scale, scale_log2;
  const float2* stats;   // [B, nh, Spad] x 2 floats: -lse*log2e ("nl") and -delta*scale ("nd"), stored per 64-row block as
                         // (nl(2c), nl(2c+1), nd(2c), nd(2c+1)) for c = 0..31; padding rows hold (-inf, 0)
  // FlashMask, causal lower-triangular form (see fa_fwd.cu): mask_start[b, c] = first query row that may NOT see key column c,
  // non-decreasing in c and > c; nullptr = plain causal.  Here a compute thread owns one kv ROW of the transposed tiles, i.e. one
  // key column c: its start row is a per-thread scalar, and the mask is one more comparison in the steps that need it.
  const int* mask_start;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// 1-D bulk copy global -> shared with mbarrier completion (size and addresses multiples of 16 bytes)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// (x0, x1) = (a0, a1) * s + (c0, c1)
__device__ __forceinline__ void fma2v(float& x0, float& x1, float a0, float a1, float s, float c0, float c1) {
  asm("{\n\t.reg .b64 ra, rs, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rs, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rs, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(x0), "=f"(x1)
      : "f"(a0), "f"(a1), "f"(s), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void mul2(float& x0, float& x1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(x0), "=f"(x1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// Two 32x32 fp32 chunks held in registers (this warp's 32 TMEM lanes) -> fp32 staging -> TMA reduce-add of two 32x32 boxes at
// coordinates (c0 + 32*i, c1, c2, c3).  One 4 KB staging buffer per warp.
__device__ __forceinline__ void reduce_regs(const uint32_t (&o)[2][32], uint8_t* buf, const CUtensorMap* tm, int lane, int c0, int c1,
                                            int c2, int c3) {
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
      tma_reduce_add_4d(tm, buf, c0 + 32 * i, c1, c2, c3);
      tma_store_commit();
    }
  }
}
// TMEM accumulator columns [32*ch0, 32*(ch0+2)) of this warp's lanes -> registers -> reduce_regs
__device__ __forceinline__ void reduce_out(uint32_t tsrc, uint8_t* buf, const CUtensorMap* tm, int lane, int ch0, int c0, int c1,
                                           int c2, int c3) {
  uint32_t o[2][32];
  tmem_ld32(tsrc + ch0 * 32, o[0]);
  tmem_ld32(tsrc + (ch0 + 1) * 32, o[1]);
  tmem_ld_wait();
  reduce_regs(o, buf, tm, lane, c0, c1, c2, c3);
}

// 10 warps = 3 on one SM sub-partition (16 K registers each): 16384 / (3 * 32) = 170 registers per thread is the hardware limit for this
// block shape (a 200-register build fails to launch), which is what __launch_bounds__(320, 1) makes ptxas target
__global__ void __launch_bounds__(NUM_THREADS, 1)
fa_bwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
               const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdK,
               const __grid_constant__ CUtensorMap tmdV, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem + OFF_K;
  uint8_t* sV = smem + OFF_V;
  uint8_t* sQ = smem + OFF_Q;          // [QST]
  uint8_t* sdO = smem + OFF_DO;        // [QST]
  uint8_t* sdS = smem + OFF_DS;        // [2]
  uint8_t* sStage = smem + OFF_STAGE;
  uint8_t* sStat = smem + OFF_STAT;    // [QST]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  // Barriers that are WAITED ON by the two compute groups in alternation exist once per group (index n & 1, phase (n >> 1) & 1):
  // a group waits for steps n, n+2, ... — consecutive waits of the same parity on a shared barrier would be satisfied by the
  // previous same-parity phase while the other group's step is still in flight.  pds_full is per group for the mirror reason:
  // arrivals of group g for step n+1 must not complete the phase the MMA warp is waiting on for step n.
  uint64_t* kv_full = bars;            // [1]
  uint64_t* qdo_full = bars + 1;       // [QST]
  uint64_t* qdo_empty = bars + 4;      // [QST]
  uint64_t* s_full = bars + 7;         // [2]   S^T(n) in buffer n&1 (and every earlier MMA) complete
  uint64_t* dp_full = bars + 9;        // [2]   dP^T(n) complete
  uint64_t* dp_free = bars + 11;       // [1]   compute(n) has loaded S^T(n), dP^T(n) into registers (128 arrivals; in step order)
  uint64_t* pds_full = bars + 12;      // [2]   P^T(n) in TMEM, dS^T(n) in smem                          (128 arrivals)
  uint64_t* dq_full = bars + 14;       // [2]   dQ^T(n) complete
  uint64_t* dq_empty = bars + 16;      // [1]   dQ^T(n) read out                                         (128 arrivals; in step order)
  uint64_t* acc_full = bars + 17;      // [1]   dK, dV final
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = static_cast<int>(blockIdx.x);     // kv tile; tile 0 has the most work and is scheduled first
  const int hq = blockIdx.y, batch = blockIdx.z;
  const int kv_head = hq / (p.nh / p.kvh);
  const int kv0 = jt * 128;
  // 64-row q sub-tiles kv0, kv0+64, ... (q >= kv0: causal).  With a document mask the steps past the end of the document(s) of
  // this kv tile are skipped: the last column has the largest start row (non-decreasing), no column is seen from that row on.
  int q_end = p.S, start_min = 0x7fffffff;
  if (p.mask_start != nullptr) {
    const int* ms = p.mask_start + static_cast<size_t>(batch) * p.S;
    q_end = min(p.S, __ldg(ms + min(kv0 + 127, p.S - 1)));
    start_min = __ldg(ms + kv0);                   // the first column's document ends first: steps below it need no mask
  }
  const int n_iter = max(1, (q_end - kv0 + 63) / 64);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1);
    for (int i = 0; i < QST; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&dp_full[i], 1);
      mbar_init(&pds_full[i], 128);
      mbar_init(&dq_full[i], 1);
    }
    mbar_init(dp_free, 128);
    mbar_init(dq_empty, 128);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<1>(tmem_ptr_smem, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const uint32_t tST = tmem_base, tdP = tmem_base + 128, tdQ = tmem_base + 192, tdV = tmem_base + 256, tdK = tmem_base + 384;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * KV_TILE_BYTES);
      tma_load_4d(&tmK, kv_full, sK, 0, kv_head, kv0, batch);
      tma_load_4d(&tmK, kv_full, sK + KV_HALF, 64, kv_head, kv0, batch);
      tma_load_4d(&tmV, kv_full, sV, 0, kv_head, kv0, batch);
      tma_load_4d(&tmV, kv_full, sV + KV_HALF, 64, kv_head, kv0, batch);
      const float2* stat_row = p.stats + (static_cast<size_t>(batch) * p.nh + hq) * p.Spad;
      // The 3-stage ring gives a Q / dO tile about one step (~1500 tensor cycles) between its slot becoming free and its first
      // MMA: enough for an L2 hit, not for a DRAM miss (ncu r02d: the compute groups spent 27 % of their time waiting for S^T).
      // So every tile is requested into L2 PF_AHEAD steps before its shared-memory load.
      constexpr int PF_AHEAD = 4;
      for (int n = 0; n < min(PF_AHEAD, n_iter); ++n) {
        tma_prefetch_l2_4d(&tmQ, 0, hq, kv0 + n * 64, batch);  tma_prefetch_l2_4d(&tmQ, 64, hq, kv0 + n * 64, batch);
        tma_prefetch_l2_4d(&tmdO, 0, hq, kv0 + n * 64, batch); tma_prefetch_l2_4d(&tmdO, 64, hq, kv0 + n * 64, batch);
      }
      for (int n = 0; n < n_iter; ++n) {
        const int st = n % QST;
        const int q0 = kv0 + n * 64;
        if (n + PF_AHEAD < n_iter) {
          const int qp = q0 + PF_AHEAD * 64;
          tma_prefetch_l2_4d(&tmQ, 0, hq, qp, batch);  tma_prefetch_l2_4d(&tmQ, 64, hq, qp, batch);
          tma_prefetch_l2_4d(&tmdO, 0, hq, qp, batch); tma_prefetch_l2_4d(&tmdO, 64, hq, qp, batch);
        }
        mbar_wait(&qdo_empty[st], static_cast<uint32_t>((n / QST) & 1) ^ 1u);
        mbar_arrive_expect_tx(&qdo_full[st], 2 * Q_TILE_BYTES + STAT_BYTES);
        tma_load_4d(&tmQ, &qdo_full[st], sQ + st * Q_TILE_BYTES, 0, hq, q0, batch);
        tma_load_4d(&tmQ, &qdo_full[st], sQ + st * Q_TILE_BYTES + Q_HALF, 64, hq, q0, batch);
        tma_load_4d(&tmdO, &qdo_full[st], sdO + st * Q_TILE_BYTES, 0, hq, q0, batch);
        tma_load_4d(&tmdO, &qdo_full[st], sdO + st * Q_TILE_BYTES + Q_HALF, 64, hq, q0, batch);
        bulk_load_1d(sStat + st * STAT_BYTES, stat_row + q0, STAT_BYTES, &qdo_full[st]);
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer -------------------------------
    // The whole warp runs the (warp-uniform) control flow, barrier waits and descriptor arithmetic; one elected lane issues the
    // tcgen05 instructions.  Inside `if (lane == 0) { ... }` the compiler treats every value as divergent and rebuilds each
    // descriptor in vector registers, then moves it to the uniform registers UTCHMMA reads (ELECT + R2UR.BROADCAST loops): ~19
    // SASS instructions and ~120 cycles per MMA against 32-64 tensor cycles per MMA (ncu r02e: the issuing thread was busy 77 %
    // of the time, the tensor pipe 36 %).  Descriptors are therefore built once per operand and ADVANCED by adding the k-step's
    // byte offset (>> 4) to the 64-bit value, all in convergent code.
    {
      const bool leader = elect_one();
      const uint32_t tbase = __shfl_sync(0xffffffffu, tmem_base, 0);      // warp-uniform by construction: lets it live in a UR
      const uint32_t uST = tbase, udP = tbase + 128, udQ = tbase + 192, udV = tbase + 256, udK = tbase + 384;
      constexpr uint32_t id_st = umma_idesc_bf16(128, 64, false, false);    // S^T / dP^T : A = K|V (K-major), B = Q|dO (K-major), N = 64
      constexpr uint32_t id_dv = umma_idesc_bf16(128, 128, false, true);    // dV : A = P^T (TMEM), B = dO (MN-major)
      constexpr uint32_t id_dk = umma_idesc_bf16(128, 128, false, true);    // dK : A = dS^T (smem, K-major), B = Q (MN-major)
      constexpr uint32_t id_dq = umma_idesc_bf16(128, 64, true, true);      // dQ^T: A = K^T (MN-major), B = dS^T (MN-major), N = 64
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aQ0 = smem_u32(sQ), adO0 = smem_u32(sdO), adS0 = smem_u32(sdS);
      const uint64_t dK_k = umma_desc_sw128(aK, 16, 1024), dV_k = umma_desc_sw128(aV, 16, 1024);   // K-major over head_dim
      const uint64_t dK_mn = umma_desc_sw128(aK, KV_HALF, 1024);                                      // K^T: MN-major over head_dim
      // byte offset of k-step kk (16 of the 128 head_dim columns) in a K-major tile stored as two 64-column halves
      auto koff = [](int kk, uint32_t half_bytes) { return static_cast<uint64_t>(((kk >> 2) * half_bytes + (kk & 3) * 32) >> 4); };
      auto issue_st_dp = [&](int n, bool first_wait) {
        const int st = n % QST;
        mbar_wait(&qdo_full[st], static_cast<uint32_t>((n / QST) & 1));
        tc_fence_after();
        const uint64_t dQ_k = umma_desc_sw128(aQ0 + st * Q_TILE_BYTES, 16, 1024);
        const uint64_t ddO_k = umma_desc_sw128(adO0 + st * Q_TILE_BYTES, 16, 1024);
        const uint32_t tS = uST + static_cast<uint32_t>((n & 1) * 64);
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss<1>(tS, dK_k + koff(kk, KV_HALF), dQ_k + koff(kk, Q_HALF), id_st, kk > 0);   // S^T = K Q^T
          umma_commit(&s_full[n & 1]);
        }
        if (first_wait) {
          mbar_wait(dp_free, static_cast<uint32_t>((n - 1) & 1));       // compute(n-1) holds dP^T(n-1) in registers
          tc_fence_after();
        }
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss<1>(udP, dV_k + koff(kk, KV_HALF), ddO_k + koff(kk, Q_HALF), id_st, kk > 0);  // dP^T = V dO^T
          umma_commit(&dp_full[n & 1]);
        }
      };
      mbar_wait(kv_full, 0);
      issue_st_dp(0, false);
      for (int n = 0; n < n_iter; ++n) {
        const int st = n % QST;
        if (n + 1 < n_iter) issue_st_dp(n + 1, true);
        mbar_wait(&pds_full[n & 1], static_cast<uint32_t>((n >> 1) & 1));
        tc_fence_after();
        const uint64_t dQ_mn = umma_desc_sw128(aQ0 + st * Q_TILE_BYTES, Q_HALF, 1024);      // Q / dO as MN-major B (k-step = 16 rows = 2 KB)
        const uint64_t ddO_mn = umma_desc_sw128(adO0 + st * Q_TILE_BYTES, Q_HALF, 1024);
        const uint64_t ddS_k = umma_desc_sw128(adS0 + (n & 1) * DS_BYTES, 16, 1024);         // dS^T K-major A (k-step = 32 bytes)
        const uint32_t tP = uST + static_cast<uint32_t>((n & 1) * 64);
        const uint32_t acc0 = n > 0 ? 1u : 0u;
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)     // dV += P^T dO : K = 64 q;  A k-step = 8 TMEM columns, B k-step = 16 dO rows
            umma_ts(udV, tP + kk * 8, ddO_mn + static_cast<uint64_t>(kk * 128), id_dv, kk > 0 ? 1u : acc0);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)     // dK += dS^T Q : A k-step = 32 bytes along the 128-byte q row, B k-step = 16 Q rows
            umma_ss<1>(udK, ddS_k + static_cast<uint64_t>(kk * 2), dQ_mn + static_cast<uint64_t>(kk * 128), id_dk, kk > 0 ? 1u : acc0);
          umma_commit(&qdo_empty[st]);       // Q / dO / stats of this step are not read again
        }
        if (n > 0) {
          mbar_wait(dq_empty, (n - 1) & 1);
          tc_fence_after();
        }
        if (leader) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)     // dQ^T = K^T dS^T : K = 128 kv;  A k-step = 16 K rows (2 KB), B k-step = 16 dS^T rows (2 KB)
            umma_ss<1>(udQ, dK_mn + static_cast<uint64_t>(kk * 128), ddS_k + static_cast<uint64_t>(kk * 128), id_dq, kk > 0);
          umma_commit(&dq_full[n & 1]);
        }
      }
      if (leader) umma_commit(acc_full);
      __syncwarp();
    }
  } else {
    // ------------------------------- compute groups: one thread per kv row -------------------------------
    const int g = (warp - 2) >> 2;                       // group 0: even steps, group 1: odd steps
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                      // kv row within the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    uint8_t* my_stage = sStage + (warp - 2) * 4096;
    // q rows (relative to kv0) this thread's key column is visible to: [r, vis_end)
    int vis_end = 0x7fffffff;
    if (p.mask_start != nullptr && kv0 + r < p.S)
      vis_end = __ldg(p.mask_start + static_cast<size_t>(batch) * p.S + kv0 + r) - kv0;
    auto read_out_dq = [&](int n) {                      // dQ^T(n): lanes = d, columns = 64 q rows of step n
      mbar_wait(&dq_full[n & 1], static_cast<uint32_t>((n >> 1) & 1));
      tc_fence_after();
      uint32_t o[2][32];
      tmem_ld32(tdQ + lane_off, o[0]);
      tmem_ld32(tdQ + lane_off + 32, o[1]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_empty);                             // the columns are free as soon as they sit in registers: dQ^T(n+1) may start
      reduce_regs(o, my_stage, &tmdQ, lane, kv0 + n * 64, quad * 32, hq, batch);
    };
    for (int n = g; n < n_iter; n += 2) {
      const int st = n % QST;
      const uint32_t tS = tST + lane_off + static_cast<uint32_t>(g * 64);
      mbar_wait(&qdo_full[st], static_cast<uint32_t>((n / QST) & 1));      // row statistics of this step are in smem
      mbar_wait(&s_full[g], static_cast<uint32_t>((n >> 1) & 1));
      mbar_wait(&dp_full[g], static_cast<uint32_t>((n >> 1) & 1));
      tc_fence_after();
      uint32_t sv[64], dv[64];
      {
        uint32_t(*a)[32] = reinterpret_cast<uint32_t(*)[32]>(sv);
        uint32_t(*b)[32] = reinterpret_cast<uint32_t(*)[32]>(dv);
        tmem_ld32(tS, a[0]); tmem_ld32(tS + 32, a[1]);
        tmem_ld32(tdP + lane_off, b[0]); tmem_ld32(tdP + lane_off + 32, b[1]);
        tmem_ld_wait();
      }
      tc_fence_before();
      mbar_arrive(dp_free);
      const uint32_t stat_a = smem_u32(sStat + st * STAT_BYTES);      // per q pair: (nl0, nl1, nd0, nd1)
      const int diag_shift = n * 64;                     // column c is visible to kv row r iff r <= c + 64 n
      uint32_t pk[32], dk[32];
      auto body = [&](auto diag_tag) {
        constexpr bool DIAG = decltype(diag_tag)::value;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const uint4 su = ld_shared_v4(stat_a + c * 16);  // warp-uniform address: one broadcast LDS.128
          float x0, x1, t0, t1, p0, p1, d0, d1;
          fma2v(x0, x1, __uint_as_float(sv[2 * c]), __uint_as_float(sv[2 * c + 1]), p.scale_log2, __uint_as_float(su.x),
                __uint_as_float(su.y));
          p0 = fast_exp2(x0); p1 = fast_exp2(x1);
          if constexpr (DIAG) {
            const int qr = 2 * c + diag_shift;               // q row of column 2c, relative to kv0
            if (r > qr || qr >= vis_end) p0 = 0.f;
            if (r > qr + 1 || qr + 1 >= vis_end) p1 = 0.f;
          }
          fma2v(t0, t1, __uint_as_float(dv[2 * c]), __uint_as_float(dv[2 * c + 1]), p.scale, __uint_as_float(su.z),
                __uint_as_float(su.w));
          mul2(d0, d1, p0, p1, t0, t1);
          pk[c] = pack_bf16x2(p0, p1);
          dk[c] = pack_bf16x2(d0, d1);
        }
      };
      // masked steps: the two that touch the diagonal, and (FlashMask) those that reach the end of this tile's first document
      if (diag_shift < 128 || kv0 + diag_shift + 63 >= start_min) body(std::true_type{}); else body(std::false_type{});
      {
        uint32_t(*a)[16] = reinterpret_cast<uint32_t(*)[16]>(pk);
        tmem_st16(tS, a[0]); tmem_st16(tS + 16, a[1]);   // P^T: 64 q as 32 packed columns over the start of S^T[g]
      }
      const uint32_t row_s = smem_u32(sdS + g * DS_BYTES) + r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        st_shared_v4(row_s + ((c ^ (r & 7)) << 4), make_uint4(dk[4 * c], dk[4 * c + 1], dk[4 * c + 2], dk[4 * c + 3]));
      tmem_st_wait();
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&pds_full[g]);
      // read out the OTHER group's previous step: its dQ^T MMA was issued ~800 tensor cycles after that group's compute ended and
      // has normally retired while this group was exponentiating, so nobody idles on dq_full
      if (n > 0) read_out_dq(n - 1);
    }
    if ((n_iter & 1) == g) read_out_dq(n_iter - 1);     // the last step's dQ^T: read by the group that would own step n_iter
    // every step's dQ^T was read out by exactly one group (128 arrivals on dq_empty each).  Epilogue: this head's dK / dV partials -> fp32 reduce-add (the GQA group's heads sum in L2)
    mbar_wait(acc_full, 0);
    tc_fence_after();
    reduce_out(tdK + lane_off, my_stage, &tmdK, lane, g * 2, g * 64, kv_head, kv0 + quad * 32, batch);
    reduce_out(tdV + lane_off, my_stage, &tmdV, lane, g * 2, g * 64, kv_head, kv0 + quad * 32, batch);
    if (lane == 0) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// stats[b, h, s] = (-lse * log2e, -scale * sum_d dO[b,s,h,d] * O[b,s,h,d]) for s < S, (-inf, 0) for S <= s < Spad   (16 lanes per row)
__global__ void fa_bwd2_stats_kernel(const bf16* __restrict__ o, const bf16* __restrict__ dout, const float* __restrict__ lse,
                                     float2* __restrict__ stats, int B, int S, int Spad, int nh, int64_t ldo, int64_t lddo,
                                     float scale) {
  const int64_t row = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 4;   // (b, s, h) flattened over Spad
  const int sub = threadIdx.x & 15;
  const int64_t total = static_cast<int64_t>(B) * Spad * nh;
  if (row >= total) return;
  const int h = static_cast<int>(row % nh);
  const int64_t tokp = row / nh;
  const int s = static_cast<int>(tokp % Spad), b = static_cast<int>(tokp / Spad);
  float acc = 0.f;
  if (s < S) {
    const int64_t tok = static_cast<int64_t>(b) * S + s;
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
  if (sub == 0) {
    float2 v;
    if (s < S) {
      v.x = -lse[(static_cast<size_t>(b) * nh + h) * S + s] * 1.4426950408889634f;
      v.y = -acc * scale;
    } else {
      v.x = -INFINITY;
      v.y = 0.f;
    }
    // layout per 64-row block: for the q pair (2c, 2c+1): nl(2c), nl(2c+1), nd(2c), nd(2c+1) — the pairs the f32x2 FMAs consume
    float* blk = reinterpret_cast<float*>(stats + (static_cast<size_t>(b) * nh + h) * Spad + (s & ~63));
    const int c = (s & 63) >> 1, e = s & 1;
    blk[c * 4 + e] = v.x;
    blk[c * 4 + 2 + e] = v.y;
  }
}

// dq[b, s, h, :] (bf16, token stride lddq) = bf16(accT[b, h, :, s])  — 64 q rows x 128 d per block, through shared memory
// (64 consecutive floats = 256 contiguous bytes per d row: with 32-row blocks the reads were 128-byte pieces 16 KB apart and the
// kernel ran at 2.7 TB/s)
__global__ void __launch_bounds__(256) fa_bwd2_dq_finish_kernel(const float* __restrict__ accT, bf16* __restrict__ dq, int S, int Spad,
                                                               int nh, int64_t lddq) {
  __shared__ float tile[128][65];
  const int s0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
  const float* src = accT + (static_cast<size_t>(b) * nh + h) * 128 * Spad;
  for (int i = threadIdx.x; i < 128 * 64; i += 256) {
    const int d = i >> 6, s = i & 63;
    tile[d][s] = src[static_cast<size_t>(d) * Spad + s0 + s];      // 64 consecutive floats per d row: coalesced
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 16; i += 256) {
    const int s = i >> 4, c = i & 15;                              // 16 chunks of 8 d per token
    if (s0 + s < S) {
      uint4 o;
      o.x = pack_bf16x2(tile[c * 8 + 0][s], tile[c * 8 + 1][s]);
      o.y = pack_bf16x2(tile[c * 8 + 2][s], tile[c * 8 + 3][s]);
      o.z = pack_bf16x2(tile[c * 8 + 4][s], tile[c * 8 + 5][s]);
      o.w = pack_bf16x2(tile[c * 8 + 6][s], tile[c * 8 + 7][s]);
      *(reinterpret_cast<uint4*>(dq + (static_cast<size_t>(b) * S + s0 + s) * lddq + h * 128) + c) = o;
    }
  }
}

// out (bf16, token stride ld) = bf16(acc fp32 [tokens, width])
__global__ void fa_bwd2_kv_finish_kernel(const float* __restrict__ acc, bf16* __restrict__ out, int64_t tokens, int width,
                                         int64_t ld) {
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
    *(reinterpret_cast<uint4*>(out + t * ld) + c) = o;
  }
}

static int make_map(CUtensorMap* tm, const void* base, int64_t B, int64_t S, int64_t heads, int64_t ld, uint32_t box_rows) {
  uint64_t dims[4] = {128, static_cast<uint64_t>(heads), static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {128 * 2, static_cast<uint64_t>(ld) * 2, static_cast<uint64_t>(S) * ld * 2};
  uint32_t box[4] = {64, 1, box_rows, 1};
  return encode_tmap_bf16(tm, base, 4, dims, strides, box);
}
// fp32 [B, S, heads, 128] contiguous, 32x32 boxes (dK / dV accumulation: lanes = kv rows, columns = d)
static int make_acc_map(CUtensorMap* tm, const void* base, int64_t B, int64_t S, int64_t heads) {
  uint64_t dims[4] = {128, static_cast<uint64_t>(heads), static_cast<uint64_t>(S), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {128 * 4, static_cast<uint64_t>(heads) * 128 * 4, static_cast<uint64_t>(S) * heads * 128 * 4};
  uint32_t box[4] = {32, 1, 32, 1};
  return encode_tmap_f32(tm, base, 4, dims, strides, box);
}
// fp32 TRANSPOSED dQ accumulation [B, heads, 128, Spad]: dims {Spad, 128, heads, B}, boxes of 32 q x 32 d (lanes = d)
static int make_dqT_map(CUtensorMap* tm, const void* base, int64_t B, int64_t Spad, int64_t heads) {
  uint64_t dims[4] = {static_cast<uint64_t>(Spad), 128, static_cast<uint64_t>(heads), static_cast<uint64_t>(B)};
  uint64_t strides[3] = {static_cast<uint64_t>(Spad) * 4, static_cast<uint64_t>(Spad) * 128 * 4,
                         static_cast<uint64_t>(Spad) * 128 * heads * 4};
  uint32_t box[4] = {32, 32, 1, 1};
  return encode_tmap_f32(tm, base, 4, dims, strides, box);
}

}  // namespace fab2

// Causal (optionally FlashMask start rows) backward through the transposed, pipelined kernel; called by b200_fa_bwd_flashmask
// (fa_bwd.cu).  Workspace layout: dQ^T accumulation [B, nh, 128, Spad] | dK acc | dV acc | stats [B, nh, Spad] float2.
int launch_fa_bwd2(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                   const int32_t* mask_start_rows, void* dq, void* dk, void* dv, void* workspace, int64_t B, int64_t S, int64_t num_heads, int64_t num_kv_heads,
                   int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo, int64_t lddq, int64_t lddk, int64_t lddv,
                   float softmax_scale, cudaStream_t stream) {
  using namespace fab2;
  const int64_t Spad = (S + 63) / 64 * 64;
  float* dq_acc = static_cast<float*>(workspace);
  float* dk_acc = dq_acc + B * num_heads * 128 * Spad;
  float* dv_acc = dk_acc + B * S * num_kv_heads * 128;
  float2* stats = reinterpret_cast<float2*>(dv_acc + B * S * num_kv_heads * 128);
  cudaError_t e = cudaMemsetAsync(dq_acc, 0, static_cast<size_t>(B) * (Spad * num_heads + 2 * S * num_kv_heads) * 128 * 4, stream);
  if (e != cudaSuccess) {
    set_last_error("fa_bwd2 memset: %s", cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  int rc;
  {
    const int64_t threads = B * Spad * num_heads * 16;
    fa_bwd2_stats_kernel<<<static_cast<unsigned>((threads + 255) / 256), 256, 0, stream>>>(
        static_cast<const bf16*>(o), static_cast<const bf16*>(dout), lse, stats, (int)B, (int)S, (int)Spad, (int)num_heads, ldo,
        lddo, softmax_scale);
    if ((rc = check_launch("fa_bwd2(stats)")) != 0) return rc;
  }
  CUtensorMap tmQ, tmK, tmV, tmdO, tmdQ, tmdK, tmdV;
  if ((rc = make_map(&tmQ, q, B, S, num_heads, ldq, 64)) != 0) return rc;
  if ((rc = make_map(&tmK, k, B, S, num_kv_heads, ldk, 128)) != 0) return rc;
  if ((rc = make_map(&tmV, v, B, S, num_kv_heads, ldv, 128)) != 0) return rc;
  if ((rc = make_map(&tmdO, dout, B, S, num_heads, lddo, 64)) != 0) return rc;
  if ((rc = make_dqT_map(&tmdQ, dq_acc, B, Spad, num_heads)) != 0) return rc;
  if ((rc = make_acc_map(&tmdK, dk_acc, B, S, num_kv_heads)) != 0) return rc;
  if ((rc = make_acc_map(&tmdV, dv_acc, B, S, num_kv_heads)) != 0) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(fa_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) {
      set_last_error("fa_bwd2 smem attr: %s", cudaGetErrorString(e));
      return static_cast<int>(e);
    }
    attr_set = true;
  }
  Params p;
  p.S = (int)S; p.Spad = (int)Spad; p.B = (int)B; p.nh = (int)num_heads; p.kvh = (int)num_kv_heads;
  p.scale = softmax_scale;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.stats = stats;
  p.mask_start = mask_start_rows;
  dim3 grid(static_cast<unsigned>((S + 127) / 128), static_cast<unsigned>(num_heads), static_cast<unsigned>(B));
  fa_bwd2_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(tmQ, tmK, tmV, tmdO, tmdQ, tmdK, tmdV, p);
  if ((rc = check_launch("fa_bwd2")) != 0) return rc;
  {
    dim3 g2(static_cast<unsigned>(Spad / 64), static_cast<unsigned>(num_heads), static_cast<unsigned>(B));
    fa_bwd2_dq_finish_kernel<<<g2, 256, 0, stream>>>(dq_acc, static_cast<bf16*>(dq), (int)S, (int)Spad, (int)num_heads, lddq);
    if ((rc = check_launch("fa_bwd2(dq finish)")) != 0) return rc;
    const int64_t tokens = B * S;
    const int kvw = static_cast<int>(num_kv_heads * 128);
    int64_t kblocks = (tokens * (kvw / 8) + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
    if (kblocks > cap) kblocks = cap;
    fa_bwd2_kv_finish_kernel<<<static_cast<unsigned>(kblocks), 256, 0, stream>>>(dk_acc, static_cast<bf16*>(dk), tokens, kvw, lddk);
    if ((rc = check_launch("fa_bwd2(dk finish)")) != 0) return rc;
    fa_bwd2_kv_finish_kernel<<<static_cast<unsigned>(kblocks), 256, 0, stream>>>(dv_acc, static_cast<bf16*>(dv), tokens, kvw, lddv);
    rc = check_launch("fa_bwd2(dv finish)");
  }
  return rc;
}

}  // namespace b200
