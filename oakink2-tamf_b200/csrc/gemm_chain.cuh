// gemm_chain.cuh -- two dependent GEMMs of an encoder layer in ONE persistent tcgen05 kernel.
//
//   phase 1   X = LayerNorm(X + A1 . W1^T + b1)        (out_proj + LN1  |  linear2 + LN2; eps 1e-5, biased variance)
//   phase 2   C2 = act(Xb . W2^T + b2) -> bf16          (linear1 + GELU  |  the NEXT layer's in_proj)
//
// Reference semantics: nn.TransformerEncoderLayer, post-norm (interaction_segment_mdm.py:63-70; segment_refine_model.py
// :88-95).  Why one kernel: M = 10 560 token rows are 41.25 row tiles of 256, so an N = 512 GEMM has 84 pair tiles for 74
// CTA pairs -- two rounds, the second 14 % full -- and a LayerNorm epilogue that needs the whole 512-wide row filled all
// of TMEM (no overlap with a mainloop) while the ~32 B/clk/SM store path of 84 SMs carried the whole result (round 1:
// 49 k cycles for 18 k of balanced tensor work, profiles/r01_gemm_cta_timelines.txt).  Here
//   * every unit of work is a 256 x 256 tile of a CTA pair (tcgen05.mma.cta_group::2, M = 256, N = 256), accumulators
//     double-buffered in TMEM, so the epilogue of unit i overlaps the mainloop of unit i + 1 for EVERY epilogue;
//   * the residual is added BY THE TENSOR CORE: after the K loop of a LayerNorm unit, eight more 64-deep k-blocks
//     multiply the unit's own 256 x 256 block of the two residual planes (Xb, Xlo) with a 64 x 64 identity that stays
//     in shared memory (M = 256, N = 64, K = 64 per 64-column block and plane: products with 1.0 are exact, the
//     accumulation is fp32), so the epilogue never loads the residual: no small-box TMA loads (one 64-byte row per ~2
//     cycles of the SM's TMA unit -- they delayed the mainloop's own loads), no staging traffic, + 1.6 k cycles of MMA;
//   * a 512-wide LayerNorm row is computed by TWO pairs (column halves): each keeps its 128 x 256 slab of
//     y = acc + b in REGISTERS (64 per thread), the halves swap per-row (sum, sum of squares) through L2 as ONE 64-bit
//     word per row that is its own flag (no fences: a gpu-scope fence costs ~1.5 k cycles), and all 148 SMs store;
//   * the units of phase 2 are scheduled into the gaps of phase 1 by a host-built static schedule (list scheduling on
//     per-unit cost estimates); a phase-2 tile of row tile m waits for a counter that the LayerNorm epilogues of
//     row tile m bump after their TMA stores have completed (release / acquire through L2, proxy fences on both
//     sides).  All CTAs of the grid are co-resident (grid <= SM count, 1 CTA / SM) and waits only point at units
//     that come EARLIER in some pair's list, so the schedule cannot deadlock; every spin is bounded (trap).
// The sync words of this kernel are zeroed by its sibling (the other chain kernel of the layer), which runs in between.
#pragma once
#include <algorithm>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

#include "gemm.cuh"

namespace tamf {

enum ChainEpi2 { CHAIN_BIAS = 1, CHAIN_GELU = 2 };

struct ChainParams {
  int M;       // token rows
  int N1, K1;  // phase 1: N1 = d (256 | 512) = LayerNorm width, halves = N1 / 256
  int N2, K2;  // phase 2: N2 % 256 == 0 (0: no phase 2)
  const float *bias1, *gamma, *beta;  // [N1]
  const float* bias2;                 // [N2]
  const int* sched_off;               // [pairs + 1]
  const int* sched;                   // unit codes: phase << 28 | row tile << 8 | column tile
  unsigned* ready;                    // [tiles_m]            LayerNorm rows of a row tile stored (counts warps)
  unsigned long long* stats;          // [tiles_m*halves*2*128] (sum | sum of squares << 32) over the CTA's 256 columns of a
                                      // row; all-ones = not posted yet (the word is its own flag)
  unsigned ready_target;              // halves * 2 CTAs * 16 warps
  unsigned* zero_ptr;                 // the sibling kernel's `ready` words, cleared here (after the dependency wait)
  int zero_n;
  unsigned long long* ones_ptr;       // the sibling kernel's `stats` words, reset to all-ones here
  int ones_n;
  long long* trace;                   // debug only: [grid][GEMM_TRACE_SLOTS] clock64 stamps
  int dbg;
};

// 20 warps = 5 warpgroups: 16 epilogue warps, then one warpgroup with the TMA producer (16), the MMA issuer (17) and two
// idle warps, so that register reallocation (setmaxnreg) always involves whole warpgroups.
constexpr int CH_THREADS = GEMM_EPI_THREADS + 128;
constexpr int CH_STAGES = 4;
constexpr int CH_BN = 256;
constexpr int CH_A_BYTES = GEMM_BM * 64 * 2;       // 128 rows x 64 k
constexpr int CH_B_BYTES = (CH_BN / 2) * 64 * 2;   // this CTA's half of the 256 W rows
constexpr int CH_STAGE_BYTES = CH_A_BYTES + CH_B_BYTES;
constexpr int CH_PIPE_BYTES = CH_STAGES * CH_STAGE_BYTES;
constexpr int CH_STG_BYTES = GEMM_EPI_WARPS * GEMM_STG_WARP;
constexpr int CH_IDENT_BYTES = 32 * 128;  // this CTA's 32 rows of the 64 x 64 bf16 identity (K-major, 128-byte swizzle)
constexpr int CH_RES_KB = 4;              // residual ring stages of a LayerNorm unit: 2 planes x 2 stages, each holding
                                          // TWO 64-column blocks (A slot, B slot) -- the ring is latency bound per stage
constexpr int CH_CTRL_BYTES = 1024;
// bias1 | gamma | beta | bias2 (256 floats each) | row statistics [2 buffers][sum, sq][4 column quarters][128 rows]
constexpr int CH_PARAM_BYTES = 4 * CH_BN * 4 + 2 * 2 * 4 * 128 * 4;
constexpr int CH_SMEM_BYTES = 1024 + CH_PIPE_BYTES + CH_STG_BYTES + CH_IDENT_BYTES + CH_CTRL_BYTES + CH_PARAM_BYTES;
static_assert(CH_SMEM_BYTES <= GEMM_SMEM_MAX, "shared memory budget (227 KB) exceeded");

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void red_relaxed_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic proxy <-> async proxy (TMA) ordering for every state space
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Bounded mbarrier wait without printf (a call site keeps many registers alive around it; the trap alone reports the
// protocol bug as a CUDA error).
__device__ __forceinline__ void mbar_wait_q(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// Lane 0 polls `*flag >= target` (acquire, gpu scope) with a bounded spin; the warp leaves together.
__device__ __forceinline__ void chain_wait_ge(const unsigned* flag, unsigned target) {
  if (lane_id() == 0) {
    if (ld_acquire_gpu(flag) < target) {
      const long long t0 = clock64();
      while (ld_acquire_gpu(flag) < target) {
        __nanosleep(40);
        if (clock64() - t0 > 4000000000LL) __trap();
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ long long chain_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return (long long)t;
}
#define CHAIN_TRACE_NS(slot)                                                                                     \
  do {                                                                                                           \
    if (p.trace) p.trace[(size_t)blockIdx.x * GEMM_TRACE_SLOTS + (slot)] = chain_globaltimer();                  \
  } while (0)
#define CHAIN_TRACE(slot)                                                                                        \
  do {                                                                                                           \
    if (p.trace && (slot) < GEMM_TRACE_SLOTS) p.trace[(size_t)blockIdx.x * GEMM_TRACE_SLOTS + (slot)] = clock64(); \
  } while (0)

template <int EPI2>
__global__ void __launch_bounds__(CH_THREADS, 1)
    gemm_chain_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                      const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                      const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmC2,
                      const __grid_constant__ CUtensorMap tmXh_st, const __grid_constant__ CUtensorMap tmXl_st,
                      const __grid_constant__ CUtensorMap tmI, const ChainParams p) {
  // tmXh: Xb [M,d], box {64, 128}: A operand of phase 2 AND the high residual plane of phase 1; tmXl: Xlo, same box.
  // tmXh_st / tmXl_st: the same planes with box {64, 32} (LayerNorm result stores); tmI: 64 x 64 identity, box {64, 32}.
  constexpr int STAGES = CH_STAGES, PW = GEMM_EPI_WARPS, PT = GEMM_EPI_THREADS, BN = CH_BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;  // identical in both CTAs of a pair
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * CH_A_BYTES;
  uint8_t* s_stage = smem + CH_PIPE_BYTES;
  uint8_t* s_ident = s_stage + CH_STG_BYTES;
  uint8_t* ctrl = s_ident + CH_IDENT_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);  // [STAGES] (the leader's copy is the live one)
  uint64_t* empty_bar = full_bar + STAGES;                  // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;                 // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                     // [2]      (the leader's copy is the live one)
  uint64_t* ident_bar = tempty_bar + 2;                     // [1]      identity tile landed (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ident_bar + 1);
  uint32_t* s_dep = tmem_slot + 1;  // units whose dependencies the scout warp has seen satisfied
  float* s_bias1 = reinterpret_cast<float*>(ctrl + CH_CTRL_BYTES);
  float* s_gamma = s_bias1 + BN;
  float* s_beta = s_gamma + BN;
  float* s_bias2 = s_beta + BN;
  float* s_stat = s_bias2 + BN;  // [2 buffers][sum, sq][4 column quarters][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader of the pair
  const int pair = blockIdx.x >> 1;
  const int halves = p.N1 / BN;
  const int u_begin = p.sched_off[pair], u_end = p.sched_off[pair + 1];  // host data (written at bind time)

  if (threadIdx.x == 0) CHAIN_TRACE(0);
  if (warp == PW && lane == 0) {
    tma_prefetch_desc(&tmA1);
    tma_prefetch_desc(&tmB1);
    tma_prefetch_desc(&tmXh);
    tma_prefetch_desc(&tmXl);
    tma_prefetch_desc(&tmXh_st);
    tma_prefetch_desc(&tmXl_st);
    tma_prefetch_desc(&tmI);
    if (p.N2 > 0) {
      tma_prefetch_desc(&tmB2);
      tma_prefetch_desc(&tmC2);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);  // the leader's arrive.expect_tx covers the bytes of BOTH CTAs' loads
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PW * 2);  // one elected lane per epilogue warp of both CTAs
    }
    mbar_init(ident_bar, 1);
    *s_dep = 0u;
    fence_mbar_init();
  }
  if (warp == PW + 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == PW) {  // the identity tile (constant data): this CTA's 32 rows, credited to the leader's barrier
    if (elect_one()) {
      if (rank == 0) mbar_arrive_expect_tx(ident_bar, 2 * CH_IDENT_BYTES);
      tma_load_2d_2sm(s_ident, &tmI, mapa_cluster(smem_u32(ident_bar), 0), 0, (int)rank * 32);
    }
    __syncwarp();
  }
  if (threadIdx.x == 0) CHAIN_TRACE(1);
  pdl_launch_dependents();
  pdl_wait();  // everything the previous kernels wrote (activations, our reset sync words) is visible from here on
  if (threadIdx.x == 0) {
    CHAIN_TRACE(2);
    CHAIN_TRACE_NS(56);  // globaltimer (ns) when the dependency wait ended: the common time base across SMs
  }

  // Registers: the block is allocated 20 warps x 96.  The producer / MMA warpgroup gives up 64 per thread (128 x 64 =
  // 8192 go back to the CTA's pool), the four epilogue warpgroups take 16 more each (512 x 16 = 8192): the LayerNorm
  // epilogue keeps 64 accumulator values per thread over two passes.
  // (each setmaxnreg dominates exactly one role's code, so ptxas allocates every role under its own limit)
  if (warp >= PW) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
   if (warp == PW) {
    // ===================== TMA producer (both CTAs of the pair) =====================
    uint32_t stage = 0, phase = 0;
    int it = 0;
    const uint32_t full_leader = mapa_cluster(smem_u32(&full_bar[0]), 0);
    for (int ui = u_begin; ui < u_end; ++ui, ++it) {
      const int code = p.sched[ui];
      const int ph = code >> 28, m = (code >> 8) & 0xFFFFF, n = code & 0xFF;
      const int m0 = m * 256 + (int)rank * GEMM_BM, n0 = n * BN;
      const CUtensorMap* ta = ph ? &tmXh : &tmA1;
      const CUtensorMap* tb = ph ? &tmB2 : &tmB1;
      const int num_kb = (ph ? p.K2 : p.K1) / 64;
      const int total_kb = num_kb + (ph ? 0 : CH_RES_KB);
      if (ph) {  // the LayerNorm rows of this row tile must have landed: the scout warp has seen the counter
        uint32_t seen;
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(s_dep)) : "memory");
        if (seen < (uint32_t)(it + 1)) {
          const long long t0 = clock64();
          do {
            asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(seen) : "r"(smem_u32(s_dep)) : "memory");
            if (clock64() - t0 > 4000000000LL) __trap();
          } while (seen < (uint32_t)(it + 1));
        }
      }
      if (lane == 0) CHAIN_TRACE(4 + 6 * it);
      for (int kb = 0; kb < total_kb; ++kb) {
        mbar_wait_q(&empty_bar[stage], phase ^ 1u);
        uint8_t* a_dst = sA + stage * CH_A_BYTES;
        uint8_t* b_dst = sB + stage * CH_B_BYTES;
        const uint32_t bar = full_leader + stage * 8;
        if (kb < num_kb) {
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * CH_STAGE_BYTES);
            tma_load_2d_2sm(a_dst, ta, bar, kb * 64, m0);
            tma_load_2d_2sm(b_dst, tb, bar, kb * 64, n0 + (int)rank * (BN / 2));
          }
        } else {  // residual stage: 128 rows x two 64-column blocks of plane (rb / 2) at columns n0 + 128 (rb % 2)
          const int rb = kb - num_kb;
          const CUtensorMap* tx = (rb >> 1) ? &tmXl : &tmXh;
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * CH_STAGE_BYTES);
            tma_load_2d_2sm(a_dst, tx, bar, n0 + 128 * (rb & 1), m0);
            tma_load_2d_2sm(b_dst, tx, bar, n0 + 128 * (rb & 1) + 64, m0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) stage = 0, phase ^= 1u;
      }
      if (lane == 0) CHAIN_TRACE(5 + 6 * it);
    }
  } else if (warp == PW + 1) {
    // ===================== MMA issuer (leader CTA) =====================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * 2, BN);
      constexpr uint32_t idesc_res = umma_idesc_bf16(GEMM_BM * 2, 64);
      const uint32_t i_addr = smem_u32(s_ident);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      bool ident_ready = false;
      int it = 0;
      for (int ui = u_begin; ui < u_end; ++ui, ++it) {
        const int code = p.sched[ui];
        const int ph = code >> 28;
        const int num_kb = (ph ? p.K2 : p.K1) / 64;
        const int total_kb = num_kb + (ph ? 0 : CH_RES_KB);
        if (!ph && !ident_ready) {
          mbar_wait_q(ident_bar, 0);
          ident_ready = true;
        }
        mbar_wait_q(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < total_kb; ++kb) {
          mbar_wait_q(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0 && kb == 0) CHAIN_TRACE(6 + 6 * it);
          const uint32_t a_addr = smem_u32(sA + stage * CH_A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * CH_B_BYTES);
          if (kb < num_kb) {
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                              (kb | k) ? 1u : 0u);
              umma_commit_2sm(&empty_bar[stage]);  // ring slot reusable in both CTAs once these MMAs have read it
              if (kb == total_kb - 1) umma_commit_2sm(&tfull_bar[acc]);
            }
          } else {  // acc[:, 64 j .. 64 j + 63] += X_plane block j . I64 for the stage's two blocks
            const uint32_t dj = d_tmem + 128u * (uint32_t)((kb - num_kb) & 1);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(dj, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(i_addr + k * 32), idesc_res, 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(dj + 64u, umma_desc_k_sw128(b_addr + k * 32), umma_desc_k_sw128(i_addr + k * 32), idesc_res, 1u);
              umma_commit_2sm(&empty_bar[stage]);
              if (kb == total_kb - 1) umma_commit_2sm(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        if (lane == 0) CHAIN_TRACE(7 + 6 * it);
        if (++acc == 2) acc = 0, acc_phase ^= 1u;
      }
    }
   } else if (warp == PW + 2) {
    // ===================== dependency scout =====================
    // Walks the unit list ahead of the producer: for a phase-2 unit it waits until the high plane of the LayerNorm rows of
    // the unit's row tile is in L2 (acquire on the counter) and publishes the number of cleared units in shared memory.
    // The producer only reads that word: polling and fencing in the producer itself drained the ring at every unit
    // boundary (+2 k cycles per tile, measured).  No proxy fence: the counter is bumped only after the TMA stores have
    // COMPLETED in L2, which is where the producer's TMA loads read (no L1 in that path, nothing can be stale).
    int it = 0;
    for (int ui = u_begin; ui < u_end; ++ui, ++it) {
      const int code = p.sched[ui];
      if (code >> 28) {
        chain_wait_ge(p.ready + ((code >> 8) & 0xFFFFF), p.ready_target);  // acquire
        if (lane == 0 && p.trace && p.trace[(size_t)blockIdx.x * GEMM_TRACE_SLOTS + 54] == 0) CHAIN_TRACE_NS(54);
      }
      if (lane == 0)
        asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(s_dep)), "r"((uint32_t)(it + 1)) : "memory");
      __syncwarp();
    }
   }
  } else {
    // ===================== epilogue (warps 0..15, both CTAs) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // LayerNorm parameters of this pair's column half (weights, constant over the chain: no dependency wait needed;
    // first read after the `bar.sync 1` of the first LayerNorm unit).  LayerNorm unit u = row tile * halves + half goes
    // to pair u % pairs: with an even pair count the half is pair % 2.
    {
      const int c0 = (pair % halves) * BN;
      for (int i = threadIdx.x; i < BN; i += PT) {
        s_bias1[i] = p.bias1 ? p.bias1[c0 + i] : 0.f;
        s_gamma[i] = p.gamma[c0 + i];
        s_beta[i] = p.beta[c0 + i];
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
    }
    // the sibling kernel's sync words (it is not running: this grid started after it completed)
    for (int i = blockIdx.x * PT + threadIdx.x; i < p.zero_n; i += gridDim.x * PT) p.zero_ptr[i] = 0u;
    for (int i = blockIdx.x * PT + threadIdx.x; i < p.ones_n; i += gridDim.x * PT) p.ones_ptr[i] = ~0ull;
    const int lq = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter (64 columns)
    const int row_in_tile = lq * 32 + lane;
    const uint32_t tempty_leader = mapa_cluster(smem_u32(&tempty_bar[0]), 0);
    const uint32_t wst = smem_u32(s_stage) + warp * GEMM_STG_WARP;
    const int cl = cq * 64;  // first column of this warp's slab inside the 256-column tile
    uint32_t acc = 0, acc_phase = 0, ln_count = 0;
    int staged_n2 = -1;
    int it = 0;
    for (int ui = u_begin; ui < u_end; ++ui, ++it) {
      const int code = p.sched[ui];
      const int ph = code >> 28, m = (code >> 8) & 0xFFFFF, n = code & 0xFF;
      const int n0 = n * BN;
      const int grow0 = m * 256 + (int)rank * GEMM_BM + lq * 32;  // first global row of this warp
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(lq * 32) << 16) + cl;
      auto release_acc = [&]() {  // hand the drained accumulator stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + acc * 8);
      };
      if (ph) {
        if (n0 != staged_n2) {  // the 64 bias values of this column quarter, shared by its 4 warps
          float* wb = s_bias2 + cl;
          const int c = n0 + cl + lane;
          asm volatile("bar.sync %0, 128;" ::"r"(2 + cq) : "memory");  // the quarter has left the previous slice
          if (lq == 0) {
            wb[lane] = p.bias2 ? p.bias2[c] : 0.f;
            wb[lane + 32] = p.bias2 ? p.bias2[c + 32] : 0.f;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(2 + cq) : "memory");
          staged_n2 = n0;
        }
      }
      mbar_wait_q(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (threadIdx.x == 0) CHAIN_TRACE(8 + 6 * it);
      uint32_t v0[32], v1[32];
      tmem_ld32(taddr, v0);
      tc_wait_ld_dep(v0);
      tmem_ld32(taddr + 32, v1);
      if (ph == 0) {
        // ---------------- x = LayerNorm(acc + b)   (acc already holds A1 . W1^T + residual) ----------------
        const int col0 = n0 + cl;  // global column of the slab
        // ---- pass 1: y = acc + bias stays in registers; row sum and sum of squares (packed fp32 pairs: the epilogue is
        //      issue bound -- 16 warps on 4 schedulers) ----
        f32x2 sa = pk2(0.f, 0.f), sb = pk2(0.f, 0.f), qa = pk2(0.f, 0.f), qb = pk2(0.f, 0.f);
        auto pass1 = [&](uint32_t (&v)[32], int cbase) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            asm volatile("" ::: "memory");  // keep the parameter loads next to their use (64 live accumulator registers)
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias1 + cbase + j * 4);
            const f32x2 y0 = add2(pk2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), pk2(b4.x, b4.y));
            const f32x2 y1 = add2(pk2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), pk2(b4.z, b4.w));
            sa = add2(sa, y0), sb = add2(sb, y1);
            qa = fma2(y0, y0, qa), qb = fma2(y1, y1, qb);
            v[4 * j] = __float_as_uint(pk_lo(y0)), v[4 * j + 1] = __float_as_uint(pk_hi(y0));
            v[4 * j + 2] = __float_as_uint(pk_lo(y1)), v[4 * j + 3] = __float_as_uint(pk_hi(y1));
          }
        };
        pass1(v0, cl);
        tc_wait_ld_dep(v1);
        release_acc();  // the whole accumulator slab of this warp is in registers
        pass1(v1, cl + 32);
        if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(57);  // pass 1 done (first unit of the pair, warp 0)
        // ---- row statistics: 4 column quarters through shared memory, the other column half through L2 ----
        float* st = s_stat + (ln_count & 1u) * (2 * 4 * 128);
        ++ln_count;
        st[cq * 128 + row_in_tile] = (pk_lo(sa) + pk_hi(sa)) + (pk_lo(sb) + pk_hi(sb));
        st[512 + cq * 128 + row_in_tile] = (pk_lo(qa) + pk_hi(qa)) + (pk_lo(qb) + pk_hi(qb));
        asm volatile("bar.sync 1, 512;" ::: "memory");
        float tot_s = (st[row_in_tile] + st[128 + row_in_tile]) + (st[256 + row_in_tile] + st[384 + row_in_tile]);
        float tot_q = (st[512 + row_in_tile] + st[640 + row_in_tile]) + (st[768 + row_in_tile] + st[896 + row_in_tile]);
        if (halves == 2) {
          const size_t mine = ((size_t)(m * 2 + n) * 2 + rank) * 128 + row_in_tile;
          const size_t theirs = ((size_t)(m * 2 + (n ^ 1)) * 2 + rank) * 128 + row_in_tile;
          // one 64-bit word per row: the sum of squares is >= 0 (or a NaN, made positive), so a posted word never
          // equals the all-ones "not posted" pattern, and the single store is its own flag (no fence needed)
          if (cq == 0)
            st_relaxed_gpu_u64(p.stats + mine, (unsigned long long)__float_as_uint(tot_s) |
                                                   ((unsigned long long)(__float_as_uint(tot_q) & 0x7fffffffu) << 32));
          unsigned long long o = ld_relaxed_gpu_u64(p.stats + theirs);
          if (o == ~0ull) {
            const long long t0 = clock64();
            while ((o = ld_relaxed_gpu_u64(p.stats + theirs)) == ~0ull) {
              __nanosleep(20);
              if (clock64() - t0 > 4000000000LL) __trap();
            }
          }
          tot_s += __uint_as_float((uint32_t)o), tot_q += __uint_as_float((uint32_t)(o >> 32));
          // a + b == b + a: both halves normalise with bit-identical statistics
        }
        if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(58);  // statistics complete (first unit of the pair)
        const float inv_n = 1.0f / (float)p.N1;
        const float mean = tot_s * inv_n;
        const float var = fmaxf(tot_q * inv_n - mean * mean, 0.f);  // biased variance (F.layer_norm), fp32
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        // ---- pass 2: normalise + affine (kept in the registers), hi plane -> staging -> TMA store, then lo plane ----
        if (elect_one()) bulk_wait_read<0>();  // the previous unit's store has finished reading the staging tile
        __syncwarp();
        const f32x2 rstd2 = pk2(rstd, rstd), nmr2 = pk2(nmr, nmr);
        // y = ((acc - mean) rstd) gamma + beta on packed pairs; the hi plane bf16(y) goes to the staging tile, the
        // register pair is replaced by the lo plane's input y - bf16(y) (exact in fp32)
        auto norm_hi = [&](uint32_t (&v)[32], int cbase, int j0) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            uint32_t hi[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              asm volatile("" ::: "memory");  // as above
              const int c = jj * 8 + e * 2;
              const float2 g2 = *reinterpret_cast<const float2*>(s_gamma + cbase + c);
              const float2 be2 = *reinterpret_cast<const float2*>(s_beta + cbase + c);
              const f32x2 y = fma2(fma2(pk2(__uint_as_float(v[c]), __uint_as_float(v[c + 1])), rstd2, nmr2), pk2(g2.x, g2.y),
                                   pk2(be2.x, be2.y));
              hi[e] = pack_bf16x2(pk_lo(y), pk_hi(y));
              const f32x2 r = add2(y, pk2(-bf16lo_f32(hi[e]), -bf16hi_f32(hi[e])));
              v[c] = __float_as_uint(pk_lo(r)), v[c + 1] = __float_as_uint(pk_hi(r));
            }
            sts128(wst + stg128_off(lane, j0 + jj), make_uint4(hi[0], hi[1], hi[2], hi[3]));
          }
        };
        norm_hi(v0, cl, 0);
        norm_hi(v1, cl + 32, 4);
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one() && !(p.dbg & 1)) {
          tma_store_2d(&tmXh_st, wst, col0, grow0);  // rows past M are clipped by the tensor map
          bulk_commit();
        }
        if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(59);  // hi plane staged + store issued
        if (elect_one()) bulk_wait_read<0>();  // the hi store has finished reading the tile
        __syncwarp();
        auto stage_lo = [&](const uint32_t (&v)[32], int j0) {  // lo = bf16(y - bf16(y)), the difference is in v
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            uint32_t lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              lo[e] = pack_bf16x2(__uint_as_float(v[jj * 8 + 2 * e]), __uint_as_float(v[jj * 8 + 2 * e + 1]));
            sts128(wst + stg128_off(lane, j0 + jj), make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
        };
        if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(60);  // hi store has read the tile
        stage_lo(v0, 0);
        stage_lo(v1, 4);
        fence_proxy_async_smem();
        __syncwarp();
        // The HIGH plane (the phase-2 A operand) is complete in L2 -> bump the row tile's counter, read by the scout warps
        // of every pair.  The low plane is only read by the next kernel's residual blocks: it completes before this CTA
        // retires (bulk_wait<0> at the end).
        if (elect_one()) {
          if (!(p.dbg & 1)) {
            tma_store_2d(&tmXl_st, wst, col0, grow0);
            bulk_commit();
          }
          if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(61);  // lo plane staged + store issued
          bulk_wait<1>();  // all but the newest group: the hi-plane store of this warp has completed
          if (threadIdx.x == 0 && it == 0) {
            CHAIN_TRACE(62);
            CHAIN_TRACE_NS(55);
          }
          // The store has COMPLETED: the rows are in L2 (the point of coherence; TMA loads read there) before the counter
          // update is even issued, so a relaxed update suffices -- a release would add a gpu-scope fence, 3.3 k cycles
          // on the critical path of every LayerNorm unit (measured).  The readers acquire.
          red_relaxed_gpu_add(p.ready + m, 1u);
        }
        __syncwarp();
      } else {
        // ---------------- bias (+ GELU) -> bf16: the warp's 32 x 64 slab leaves as one TMA store ----------------
        const float* wb = s_bias2 + cl;
        if (elect_one()) bulk_wait_read<0>();  // the previous unit's store has finished reading the staging tile
        __syncwarp();
        auto half = [&](const uint32_t (&vv)[32], int j0) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            const uint32_t* v = vv + jj * 8;
            uint32_t o[4];
            if constexpr (EPI2 == CHAIN_GELU) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {  // packed pairs: bias add + GELU as FADD2 / FFMA2 chains
                const float2 b2 = *reinterpret_cast<const float2*>(wb + j * 8 + 2 * e);
                const f32x2 y = gelu_erf2(add2(pk2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), pk2(b2.x, b2.y)));
                o[e] = pack_bf16x2(pk_lo(y), pk_hi(y));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 b2 = *reinterpret_cast<const float2*>(wb + j * 8 + 2 * e);
                o[e] = pack_bf16x2(__uint_as_float(v[2 * e]) + b2.x, __uint_as_float(v[2 * e + 1]) + b2.y);
              }
            }
            sts128(wst + stg128_off(lane, j), make_uint4(o[0], o[1], o[2], o[3]));
          }
        };
        half(v0, 0);
        tc_wait_ld_dep(v1);
        release_acc();
        half(v1, 4);
        fence_proxy_async_smem();
        __syncwarp();
        if (elect_one() && !(p.dbg & 1)) {
          tma_store_2d(&tmC2, wst, n0 + cl, grow0);  // rows past M are clipped
          bulk_commit();
        }
      }
      if (threadIdx.x == 0) CHAIN_TRACE(9 + 6 * it);
      if (++acc == 2) acc = 0, acc_phase ^= 1u;
    }
    if (elect_one()) bulk_wait<0>();  // this thread's TMA stores have been performed before the CTA retires
  }
  tc_fence_before();
  cluster_sync_all();  // the peer may still signal our barriers / read our TMEM half
  if (threadIdx.x == 0) CHAIN_TRACE(3);
  if (warp == PW + 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ---- host side ------------------------------------------------------------------------------------
// Static schedule of one chain kernel: per CTA pair the list of unit codes (phase << 28 | row tile << 8 | column tile).
struct ChainSchedule {
  int pairs = 0, tiles_m = 0, halves = 0, tiles_n2 = 0;
  std::vector<int> off, units;
};

// List scheduling on cost estimates (cycles): LayerNorm units u = m * halves + h go round-robin to pair u % pairs (the
// two column halves of a row tile land on neighbouring pairs in the same round); every phase-2 tile, in row-tile order,
// goes to the pair that can finish it first given when its row tile's LayerNorm units end.
inline ChainSchedule build_chain_schedule(int M, int N1, int N2, int slots, double cost_ln, double cost_ln_tail,
                                          double cost_t2) {
  ChainSchedule s;
  s.tiles_m = (M + 255) / 256;
  s.halves = N1 / CH_BN;
  s.tiles_n2 = N2 / CH_BN;
  const int u1 = s.tiles_m * s.halves, u2 = s.tiles_m * s.tiles_n2;
  int pairs = slots < (u1 + u2) ? slots : (u1 + u2);
  if (s.halves == 2) pairs &= ~1;  // neighbours (2k, 2k + 1) exchange statistics: keep the pair count even
  if (pairs < s.halves) pairs = s.halves;
  s.pairs = pairs;
  std::vector<std::vector<int>> lists(pairs);
  std::vector<double> busy(pairs, 0.0), ready(s.tiles_m, 0.0);
  for (int u = 0; u < u1; ++u) {
    const int pr = u % pairs, m = u / s.halves, h = u % s.halves;
    lists[pr].push_back((0 << 28) | (m << 8) | h);
    busy[pr] += cost_ln;
    const double done = busy[pr] + cost_ln_tail;  // the rows are in L2 an epilogue after the mainloop
    if (done > ready[m]) ready[m] = done;
  }
  // phase-2 tiles in order of the estimated time their row tile is ready (ties: row tile index): the row tiles of pairs
  // that run two LayerNorm units come last -- those pairs finish even their FIRST unit late (its epilogue shares the
  // SM with the second unit's mainloop)
  std::vector<int> order(s.tiles_m);
  for (int m = 0; m < s.tiles_m; ++m) order[m] = m;
  {
    std::vector<int> per_pair(pairs, 0);
    for (int u = 0; u < u1; ++u) ++per_pair[u % pairs];
    std::vector<double> key(s.tiles_m, 0.0);
    for (int u = 0; u < u1; ++u) {
      const int m = u / s.halves;
      const double k = ready[m] + (per_pair[u % pairs] > 1 ? 0.5 * cost_ln_tail : 0.0);
      if (k > key[m]) key[m] = k;
    }
    for (int m = 0; m < s.tiles_m; ++m) ready[m] = key[m];
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  }
  for (int mi = 0; mi < s.tiles_m; ++mi) {
    const int m = order[mi];
    for (int n = 0; n < s.tiles_n2; ++n) {
      int best = 0;
      double best_end = 1e300;
      for (int pr = 0; pr < pairs; ++pr) {
        const double start = busy[pr] > ready[m] ? busy[pr] : ready[m];
        if (start + cost_t2 < best_end - 1e-9) best_end = start + cost_t2, best = pr;
      }
      lists[best].push_back((1 << 28) | (m << 8) | n);
      busy[best] = best_end;
    }
  }
  s.off.assign(pairs + 1, 0);
  for (int pr = 0; pr < pairs; ++pr) {
    s.off[pr + 1] = s.off[pr] + (int)lists[pr].size();
    s.units.insert(s.units.end(), lists[pr].begin(), lists[pr].end());
  }
  return s;
}

template <int EPI2>
int configure_gemm_chain() {
  TAMF_CUDA_CHECK(cudaFuncSetAttribute(gemm_chain_kernel<EPI2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES));
  return TAMF_OK;
}

struct ChainMaps {
  // A1 / B1: phase-1 operands (box {64, 128}); Xh / Xl: the residual planes, box {64, 128} (Xh is also the phase-2 A
  // operand); B2 / C2: phase-2 weights and bf16 output (box {64, 32}), null when N2 == 0; Xh_st / Xl_st: the planes with
  // box {64, 32} (LayerNorm stores); I: chain_identity_map()
  const CUtensorMap *A1, *B1, *Xh, *Xl, *B2, *C2, *Xh_st, *Xl_st, *I;
};

// The 64 x 64 bf16 identity the residual blocks are multiplied with (one per device, never freed) and its tensor map.
inline int chain_identity_map(CUtensorMap* out) {
  static std::mutex mu;
  static std::map<int, void*> per_dev;
  int dev = 0;
  TAMF_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  void*& d = per_dev[dev];
  if (!d) {
    std::vector<uint16_t> h(64 * 64, 0);
    for (int i = 0; i < 64; ++i) h[i * 64 + i] = 0x3F80;  // bf16 1.0
    TAMF_CUDA_CHECK(cudaMalloc(&d, h.size() * 2));
    TAMF_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  }
  return make_tmap_2d_bf16(out, d, 64, 64, 128, 64, 32);
}

template <int EPI2>
int launch_gemm_chain(const ChainMaps& tm, const ChainParams& p, int pairs, cudaStream_t stream) {
  TAMF_REQUIRE(p.N1 == 256 || p.N1 == 512, TAMF_E_BADARG, "gemm_chain: LayerNorm width must be 256 or 512");
  TAMF_REQUIRE(p.K1 > 0 && p.K1 % 64 == 0 && p.N2 % 256 == 0 && (p.N2 == 0 || (p.K2 > 0 && p.K2 % 64 == 0)), TAMF_E_BADARG,
               "gemm_chain: K must be a multiple of 64, N2 a multiple of 256");
  TAMF_REQUIRE(tm.A1 && tm.B1 && tm.Xh && tm.Xl && tm.Xh_st && tm.Xl_st && tm.I && (p.N2 == 0 || (tm.B2 && tm.C2)),
               TAMF_E_BADARG, "gemm_chain: missing tensor map");
  TAMF_REQUIRE(pairs >= 1 && 2 * pairs <= num_sms(), TAMF_E_BADARG, "gemm_chain: the grid must be co-resident");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
  ++na;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const CUtensorMap& b2 = tm.B2 ? *tm.B2 : *tm.B1;  // unused maps are passed as copies
  const CUtensorMap& c2 = tm.C2 ? *tm.C2 : *tm.A1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_chain_kernel<EPI2>, *tm.A1, *tm.B1, *tm.Xh, *tm.Xl, b2, c2, *tm.Xh_st,
                                     *tm.Xl_st, *tm.I, p);
  count_launch();
  if (e != cudaSuccess) {
    set_error(std::string("gemm_chain launch failed: ") + cudaGetErrorString(e));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

}  // namespace tamf
