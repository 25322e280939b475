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
//   * a 512-wide LayerNorm row is computed by TWO pairs (column halves): each keeps its 128 x 256 slab of
//     y = acc + b + x in REGISTERS (64 per thread), the halves swap per-row (sum, sum of squares) through L2
//     (1 KB per CTA, release/acquire flags), and all 148 SMs store the result;
//   * the units of phase 2 are scheduled into the gaps of phase 1 by a host-built static schedule (list scheduling on
//     per-unit cost estimates); a phase-2 tile of row tile m waits for a counter that the LayerNorm epilogues of
//     row tile m bump after their TMA stores have completed (release / acquire through L2, proxy fences on both
//     sides).  All CTAs of the grid are co-resident (grid <= SM count, 1 CTA / SM) and waits only point at units
//     that come EARLIER in some pair's list, so the schedule cannot deadlock; every spin is bounded (trap).
// The sync words of this kernel are zeroed by its sibling (the other chain kernel of the layer), which runs in between.
#pragma once
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
  unsigned* sflag;                    // [tiles_m*halves*2*4]  row statistics of (unit, CTA rank, lane quarter) posted
  float2* stats;                      // [tiles_m*halves*2*128] (sum, sum of squares) over the CTA's 256 columns
  unsigned ready_target;              // halves * 2 CTAs * 16 warps
  unsigned* zero_ptr;                 // the sibling kernel's sync words, cleared here (after the dependency wait)
  int zero_n;
  long long* trace;                   // debug only: [grid][GEMM_TRACE_SLOTS] clock64 stamps
  int dbg;
};

constexpr int CH_STAGES = 4;
constexpr int CH_BN = 256;
constexpr int CH_A_BYTES = GEMM_BM * 64 * 2;       // 128 rows x 64 k
constexpr int CH_B_BYTES = (CH_BN / 2) * 64 * 2;   // this CTA's half of the 256 W rows
constexpr int CH_STAGE_BYTES = CH_A_BYTES + CH_B_BYTES;
constexpr int CH_PIPE_BYTES = CH_STAGES * CH_STAGE_BYTES;
constexpr int CH_STG_BYTES = GEMM_EPI_WARPS * GEMM_STG_WARP;
constexpr int CH_CTRL_BYTES = 1024;
// bias1 | gamma | beta | bias2 (256 floats each) | row statistics [2 buffers][sum, sq][4 column quarters][128 rows]
constexpr int CH_PARAM_BYTES = 4 * CH_BN * 4 + 2 * 2 * 4 * 128 * 4;
constexpr int CH_SMEM_BYTES = 1024 + CH_PIPE_BYTES + CH_STG_BYTES + CH_CTRL_BYTES + CH_PARAM_BYTES;
static_assert(CH_SMEM_BYTES <= GEMM_SMEM_MAX, "shared memory budget (227 KB) exceeded");

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release_gpu_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// generic proxy <-> async proxy (TMA) ordering for every state space
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// Lane 0 polls `*flag >= target` (acquire, gpu scope) with a bounded spin; the warp leaves together.
__device__ __forceinline__ void chain_wait_ge(const unsigned* flag, unsigned target) {
  if (lane_id() == 0) {
    if (ld_acquire_gpu(flag) < target) {
      const long long t0 = clock64();
      while (ld_acquire_gpu(flag) < target) {
        __nanosleep(40);
        if (clock64() - t0 > 4000000000LL) {
          printf("tamf: chain flag timeout block %d thread %d (%u < %u)\n", blockIdx.x, threadIdx.x, ld_acquire_gpu(flag),
                 target);
          __trap();
        }
      }
    }
  }
  __syncwarp();
}

#define CHAIN_TRACE(slot)                                                                                        \
  do {                                                                                                           \
    if (p.trace && (slot) < GEMM_TRACE_SLOTS) p.trace[(size_t)blockIdx.x * GEMM_TRACE_SLOTS + (slot)] = clock64(); \
  } while (0)

template <int EPI2>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_chain_kernel(const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                      const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                      const __grid_constant__ CUtensorMap tmC2, const __grid_constant__ CUtensorMap tmXh,
                      const __grid_constant__ CUtensorMap tmXl, const ChainParams p) {
  constexpr int STAGES = CH_STAGES, PW = GEMM_EPI_WARPS, PT = GEMM_EPI_THREADS, BN = CH_BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;  // identical in both CTAs of a pair
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * CH_A_BYTES;
  uint8_t* s_stage = smem + CH_PIPE_BYTES;
  uint8_t* ctrl = s_stage + CH_STG_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);  // [STAGES] (the leader's copy is the live one)
  uint64_t* empty_bar = full_bar + STAGES;                  // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;                 // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                     // [2]      (the leader's copy is the live one)
  uint64_t* rbar = tempty_bar + 2;                          // [16]     residual tile of epilogue warp w landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + PW);
  float* s_bias1 = reinterpret_cast<float*>(ctrl + CH_CTRL_BYTES);
  float* s_gamma = s_bias1 + BN;
  float* s_beta = s_gamma + BN;
  float* s_bias2 = s_beta + BN;
  float* s_stat = s_bias2 + BN;  // [2][2][4][128]

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
    if (p.N2 > 0) {
      tma_prefetch_desc(&tmA2);
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
    for (int i = 0; i < PW; ++i) mbar_init(&rbar[i], 1);
    fence_mbar_init();
  }
  if (warp == PW + 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  if (warp < PW) {
    // LayerNorm parameters of this pair's column half (weights: constant over the chain, staged before the dependency
    // wait).  Every LayerNorm unit of a pair has the same column half (host schedule: an even number of pairs).
    const int code0 = u_begin < u_end ? p.sched[u_begin] : -1;
    if (code0 >= 0 && (code0 >> 28) == 0) {
      const int c0 = (code0 & 0xFF) * BN;
      for (int i = threadIdx.x; i < BN; i += PT) {
        s_bias1[i] = p.bias1 ? p.bias1[c0 + i] : 0.f;
        s_gamma[i] = p.gamma[c0 + i];
        s_beta[i] = p.beta[c0 + i];
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) CHAIN_TRACE(1);
  pdl_launch_dependents();
  pdl_wait();  // everything the previous kernels wrote (activations, our zeroed sync words) is visible from here on
  if (threadIdx.x == 0) CHAIN_TRACE(2);

  if (warp == PW) {
    // ===================== TMA producer (both CTAs of the pair) =====================
    uint32_t stage = 0, phase = 0;
    int it = 0;
    const uint32_t full_leader = mapa_cluster(smem_u32(&full_bar[0]), 0);
    for (int ui = u_begin; ui < u_end; ++ui, ++it) {
      const int code = p.sched[ui];
      const int ph = code >> 28, m = (code >> 8) & 0xFFFFF, n = code & 0xFF;
      const int m0 = m * 256 + (int)rank * GEMM_BM, n0 = n * BN;
      const CUtensorMap* ta = ph ? &tmA2 : &tmA1;
      const CUtensorMap* tb = ph ? &tmB2 : &tmB1;
      const int num_kb = (ph ? p.K2 : p.K1) / 64;
      if (ph) {  // the LayerNorm rows of this row tile must have landed (both column halves, both CTAs)
        chain_wait_ge(p.ready + m, p.ready_target);
        fence_proxy_async_all();
      }
      if (lane == 0) CHAIN_TRACE(4 + 6 * it);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* a_dst = sA + stage * CH_A_BYTES;
        uint8_t* b_dst = sB + stage * CH_B_BYTES;
        const uint32_t bar = full_leader + stage * 8;
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * CH_STAGE_BYTES);
          tma_load_2d_2sm(a_dst, ta, bar, kb * 64, m0);
          tma_load_2d_2sm(b_dst, tb, bar, kb * 64, n0 + (int)rank * (BN / 2));
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
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      int it = 0;
      for (int ui = u_begin; ui < u_end; ++ui, ++it) {
        const int code = p.sched[ui];
        const int num_kb = ((code >> 28) ? p.K2 : p.K1) / 64;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0 && kb == 0) CHAIN_TRACE(6 + 6 * it);
          const uint32_t a_addr = smem_u32(sA + stage * CH_A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * CH_B_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16_2sm(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                            (kb | k) ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage]);  // ring slot reusable in both CTAs once these MMAs have read it
            if (kb == num_kb - 1) umma_commit_2sm(&tfull_bar[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        if (lane == 0) CHAIN_TRACE(7 + 6 * it);
        if (++acc == 2) acc = 0, acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 0..15, both CTAs) =====================
    // the sibling kernel's sync words (it is not running: this grid started after it completed)
    for (int i = blockIdx.x * PT + threadIdx.x; i < p.zero_n; i += gridDim.x * PT) p.zero_ptr[i] = 0u;
    const int lq = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter (64 columns)
    const int row_in_tile = lq * 32 + lane;
    const uint32_t tempty_leader = mapa_cluster(smem_u32(&tempty_bar[0]), 0);
    const uint32_t wst = smem_u32(s_stage) + warp * GEMM_STG_WARP;
    uint64_t* rb = &rbar[warp];
    uint32_t acc = 0, acc_phase = 0, rpar = 0, ln_count = 0;
    int staged_n2 = -1;
    int it = 0;
    for (int ui = u_begin; ui < u_end; ++ui, ++it) {
      const int code = p.sched[ui];
      const int ph = code >> 28, m = (code >> 8) & 0xFFFFF, n = code & 0xFF;
      const int n0 = n * BN;
      const int grow0 = m * 256 + (int)rank * GEMM_BM + lq * 32;  // first global row of this warp
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(lq * 32) << 16) + cq * 64;
      auto release_acc = [&]() {  // hand the drained accumulator stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + acc * 8);
      };
      if (ph == 0) {
        // ---------------- x = LayerNorm(x + acc + b) for a 32-row x 64-column slab per warp ----------------
        const int col0 = n0 + cq * 64;  // global column of the slab
        const int cl0 = cq * 64;        // its column inside the CTA's half (s_bias1 / s_gamma / s_beta)
        // residual (bf16 hi + lo planes, thread = row): columns 0..31 go to registers, 32..63 wait in the staging tile.
        // None of this depends on the accumulator: it runs while the mainloop of this unit is still in flight.
        if (elect_one()) {
          bulk_wait_read<0>();  // the previous unit's TMA store has finished reading the staging tile
          mbar_arrive_expect_tx(rb, 4096);
          tma_load_2d_u32(wst, &tmXh, smem_u32(rb), col0, grow0);
          tma_load_2d_u32(wst + 2048, &tmXl, smem_u32(rb), col0, grow0);
        }
        mbar_wait(rb, rpar);
        rpar ^= 1u;
        uint4 rh[4], rl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          rh[j] = lds128(wst + stg64_off(lane, j));
          rl[j] = lds128(wst + 2048 + stg64_off(lane, j));
        }
        __syncwarp();  // every lane holds its row: the tile may be refilled
        if (elect_one()) {
          mbar_arrive_expect_tx(rb, 4096);
          tma_load_2d_u32(wst, &tmXh, smem_u32(rb), col0 + 32, grow0);
          tma_load_2d_u32(wst + 2048, &tmXl, smem_u32(rb), col0 + 32, grow0);
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (threadIdx.x == 0) CHAIN_TRACE(8 + 6 * it);
        // ---- pass 1: y = acc + bias + residual stays in registers; row sum and sum of squares ----
        uint32_t v0[32], v1[32];
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
        // `from_regs`: the residual pieces come from rh / rl (columns 0..31); otherwise straight from the staging tile
        auto pass1 = [&](uint32_t (&v)[32], auto from_regs, int cbase) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 8 columns per 16-byte piece of each plane
            uint4 h4, l4;
            if constexpr (decltype(from_regs)::value) {
              h4 = rh[j], l4 = rl[j];
            } else {
              h4 = lds128(wst + stg64_off(lane, j)), l4 = lds128(wst + 2048 + stg64_off(lane, j));
            }
            const float4 ba = *reinterpret_cast<const float4*>(s_bias1 + cbase + j * 8);
            const float4 bb = *reinterpret_cast<const float4*>(s_bias1 + cbase + j * 8 + 4);
            const float y0 = __uint_as_float(v[8 * j]) + ba.x + (bf16lo_f32(h4.x) + bf16lo_f32(l4.x));
            const float y1 = __uint_as_float(v[8 * j + 1]) + ba.y + (bf16hi_f32(h4.x) + bf16hi_f32(l4.x));
            const float y2 = __uint_as_float(v[8 * j + 2]) + ba.z + (bf16lo_f32(h4.y) + bf16lo_f32(l4.y));
            const float y3 = __uint_as_float(v[8 * j + 3]) + ba.w + (bf16hi_f32(h4.y) + bf16hi_f32(l4.y));
            const float y4 = __uint_as_float(v[8 * j + 4]) + bb.x + (bf16lo_f32(h4.z) + bf16lo_f32(l4.z));
            const float y5 = __uint_as_float(v[8 * j + 5]) + bb.y + (bf16hi_f32(h4.z) + bf16hi_f32(l4.z));
            const float y6 = __uint_as_float(v[8 * j + 6]) + bb.z + (bf16lo_f32(h4.w) + bf16lo_f32(l4.w));
            const float y7 = __uint_as_float(v[8 * j + 7]) + bb.w + (bf16hi_f32(h4.w) + bf16hi_f32(l4.w));
            s0 += y0 + y4, s1 += y1 + y5, s2 += y2 + y6, s3 += y3 + y7;
            q0 = fmaf(y0, y0, q0), q1 = fmaf(y1, y1, q1), q2 = fmaf(y2, y2, q2), q3 = fmaf(y3, y3, q3);
            q0 = fmaf(y4, y4, q0), q1 = fmaf(y5, y5, q1), q2 = fmaf(y6, y6, q2), q3 = fmaf(y7, y7, q3);
            v[8 * j] = __float_as_uint(y0), v[8 * j + 1] = __float_as_uint(y1);
            v[8 * j + 2] = __float_as_uint(y2), v[8 * j + 3] = __float_as_uint(y3);
            v[8 * j + 4] = __float_as_uint(y4), v[8 * j + 5] = __float_as_uint(y5);
            v[8 * j + 6] = __float_as_uint(y6), v[8 * j + 7] = __float_as_uint(y7);
          }
        };
        tmem_ld32(taddr, v0);
        tc_wait_ld();
        pass1(v0, std::true_type{}, cl0);
        tmem_ld32(taddr + 32, v1);
        mbar_wait(rb, rpar);  // residual columns 32..63 have landed in the staging tile
        rpar ^= 1u;
        tc_wait_ld();
        release_acc();  // the whole accumulator slab of this warp is in registers
        pass1(v1, std::false_type{}, cl0 + 32);
        __syncwarp();  // every lane has read the staging tile: it becomes the output staging of pass 2
        // ---- row statistics: 4 column quarters through shared memory, the other column half through L2 ----
        float* st = s_stat + (ln_count & 1u) * (2 * 4 * 128);
        ++ln_count;
        st[cq * 128 + row_in_tile] = (s0 + s1) + (s2 + s3);
        st[512 + cq * 128 + row_in_tile] = (q0 + q1) + (q2 + q3);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        float tot_s = (st[row_in_tile] + st[128 + row_in_tile]) + (st[256 + row_in_tile] + st[384 + row_in_tile]);
        float tot_q = (st[512 + row_in_tile] + st[640 + row_in_tile]) + (st[768 + row_in_tile] + st[896 + row_in_tile]);
        if (halves == 2) {
          const int mine = ((m * 2 + n) * 2 + (int)rank), theirs = ((m * 2 + (n ^ 1)) * 2 + (int)rank);
          if (cq == 0) {
            p.stats[(size_t)mine * 128 + row_in_tile] = make_float2(tot_s, tot_q);
            __syncwarp();
            if (lane == 0) {
              __threadfence();
              st_release_gpu(p.sflag + mine * 4 + lq, 1u);
            }
          }
          chain_wait_ge(p.sflag + theirs * 4 + lq, 1u);
          const float2 o = __ldcg(p.stats + (size_t)theirs * 128 + row_in_tile);
          tot_s += o.x, tot_q += o.y;  // a + b == b + a: both halves normalise with bit-identical statistics
        }
        if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(58);  // statistics complete (first unit of the pair)
        const float inv_n = 1.0f / (float)p.N1;
        const float mean = tot_s * inv_n;
        const float var = fmaxf(tot_q * inv_n - mean * mean, 0.f);  // biased variance (F.layer_norm), fp32
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        // ---- pass 2: normalise + affine -> hi / lo bf16 tiles -> TMA stores ----
        auto pass2 = [&](const uint32_t (&v)[32], int cbase, int gcol) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = j * 8 + e * 2;
              const float2 g2 = *reinterpret_cast<const float2*>(s_gamma + cbase + c);
              const float2 be2 = *reinterpret_cast<const float2*>(s_beta + cbase + c);
              const float y0 = fmaf(fmaf(__uint_as_float(v[c]), rstd, nmr), g2.x, be2.x);
              const float y1 = fmaf(fmaf(__uint_as_float(v[c + 1]), rstd, nmr), g2.y, be2.y);
              split_bf16x2(y0, y1, hi[e], lo[e]);
            }
            sts128(wst + stg64_off(lane, j), make_uint4(hi[0], hi[1], hi[2], hi[3]));
            sts128(wst + 2048 + stg64_off(lane, j), make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one() && !(p.dbg & 1)) {
            tma_store_2d(&tmXh, wst, gcol, grow0);  // rows past M are clipped by the tensor map
            tma_store_2d(&tmXl, wst + 2048, gcol, grow0);
            bulk_commit();
          }
        };
        pass2(v0, cl0, col0);
        if (elect_one()) bulk_wait_read<0>();  // the stores of columns 0..31 have finished reading the staging tile
        __syncwarp();
        pass2(v1, cl0 + 32, col0 + 32);
        // the rows are complete in L2 -> bump the row tile's counter (read by the phase-2 producers of every pair)
        if (elect_one()) {
          bulk_wait<0>();
          fence_proxy_async_all();
          __threadfence();
          red_release_gpu_add(p.ready + m, 1u);
        }
        __syncwarp();
      } else {
        // ---------------- bias (+ GELU) -> bf16: the warp's 32 x 64 slab leaves as one TMA store ----------------
        const int cl = cq * 64;
        float* wb = s_bias2 + cl;
        if (n0 != staged_n2) {  // the 64 bias values of this column quarter, shared by its 4 warps
          const int c = n0 + cl + lane;
          asm volatile("bar.sync %0, 128;" ::"r"(2 + cq) : "memory");  // the quarter has left the previous slice
          if (lq == 0) {
            wb[lane] = p.bias2 ? p.bias2[c] : 0.f;
            wb[lane + 32] = p.bias2 ? p.bias2[c + 32] : 0.f;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(2 + cq) : "memory");
          staged_n2 = n0;
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (threadIdx.x == 0) CHAIN_TRACE(8 + 6 * it);
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr, v0);
        tc_wait_ld_dep(v0);
        tmem_ld32(taddr + 32, v1);
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
  for (int m = 0; m < s.tiles_m; ++m) {
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
  const CUtensorMap *A1, *B1, *A2, *B2, *C2, *Xh, *Xl;  // A2 / B2 / C2 may be null when N2 == 0
};

template <int EPI2>
int launch_gemm_chain(const ChainMaps& tm, const ChainParams& p, int pairs, cudaStream_t stream) {
  TAMF_REQUIRE(p.N1 == 256 || p.N1 == 512, TAMF_E_BADARG, "gemm_chain: LayerNorm width must be 256 or 512");
  TAMF_REQUIRE(p.K1 > 0 && p.K1 % 64 == 0 && p.N2 % 256 == 0 && (p.N2 == 0 || (p.K2 > 0 && p.K2 % 64 == 0)), TAMF_E_BADARG,
               "gemm_chain: K must be a multiple of 64, N2 a multiple of 256");
  TAMF_REQUIRE(tm.A1 && tm.B1 && tm.Xh && tm.Xl && (p.N2 == 0 || (tm.A2 && tm.B2 && tm.C2)), TAMF_E_BADARG,
               "gemm_chain: missing tensor map");
  TAMF_REQUIRE(pairs >= 1 && 2 * pairs <= num_sms(), TAMF_E_BADARG, "gemm_chain: the grid must be co-resident");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(GEMM_THREADS);
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
  const CUtensorMap& a2 = tm.A2 ? *tm.A2 : *tm.A1;  // unused maps are passed as copies
  const CUtensorMap& b2 = tm.B2 ? *tm.B2 : *tm.B1;
  const CUtensorMap& c2 = tm.C2 ? *tm.C2 : *tm.A1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_chain_kernel<EPI2>, *tm.A1, *tm.B1, a2, b2, c2, *tm.Xh, *tm.Xl, p);
  count_launch();
  if (e != cudaSuccess) {
    set_error(std::string("gemm_chain launch failed: ") + cudaGetErrorString(e));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

}  // namespace tamf
