// attn_tc.cuh -- frame-sequence self-attention on the 5th-gen tensor cores: softmax(Q K^T / sqrt(hd)) V, no mask (the
// reference applies none: nn.TransformerEncoderLayer built at interaction_segment_mdm.py:63-70, applied :171).
//
// One CTA per (head, sequence).  S <= 176 tokens, so the whole head of one sequence is two 128-row query tiles against
// 176 keys and the softmax is single-pass:
//   * Q, K, V (176 x hd each) arrive as 128B-swizzled TMA boxes of the [B][S][3d] view of the packed in_proj output
//     (tokens >= S zero-filled by the tensor map);
//   * scores:  tcgen05.mma M=128, N=176, K=hd, both operands K-major from shared memory, fp32 accumulators in TMEM
//     (tile 0 -> columns [0,176), tile 1 -> [176,352));
//   * softmax: thread = query row; the 176 scores of the row come out of TMEM once (tcgen05.ld), fp32 max / exp2 /
//     sum in registers, probabilities rounded to bf16 and written to shared memory in the K-major 128B-swizzled
//     layout the tensor core reads its A operand in;
//   * P V:     tcgen05.mma M=128, N=hd, K=176 with V as the MN-major B operand (V is [key][hd] in shared memory exactly
//     as TMA delivered it -- no transpose pass); O tile 0 -> TMEM columns [352,480), O tile 1 reuses [0,128) (the
//     scores of tile 0 are dead by then);
//   * epilogue: O * (1 / row sum) -> bf16 -> warp-private swizzled staging in the dead Q/K region -> TMA stores of the
//     [B][S][d] view (rows >= S clipped).
// Warps 0-3 own the rows of tile 0, warps 4-5 rows 128..191 (tile 1: at most 48 valid rows), warp 6 issues TMA and MMA.
// Tile-1 rows beyond the 176 loaded tokens read whatever follows in shared memory: each accumulator row depends on
// its own A row only, and those rows are never stored.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace tamf {

constexpr int ATC_KP = 176;         // padded token count (multiple of 16)
constexpr int ATC_THREADS = 7 * 32;
constexpr int ATC_TMEM_COLS = 512;
constexpr int ATC_O0_COL = 2 * ATC_KP;  // 352

template <int HD>
struct AttnTcCfg {
  static constexpr int NB = HD / 64;           // 64-column blocks per Q / K / V operand
  static constexpr int BLK = ATC_KP * 128;     // bytes of one [176 rows][128 B] block
  static constexpr int P0_BLK = 128 * 128;     // tile 0 probabilities: 3 blocks of [128 rows][128 B]
  static constexpr int P1_BLK = 64 * 128;      // tile 1 probabilities: 3 blocks of [64 rows][128 B]
  static constexpr int OFF_P1 = 0;             // first, so that the 128-row MMA reads past its 64 rows stay in bounds
  static constexpr int OFF_P0 = 3 * P1_BLK;
  static constexpr int OFF_Q = OFF_P0 + 3 * P0_BLK;
  static constexpr int OFF_K = OFF_Q + NB * BLK;
  static constexpr int OFF_V = OFF_K + NB * BLK;
  static constexpr int END = OFF_V + NB * BLK;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + END;
  static_assert(6 * NB * 4096 <= 2 * NB * BLK, "output staging must fit in the dead Q/K region");
  static_assert(OFF_Q + (NB - 1) * BLK + 256 * 128 <= END, "tile-1 Q over-read must stay inside the allocation");
};

// MN-major operand with 128-byte swizzle (cute/atom/mma_traits_sm100.hpp, canonical Major-MN layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units): 64 contiguous MN elements per K row, rows 128 B apart, 8-row
// groups SBO = 1024 B apart, 64-element MN groups LBO bytes apart.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float atc_ex2(float x) {  // x <= 0 here: flush-to-zero of denormal results is harmless
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a pair of arguments x <= 0 on the FMA pipe (packed fp32 pairs), for the share of the softmax exponentials
// the MUFU unit (16 results / clk / SM) cannot take: x = n + f with n = rint(x) (magic-number rounding), |f| <= 0.5,
// 2^f by the degree-4 Taylor polynomial of e^(f ln 2) (relative error < 5e-5, far below the bf16 rounding of the
// probabilities), 2^n added into the exponent field.  Arguments below -125 (and the -inf of masked keys) give 2^-125:
// such keys meet zero rows of V and add nothing visible to the row sum.
__device__ __forceinline__ void atc_ex2_poly2(float x0, float x1, float& p0, float& p1) {
  typedef unsigned long long u64;
  auto pk = [](float lo, float hi) { return (u64)__float_as_uint(lo) | ((u64)__float_as_uint(hi) << 32); };
  auto fma2 = [](u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
  };
  auto add2 = [](u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
  };
  const u64 x = pk(fmaxf(x0, -125.f), fmaxf(x1, -125.f));
  const u64 t = add2(x, pk(12582912.f, 12582912.f));       // 1.5 * 2^23 + n in the low mantissa bits
  const u64 n = add2(t, pk(-12582912.f, -12582912.f));
  const u64 f = fma2(n, pk(-1.f, -1.f), x);
  u64 q = pk(9.618129e-3f, 9.618129e-3f);
  q = fma2(q, f, pk(5.550411e-2f, 5.550411e-2f));
  q = fma2(q, f, pk(2.402265e-1f, 2.402265e-1f));
  q = fma2(q, f, pk(6.931472e-1f, 6.931472e-1f));
  q = fma2(q, f, pk(1.f, 1.f));
  p0 = __uint_as_float((uint32_t)q + ((uint32_t)t << 23));
  p1 = __uint_as_float((uint32_t)(q >> 32) + ((uint32_t)(t >> 32) << 23));
}
__device__ __forceinline__ void atc_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// tmKV: QKV viewed [B][S][3d], box {64, 176, 1}; tmO: ATT viewed [B][S][d], box {64, 32, 1}.
template <int HD>
__global__ void __launch_bounds__(ATC_THREADS, 1)
    attn_tc_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmO, int S, int d,
                   int dbg, long long* trace, long long* ktime, const unsigned* rq, unsigned rq_target, unsigned* ra,
                   int stack_layers, int H, int B) {
  // stack_layers > 0 (stack form of the encoder, encoder.cu): a PERSISTENT grid of gridDim.x CTAs walks the units
  // g = (layer * B + b) * H + h of ALL layers in order, CTA c taking g = c, c + gridDim.x, ...; a unit of layer l >= 1
  // waits for rq[row tile] >= l * rq_target (the in_proj tiles the layer kernel stored for layer l), layer 0 reads the
  // plain in_proj GEMM's output (complete before this kernel starts: stream order).  Otherwise one unit per CTA,
  // grid (H, B).
  // rq / ra (layer-kernel form of the encoder, layer_chain.cuh): instead of waiting for the whole previous grid, the CTA
  // of sequence b waits until the in_proj tiles of the row tiles its tokens lie in have been stored (rq[row tile] >=
  // rq_target, bumped by the layer kernel's epilogue warps after their TMA stores completed), and announces its own
  // output with ra[b] += 1 per 32-row slab (ceil(S / 32) per CTA) once the slab's store has completed.  rq == null: plain programmatic dependency (layer 0).
  // debug only (tools/attn_trace.py): per-CTA clock64 stamps [grid][16]; null in the product path
#define ATC_TRACE(slot)                                                                                   \
  do {                                                                                                    \
    if (trace) trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64();           \
  } while (0)
  using C = AttnTcCfg<HD>;
  constexpr int NB = C::NB;
  extern __shared__ uint8_t atc_raw[];
  // 0: Q+K landed, 1: V landed, 2+t: scores of tile t, 4+t: P of tile t, 6+t: O of tile t, 8: unit drained (every
  // softmax warp has read its accumulators and its output stores have left shared memory: the next unit may load)
  __shared__ uint64_t bars[9];
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (smem_u32(atc_raw) + 1023u) & ~1023u;
  const uint32_t sP1 = base + C::OFF_P1, sP0 = base + C::OFF_P0, sQ = base + C::OFF_Q, sK = base + C::OFF_K,
                 sV = base + C::OFF_V;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = S > 128 ? 2 : 1;
  const bool stack = stack_layers > 0;
  const int g_first = stack ? (int)blockIdx.x : (int)(blockIdx.y * gridDim.x + blockIdx.x);
  const int g_step = stack ? (int)gridDim.x : 1;
  const int g_end = stack ? stack_layers * B * H : g_first + 1;
  const int heads = stack ? H : (int)gridDim.x;
  if (threadIdx.x == 0) {
    ATC_TRACE(0);
    ktime_entry(ktime);
    if (trace) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 14] = (long long)gt;
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 13] = smid;
    }
  }

  if (warp == 6) {
    if (lane == 0) {
      tma_prefetch_desc(&tmKV);
      tma_prefetch_desc(&tmO);
      mbar_init(&bars[0], 1), mbar_init(&bars[1], 1);
      mbar_init(&bars[2], 1), mbar_init(&bars[3], 1);
      mbar_init(&bars[4], 4), mbar_init(&bars[5], 2);
      mbar_init(&bars[6], 1), mbar_init(&bars[7], 1);
      mbar_init(&bars[8], ntile == 2 ? 6 : 4);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(&tmem_slot, ATC_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_launch_dependents();
  if (threadIdx.x == 0) ATC_TRACE(1);

  if (warp == 6) {
   int iter = 0;
   for (int g = g_first; g < g_end; g += g_step, ++iter) {
    const uint32_t ph = (uint32_t)(iter & 1);
    const int layer = g / (B > 0 && stack ? B * heads : 1 << 30), h = g % heads, b = (g / heads) % (stack ? B : 1 << 30);
    const unsigned unit_target = stack ? (unsigned)layer * rq_target : rq_target;
    if (iter > 0) {  // the previous unit has left shared memory and TMEM
      mbar_wait(&bars[8], ph ^ 1u);
      tc_fence_after();
    }
    // all lanes run the issue sequence (uniform descriptors / coordinates, see elect_one()); one elected lane issues
    if (elect_one()) {
      if (rq == nullptr) {
        pdl_wait();  // QKV is the previous kernel's output
      } else if (!stack || layer > 0) {
        const int m0 = (b * S) >> 8, m1 = (b * S + S - 1) >> 8;
        for (int m = m0; m <= m1; ++m) {
          unsigned v;
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(rq + m) : "memory");
          if ((int)(v - unit_target) < 0) {
            const long long t0 = clock64();
            do {
              __nanosleep(40);
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(rq + m) : "memory");
              if (clock64() - t0 > 4000000000LL) __trap();
            } while ((int)(v - unit_target) < 0);
          }
        }
        // the rows were written by TMA stores and are read by the TMA loads below: order the two proxies behind the acquire
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      if (iter == 0) ktime_ready(ktime);
      mbar_arrive_expect_tx(&bars[0], 2 * NB * C::BLK);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        tma_load_3d_u32(sQ + j * C::BLK, &tmKV, &bars[0], h * HD + j * 64, 0, b);
        tma_load_3d_u32(sK + j * C::BLK, &tmKV, &bars[0], d + h * HD + j * 64, 0, b);
      }
      mbar_arrive_expect_tx(&bars[1], NB * C::BLK);
#pragma unroll
      for (int j = 0; j < NB; ++j) tma_load_3d_u32(sV + j * C::BLK, &tmKV, &bars[1], 2 * d + h * HD + j * 64, 0, b);
    }
    __syncwarp();
    // ---- scores = Q K^T, one 128 x 176 accumulator per query tile ----
    constexpr uint32_t id_s = umma_idesc_bf16(128, ATC_KP);
    mbar_wait(&bars[0], ph);
    tc_fence_after();
    if (lane == 0) ATC_TRACE(2);
    for (int t = 0; t < ntile; ++t) {
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * C::BLK + (ks & 3) * 32;
          umma_bf16(tmem + t * ATC_KP, umma_desc_k_sw128(sQ + off + t * (128 * 128)), umma_desc_k_sw128(sK + off), id_s,
                    ks ? 1u : 0u);
        }
        umma_commit(&bars[2 + t]);
      }
      __syncwarp();
    }
    // ---- O = P V ----
    constexpr uint32_t id_o = umma_idesc_bf16_bmn(128, HD);
    mbar_wait(&bars[1], ph);
    if (lane == 0) ATC_TRACE(3);
    for (int t = 0; t < ntile; ++t) {
      mbar_wait(&bars[4 + t], ph);
      tc_fence_after();
      if (lane == 0) ATC_TRACE(4 + t);
      const uint32_t pb = t ? sP1 : sP0, pblk = t ? C::P1_BLK : C::P0_BLK;
      const uint32_t ocol = t ? 0u : (uint32_t)ATC_O0_COL;
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < ATC_KP / 16; ++ks) {
          umma_bf16(tmem + ocol, umma_desc_k_sw128(pb + (ks >> 2) * pblk + (ks & 3) * 32),
                    umma_desc_mn_sw128(sV + ks * (16 * 128), C::BLK), id_o, ks ? 1u : 0u);
        }
        umma_commit(&bars[6 + t]);
      }
      __syncwarp();
    }
   }
  } else {
    const int t = warp >> 2, lq = warp & 3;
    if (t < ntile) {
     int iter = 0;
     for (int g = g_first; g < g_end; g += g_step, ++iter) {
      const uint32_t ph = (uint32_t)(iter & 1);
      const int h = g % heads, b = (g / heads) % (stack ? B : 1 << 30);
      const int r = lq * 32 + lane;  // row inside the tile
      const uint32_t lane_sel = (uint32_t)(lq * 32) << 16;
      // ---- softmax over the 176 keys of this row ----
      mbar_wait(&bars[2 + t], ph);
      tc_fence_after();
      if (lane == 0 && lq == 0) ATC_TRACE(6 + t);
      // The row's 176 scores come out of TMEM with all six loads in flight at once (a chunk-by-chunk pipeline against the
      // maximum was slower: tcgen05.wait::ld waits for every outstanding load, so each chunk paid the full latency).
      // Row maximum on four independent chains.  Keys >= S are zero-filled by TMA and masked here; with S >= 160 (the
      // production shape) only the last 16 columns can be masked, which spares the other 160 a compare and a select.
      uint32_t v[ATC_KP];
      {
        const uint32_t ta = tmem + lane_sel + t * ATC_KP;
#pragma unroll
        for (int c = 0; c < 5; ++c) tmem_ld32(ta + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
        tmem_ld16(ta + 160, &v[160]);
        tc_wait_ld();
      }
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (S >= 160) {  // warp-uniform
#pragma unroll
        for (int k = 0; k < 160; ++k) mx4[k & 3] = fmaxf(mx4[k & 3], __uint_as_float(v[k]));
      } else {
#pragma unroll
        for (int k = 0; k < 160; ++k) {
          float sc = __uint_as_float(v[k]);
          if (k >= S) sc = -INFINITY;
          v[k] = __float_as_uint(sc);
          mx4[k & 3] = fmaxf(mx4[k & 3], sc);
        }
      }
#pragma unroll
      for (int k = 160; k < ATC_KP; ++k) {
        float sc = __uint_as_float(v[k]);
        if (k >= S) sc = -INFINITY;
        v[k] = __float_as_uint(sc);
        mx4[k & 3] = fmaxf(mx4[k & 3], sc);
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float c2 = rsqrtf((float)HD) * 1.4426950408889634f;
      const float mc = -mx * c2;
      float sum0 = 0.f, sum1 = 0.f;
      const uint32_t prow = (t ? sP1 : sP0) + r * 128;
      const uint32_t pblk = t ? C::P1_BLK : C::P0_BLK;
      // Experiment kept behind TAMF_ATTN_DBG=2: exp2 of 3 in 8 elements as a polynomial on the FMA pipe (atc_ex2_poly2),
      // because two softmax warps share sub-partitions 0 and 1 (tile 0 and tile 1 rows) and 176 MUFU results per row set
      // their pace.  It shortens the softmax phase of a CTA by ~0.4 k cycles (3.6 k -> 3.2 k) but neither the grid span
      // (two waves of 148 + 108 CTAs, 13.8 us) nor the in-chain time (20.5 us per launch) moves, so the exact-path MUFU
      // form stays the default (profiles/r01_exp_attn_softmax.txt).
      auto exp_pass = [&](auto poly) {
#pragma unroll
        for (int c8 = 0; c8 < ATC_KP / 8; ++c8) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a0 = fmaf(__uint_as_float(v[c8 * 8 + 2 * e]), c2, mc);
            const float a1 = fmaf(__uint_as_float(v[c8 * 8 + 2 * e + 1]), c2, mc);
            float p0, p1;
            if (decltype(poly)::value && (e == 3 || (e == 2 && (c8 & 1)))) {
              atc_ex2_poly2(a0, a1, p0, p1);
            } else {
              p0 = atc_ex2(a0), p1 = atc_ex2(a1);
            }
            sum0 += p0, sum1 += p1;
            pk[e] = pack_bf16x2(p0, p1);
          }
          atc_sts128(prow + (c8 >> 3) * pblk + (((c8 & 7) ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      };
      if (dbg & 2)  // experiment (TAMF_ATTN_DBG=2): 3 in 8 exponentials on the FMA pipe
        exp_pass(std::true_type{});
      else
        exp_pass(std::false_type{});
      const float inv = 1.0f / (sum0 + sum1);
      tc_fence_before();
      fence_proxy_async_smem();  // P is read by the tensor core through the async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[4 + t]);

      // ---- O * inv -> bf16 -> warp-private staging (dead Q/K region) -> TMA store ----
      mbar_wait(&bars[6 + t], ph);
      tc_fence_after();
      if (lane == 0 && lq == 0) ATC_TRACE(8 + t);
      const uint32_t oa = tmem + lane_sel + (t ? 0u : (uint32_t)ATC_O0_COL);
      const uint32_t stage = sQ + warp * (NB * 4096);
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        uint32_t o[64];
        tmem_ld32(oa + j * 64, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
        tmem_ld32(oa + j * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
        tc_wait_ld();
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            pk[e] = pack_bf16x2(__uint_as_float(o[c8 * 8 + 2 * e]) * inv, __uint_as_float(o[c8 * 8 + 2 * e + 1]) * inv);
          atc_sts128(stage + j * 4096 + lane * 128 + ((c8 ^ (lane & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      const int row0 = t * 128 + lq * 32;
      if (lane == 0 && row0 < S && !(dbg & 1)) {
#pragma unroll
        for (int j = 0; j < NB; ++j) tma_store_3d(&tmO, stage + j * 4096, h * HD + j * 64, row0, b);
        bulk_commit();
        if (ra) {
          // announce the 32 rows to the layer kernel: the store has completed (its writes are visible to this thread),
          // release at gpu scope.  One update per storing warp, ceil(S / 32) per CTA; the fence of an early warp runs
          // under the softmax / PV work of the later ones.
          bulk_wait<0>();
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(ra + b), "r"(1u) : "memory");
        } else {
          bulk_wait_read<0>();  // the staging tile must outlive the store's shared-memory reads; kernel end flushes the writes
        }
      }
      if (lane == 0 && lq == 0) ATC_TRACE(10 + t);
      // unit drained: accumulators read, output stores have left the staging tile (lane 0 waited for its bulk group)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[8]);
     }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    ATC_TRACE(12);
    ktime_exit(ktime);
    if (trace) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      trace[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 15] = (long long)gt;
    }
  }
  if (warp == 6) tmem_dealloc(tmem, ATC_TMEM_COLS);
}

struct AttnTcMaps {
  CUtensorMap kv, o;
};
// qkv bf16 [B*S, 3d] (packed in_proj output), out bf16 [B*S, d]
inline int make_attn_tc_maps(AttnTcMaps* m, const void* qkv, const void* out, int B, int S, int d) {
  int rc;
  const uint64_t ld = (uint64_t)3 * d * 2;
  if ((rc = make_tmap_3d_bf16(&m->kv, qkv, 3 * (uint64_t)d, S, B, ld, ld * S, ATC_KP))) return rc;
  if ((rc = make_tmap_3d_bf16(&m->o, out, d, S, B, (uint64_t)d * 2, (uint64_t)d * 2 * S, 32))) return rc;
  return TAMF_OK;
}

template <int HD>
int configure_attn_tc() {
  TAMF_CUDA_CHECK(cudaFuncSetAttribute(attn_tc_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       AttnTcCfg<HD>::SMEM_BYTES));
  return TAMF_OK;
}

template <int HD>
int launch_attn_tc(const AttnTcMaps& m, int B, int S, int H, int d, cudaStream_t stream, long long* trace = nullptr,
                   long long* ktime = nullptr, const unsigned* rq = nullptr, unsigned rq_target = 0, unsigned* ra = nullptr,
                   int stack_layers = 0, int stack_ctas = 0) {
  // stack_layers > 0: ONE persistent launch of `stack_ctas` CTAs (even: launched as clusters of 2 so that it occupies
  // whole TPCs next to the layer kernel's CTA pairs) for the attention of all layers; rq_target is then per layer.
  TAMF_REQUIRE(S <= ATC_KP, TAMF_E_BADARG, "attention: at most 176 tokens per sequence");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = stack_layers > 0 ? dim3(stack_ctas) : dim3(H, B);
  cfg.blockDim = dim3(ATC_THREADS);
  cfg.dynamicSmemBytes = AttnTcCfg<HD>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (stack_layers > 0) {
    TAMF_REQUIRE(stack_ctas >= 2 && stack_ctas % 2 == 0 && rq && ra, TAMF_E_BADARG, "attention (stack form): bad launch");
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
    ++na;
  } else if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  static const int dbg = getenv("TAMF_ATTN_DBG") ? atoi(getenv("TAMF_ATTN_DBG")) : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, attn_tc_kernel<HD>, m.kv, m.o, S, d, dbg, trace, ktime, rq, rq_target, ra,
                                     stack_layers, H, B);
  count_launch();
  if (e != cudaSuccess) {
    set_error(std::string("attention launch failed: ") + cudaGetErrorString(e));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

}  // namespace tamf
