// gemm.cuh -- persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   C[M,N] = A[M,K] (bf16, K-major) . W[N,K]^T (bf16, K-major = nn.Linear layout), fp32 accumulation in TMEM.
//
// CG = 2 (the hot-path configuration): a cluster of two CTAs on the two SMs of a TPC computes a 256 x BN tile with
// tcgen05.mma.cta_group::2 (M = 256).  Each CTA stages its own 128 rows of A and HALF of the W tile, so the
// shared-memory traffic per MMA drops from 12 KB to 8 KB per CTA -- with one CTA per tile the mainloop is bound by
// the 128 B/clk of shared memory (UMMA operand reads + TMA writes), not by the tensor pipe (measured).
// CG = 1 is the single-CTA form (M = 128), kept for the self-test and small problems.
//
// One CTA per SM, 576 threads:
//   warps 0-15 epilogue: tcgen05.ld the CTA's 128 x BN fp32 accumulator (warp w owns TMEM lanes 32*(w%4).., column
//              quarter w/4) and apply the fused epilogue.  TMEM hands every thread one accumulator ROW, so results
//              are transposed through a warp-private XOR-swizzled shared-memory tile and leave the SM as full
//              sectors; the fp32 residual of the LayerNorm epilogue arrives the same way (cp.async, one chunk
//              ahead).  16 warps (4 per scheduler): the epilogue is latency-, not issue-bound.
//   warp 16    TMA producer: cp.async.bulk.tensor 2D loads of the 128 x BK A tile and (BN/CG) x BK W tile into a
//              STAGES-deep shared-memory ring, signalled through mbarriers (expect_tx / complete_tx; with CG = 2 both
//              CTAs credit the leader's barrier).
//   warp 17    MMA issuer (leader CTA only when CG = 2): one thread issues tcgen05.mma.kind::f16 (N <= 256, K = 16) on
//              shared-memory descriptors; tcgen05.commit (multicast to both CTAs) releases ring slots and publishes
//              accumulators.
// The accumulator is double-buffered in TMEM when 2*BN <= 512 columns, so the epilogue of tile i overlaps the
// mainloop of tile i+1 (persistent static round-robin tile schedule).  Launched with programmatic dependent launch:
// barrier init, TMEM allocation, descriptor prefetch and the staging of bias/gamma/beta overlap the previous
// kernel's tail; griddepcontrol.wait precedes every access to activations.
//
// Epilogues (what the reference runs as separate ATen kernels, SURVEY.md 2.4 K1,K3-K6,K9):
//   EPI_BIAS_BF16      y = acc + b                                   -> bf16            (QKV in_proj)
//   EPI_BIAS_GELU_BF16 y = gelu_erf(acc + b)                         -> bf16            (linear1 + F.gelu)
//   EPI_BIAS_SILU_BF16 y = silu(acc + b)                             -> bf16            (input_merge.0, folded)
//   EPI_TOKEN_OUT      y = nan_to_num(acc + b) + pe[P0+tau]          -> token rows as a bf16 hi/lo pair (input_merge.2)
//   EPI_RES_LN         x = LayerNorm(x + acc + b) (eps 1e-5, biased var) in place -> bf16 hi/lo pair (BN == N == d)
// The residual stream lives in HBM as TWO bf16 planes, Xb = bf16(x) (which is also the next GEMM's A operand) and
// Xlo = bf16(x - Xb): x is recovered as Xb + Xlo to 2^-17 relative (fp32-grade for a LayerNorm input), but an LN
// epilogue stores 4 instead of 6 bytes per element -- its second pass is bound by the ~32 B/clk/SM store path.
//   EPI_POSTERIOR      x0 = nan_to_num(acc + b); x_{t-1} = c1 x0 + c2 x_t + sigma eps  -> [B,99,1,T] fp32
//   EPI_RESIDUAL_OUT   y = nan_to_num(x_in + acc + b) -> [B,T,99] fp32 (MF-MDM R: segment_refine_model.py:215-217)
//   EPI_F32            y = acc + b -> fp32 (self-test)
#pragma once
#include "common.cuh"

namespace tamf {

enum GemmEpi {
  EPI_BIAS_BF16 = 0,
  EPI_BIAS_GELU_BF16 = 1,
  EPI_BIAS_SILU_BF16 = 2,
  EPI_TOKEN_OUT = 3,
  EPI_RES_LN = 4,
  EPI_POSTERIOR = 5,
  EPI_F32 = 6,
  EPI_RESIDUAL_OUT = 7,
};

struct GemmParams {
  int M, N, K;
  const float* bias;  // [N] (may be null -> 0)
  // bf16 / fp32 row-major outputs
  __nv_bfloat16* out_bf16;
  int ld_bf16;
  float* out_f32;
  int ld_f32;
  // EPI_TOKEN_OUT: GEMM row m = b*T + tau  ->  token row b*S + P0 + tau
  const float* pe;  // [rows, N]
  int T, S, P0;
  // EPI_RES_LN: residual in / normalised out as the bf16 pair Xb (hi) + Xlo (lo), both [M,N].  EPI_TOKEN_OUT writes
  // the same pair.
  __nv_bfloat16* Xlo;
  __nv_bfloat16* Xb;
  const float* gamma;
  const float* beta;
  // EPI_RES_LN: rows of the result each TMEM lane quarter of a CTA owns (1..32, gemm_ln_rq()).  A CTA covers 4 * ln_rq
  // consecutive rows: quarter q computes rows [m0 + q ln_rq, +32) (A arrives as four 32-row boxes) and keeps the first
  // ln_rq of them.  32 = dense 128-row tiles; fewer rows spread the store-bound epilogue over more SMs (see gemm_ln_rq).
  int ln_rq;
  // EPI_POSTERIOR / EPI_RESIDUAL_OUT: GEMM row m = b*S + s (token); frame tau = s - P0
  const float* x_t;    // [B,nfeat,1,T] (POSTERIOR) | x_in [B,T,nfeat] (RESIDUAL_OUT)
  float* x_out;        // x_{t-1}, may alias x_t; null -> forward only
  float* x0_out;       // may be null
  const float* noise;  // null -> Philox
  const int* t_ptr;    // [B]
  const float *c1, *c2, *sigma;  // [num_steps] fp32
  unsigned long long seed;
  const unsigned long long* seed_ptr;  // non-null: the Philox seed is read from device memory (one graph for every chain)
  int nfeat;
  // host pointers to the TMA-store maps (copied into kernel parameters by launch_gemm):
  //   tmC: bf16 output [M,N], box {64 cols, 32 rows} (bias/GELU/SiLU epilogues) | Xb [M,N], box {32, ln_rq} (LN)
  //   tmX: Xlo [M,N] bf16, box {32, ln_rq}; the LN epilogue loads (residual) and stores (normalised) through tmC + tmX
  const CUtensorMap* tmC;
  const CUtensorMap* tmX;
  // debug only (tools/gemm_trace.py): per-CTA event timestamps, [grid][GEMM_TRACE_SLOTS] clock64 values; null in product
  long long* trace;
  long long* ktime;  // debug only: in-graph timing slots of this launch (common.cuh ktime_*); null in product
  int dbg;  // debug only: 1 skip global stores, 2 skip staging, 4 skip the whole epilogue body (bf16 epilogues),
            // 16 stop streaming W after the first ring fill (probe: the mainloop rate does not change, so TMA writes
            // into shared memory are not what bounds it)
};

constexpr int GEMM_TRACE_SLOTS = 64;
constexpr int GEMM_BM = 128;                           // accumulator rows per CTA (TMEM lanes)
constexpr int GEMM_EPI_WARPS = 16;
constexpr int GEMM_EPI_THREADS = GEMM_EPI_WARPS * 32;  // 512
constexpr int GEMM_THREADS = GEMM_EPI_THREADS + 64;    // + TMA producer warp + MMA issuer warp
constexpr int GEMM_CTRL_BYTES = 512;                   // mbarriers + TMEM slot | 16 x 2 residual-tile mbarriers (LN)
constexpr int GEMM_STG_WARP = 4096;                    // warp-private staging tile, 32 rows x 128 B (TMA-store source)
constexpr int GEMM_LN_STG_WARP = 8192;                 // LN: 2 x (2 KB hi + 2 KB lo) residual tiles in (pass 1) | the same out (pass 2)
constexpr int GEMM_SMEM_MAX = 232448;                  // 227 KB

constexpr bool epi_is_ln(int e) { return e == EPI_RES_LN; }
// bf16 [M,N] outputs that leave the SM as TMA stores of warp-private 32 x 64 staging tiles
constexpr bool epi_tma_bf16(int e) { return e == EPI_BIAS_BF16 || e == EPI_BIAS_GELU_BF16 || e == EPI_BIAS_SILU_BF16; }
constexpr bool epi_staged(int e) {
  return e == EPI_BIAS_BF16 || e == EPI_BIAS_GELU_BF16 || e == EPI_BIAS_SILU_BF16 || e == EPI_TOKEN_OUT;
}
// Tile geometry shared with the host code that encodes the tensor maps.
constexpr int gemm_bk(int bn, int cg) { return (bn == 512 && cg == 1) ? 32 : 64; }  // K extent of a pipeline stage
constexpr int gemm_un(int bn) { return bn > 256 ? 256 : bn; }                        // N of one MMA instruction
constexpr int gemm_b_box_rows(int bn, int cg) { return gemm_un(bn) / cg; }           // W rows per TMA box (per CTA)

template <int BN, int EPI, int CG>
struct GemmCfg {
  static constexpr int BK = gemm_bk(BN, CG);
  static constexpr int UN = gemm_un(BN);
  static constexpr int NH = BN / UN;            // MMA instructions per k-step
  static constexpr int BOX_B = UN / CG;         // W rows per box held by this CTA
  static constexpr int A_BYTES = GEMM_BM * BK * 2;
  static constexpr int B_BYTES = NH * BOX_B * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int ACC_STAGES = (2 * BN <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = (ACC_STAGES * BN <= 256) ? 256 : 512;
  // LN: bias | gamma | beta | row statistics sum[4][128], sumsq[4][128];  otherwise: the tile's bias slice
  // LN: bias | gamma | beta | row statistics;  otherwise the tile's bias slice (BN floats: a 5th ring stage of the
  // BN = 256 kernels depends on this staying at 1 KB)
  static constexpr int PARAM_BYTES = epi_is_ln(EPI) ? (3 * BN * 4 + 2 * 4 * 128 * 4) : (BN * 4);
  static constexpr int STG_BYTES = epi_staged(EPI) ? GEMM_EPI_WARPS * GEMM_STG_WARP : 0;
  static constexpr int RING_BUDGET = GEMM_SMEM_MAX - 1024 - GEMM_CTRL_BYTES - PARAM_BYTES - STG_BYTES;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  // Shared-memory map: [pipeline ring | staging | control | parameters].  LN epilogues stage through the drained
  // ring (one tile in flight); the other staged epilogues own a staging area (they overlap the next mainloop).
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + PIPE_BYTES + GEMM_CTRL_BYTES + PARAM_BYTES + STG_BYTES;
  static_assert(STAGES >= 2, "pipeline needs at least two stages");
  static_assert(PIPE_BYTES % 1024 == 0, "staging tiles behind the ring must stay 1024-byte aligned");
  static_assert(!epi_is_ln(EPI) || PIPE_BYTES >= GEMM_EPI_WARPS * GEMM_LN_STG_WARP, "LN staging must fit in the ring");
  static_assert(SMEM_BYTES <= GEMM_SMEM_MAX, "shared memory budget (227 KB) exceeded");
  static_assert((2 * STAGES + 3 * ACC_STAGES) * 8 + 8 <= 256, "control block too small");
};

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float fast_ex2(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// GELU with the exact (erf) form of F.gelu, the reference's activation="gelu":  x Phi(x),  Phi(x) = 0.5 (1 + erf(x / sqrt 2)).
// Phi(x) - 0.5 = x P(x^2) on |x| <= 4 (degree-8 weighted least-squares fit, |error of x Phi| <= 1.7e-5 in fp32 Horner form,
// i.e. below half a bf16 ulp of any |y| >= 0.01); outside, Phi is held at Phi(+-4) (1 - 3.2e-5 / 3.2e-5).  13 FMA/ALU
// instructions and no MUFU: the epilogue of linear1 must keep pace with a 128 x 256 x 512 MMA tile (4096 cycles for
// 32768 elements per SM), which two MUFU operations per element (16/clk/SM) alone would already exceed.
__device__ __forceinline__ float gelu_erf(float x) {
  const float xc = fminf(fmaxf(x, -4.0f), 4.0f);
  const float s = xc * xc;
  float p = 6.699732416e-11f;
  p = fmaf(p, s, -6.040797371e-09f);
  p = fmaf(p, s, 2.434135770e-07f);
  p = fmaf(p, s, -5.851926342e-06f);
  p = fmaf(p, s, 9.488355446e-05f);
  p = fmaf(p, s, -1.112714840e-03f);
  p = fmaf(p, s, 9.816041892e-03f);
  p = fmaf(p, s, -6.632534796e-02f);
  p = fmaf(p, s, 3.988829340e-01f);
  return x * fmaf(xc, p, 0.5f);
}
// x sigmoid(x) with the two MUFU approximations (ex2, rcp; ~2 ulp each) instead of expf and an IEEE division: the
// result is rounded to bf16 (8 bits) right after, and the exact form cost ~30 instructions per element.
// x -> -inf: ex2 -> +inf, rcp -> 0, x * 0 = -0 (the limit); x -> +inf: x * 1.
__device__ __forceinline__ float silu(float x) { return x * fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * x)); }

// Packed fp32 pairs (sm_100 FFMA2 / FADD2 / FMUL2: one issue slot for two lanes of fp32 math, same rounding as the
// scalar forms).  The epilogues are issue-bound, so every FMA chain that can run on pairs does.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float pk_lo(f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float pk_hi(f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// gelu_erf on a pair: identical polynomial and operation order as the scalar form (bit-identical results).
__device__ __forceinline__ f32x2 gelu_erf2(f32x2 x) {
  const f32x2 xc = pk2(fminf(fmaxf(pk_lo(x), -4.0f), 4.0f), fminf(fmaxf(pk_hi(x), -4.0f), 4.0f));
  const f32x2 s = mul2(xc, xc);
  f32x2 p = pk2(6.699732416e-11f, 6.699732416e-11f);
  p = fma2(p, s, pk2(-6.040797371e-09f, -6.040797371e-09f));
  p = fma2(p, s, pk2(2.434135770e-07f, 2.434135770e-07f));
  p = fma2(p, s, pk2(-5.851926342e-06f, -5.851926342e-06f));
  p = fma2(p, s, pk2(9.488355446e-05f, 9.488355446e-05f));
  p = fma2(p, s, pk2(-1.112714840e-03f, -1.112714840e-03f));
  p = fma2(p, s, pk2(9.816041892e-03f, 9.816041892e-03f));
  p = fma2(p, s, pk2(-6.632534796e-02f, -6.632534796e-02f));
  p = fma2(p, s, pk2(3.988829340e-01f, 3.988829340e-01f));
  return mul2(x, fma2(xc, p, pk2(0.5f, 0.5f)));
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Staging tiles hand data between the TMEM side (one thread = one accumulator row) and the global side (8 or 4
// lanes = one row segment, full sectors).  16-byte piece p of row r is XOR-swizzled so both sides are conflict-free.
__device__ __forceinline__ uint32_t stg128_off(int r, int p) { return (uint32_t)(r * 128 + ((p ^ (r & 7)) << 4)); }
__device__ __forceinline__ uint32_t stg64_off(int r, int p) { return (uint32_t)(r * 64 + ((p ^ ((r >> 1) & 3)) << 4)); }

__device__ __forceinline__ float bf16lo_f32(uint32_t w) { return __uint_as_float(w << 16); }          // element 0 of a pair
__device__ __forceinline__ float bf16hi_f32(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }  // element 1
// x -> (hi, lo) bf16 planes for a pair of values: hi = bf16(x), lo = bf16(x - hi)
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);
  lo = pack_bf16x2(a - bf16lo_f32(hi), b - bf16hi_f32(hi));
}

// K-major operand tile written by TMA with 64-byte swizzle (rows of 32 bf16): 8-row groups 512 B apart.
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(512 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)4 << 61);
}

#define GEMM_TRACE(slot)                                                                                    \
  do {                                                                                                      \
    if (p.trace && (slot) < GEMM_TRACE_SLOTS) p.trace[(size_t)blockIdx.x * GEMM_TRACE_SLOTS + (slot)] = clock64(); \
  } while (0)

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmX, const GemmParams p) {
  using Cfg = GemmCfg<BN, EPI, CG>;
  constexpr int STAGES = Cfg::STAGES, ACC = Cfg::ACC_STAGES, BK = Cfg::BK, UN = Cfg::UN, NH = Cfg::NH;
  constexpr bool LN = epi_is_ln(EPI);
  constexpr int PW = GEMM_EPI_WARPS, PT = GEMM_EPI_THREADS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;  // identical in both CTAs of a pair (same kernel, same layout)
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* ctrl = smem + Cfg::PIPE_BYTES + Cfg::STG_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);  // [STAGES]  (CG = 2: the leader's copy is the live one)
  uint64_t* empty_bar = full_bar + STAGES;                  // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;                 // [ACC]
  uint64_t* tempty_bar = tfull_bar + ACC;                   // [ACC]     (CG = 2: the leader's copy is the live one)
  uint64_t* ldone_bar = tempty_bar + ACC;                   // [ACC]     LN only: THIS CTA's epilogue has left the ring
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ldone_bar + ACC);
  uint64_t* rbar = reinterpret_cast<uint64_t*>(ctrl + 256);  // [16][2]  LN only: residual tile landed (per warp, per buffer)
  float* s_bias = reinterpret_cast<float*>(ctrl + GEMM_CTRL_BYTES);
  float* s_gamma = s_bias + BN;     // LN only
  float* s_beta = s_bias + 2 * BN;  // LN only
  float* s_sum = s_bias + 3 * BN;   // LN only: [4][128] partial row sums, then [4][128] partial sums of squares
  float* s_sq = s_sum + 4 * 128;
  uint8_t* s_stage = LN ? smem : (smem + Cfg::PIPE_BYTES);  // 1024-byte aligned (swizzled TMA tiles)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;  // 0 = leader
  // rows per CTA / per TMEM lane quarter: dense 128 / 32, except the LayerNorm tiles (GemmParams::ln_rq)
  const int RQ = LN ? p.ln_rq : 32;
  const int ROWS_CTA = 4 * RQ;
  const int tiles_m = (p.M + ROWS_CTA * CG - 1) / (ROWS_CTA * CG);
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + BK - 1) / BK;
  // Tile schedule: tiles are numbered n-major (tile = n * tiles_m + m) and every CTA pair takes a CONTIGUOUS run of
  // ceil(num_tiles / pairs) tiles, so consecutive tiles of a pair share the weight tile and the bias slice (restaged
  // only when n changes) -- same makespan as a strided schedule (252 / 336 tiles over 74 pairs: 4 / 5 tiles per pair).
  const int n_pairs = gridDim.x / CG;
  const int per_pair = (num_tiles + n_pairs - 1) / n_pairs;
  const int first_tile = (blockIdx.x / CG) * per_pair;
  const int last_tile = min(num_tiles, first_tile + per_pair);

  if (threadIdx.x == 0) {
    GEMM_TRACE(0);
    ktime_entry(p.ktime);
  }
  if (warp == PW && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);  // CG = 2: the leader's arrive.expect_tx covers the bytes of BOTH CTAs' loads
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PW * CG);  // one elected lane per epilogue warp (of both CTAs)
      mbar_init(&ldone_bar[a], PW);
    }
    if constexpr (LN) {
      for (int i = 0; i < 2 * PW; ++i) mbar_init(&rbar[i], 1);
    }
    if constexpr (LN || epi_tma_bf16(EPI)) {
      tma_prefetch_desc(&tmC);
      if constexpr (LN) tma_prefetch_desc(&tmX);
    }
    if constexpr (EPI == EPI_TOKEN_OUT) {
      if (p.tmC) tma_prefetch_desc(&tmC), tma_prefetch_desc(&tmX);
    }
    fence_mbar_init();
  }
  if (warp == PW + 1) {
    if constexpr (CG == 2) {
      tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  if constexpr (LN) {
    if (warp < PW) {  // weights only (constant over the chain): staged before the dependency wait
      for (int i = threadIdx.x; i < BN; i += PT) {
        s_bias[i] = p.bias ? p.bias[i] : 0.f;
        s_gamma[i] = p.gamma[i];
        s_beta[i] = p.beta[i];
      }
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) GEMM_TRACE(1);
  pdl_launch_dependents();  // the next kernel may run its own prologue on SMs this grid has already left
  pdl_wait();               // everything the previous kernel wrote is visible from here on
  if (threadIdx.x == 0) {
    GEMM_TRACE(2);
    ktime_ready(p.ktime);
  }

  if (warp == PW) {
    // ===================== TMA producer (both CTAs of a pair) =====================
    // all lanes run the loop (uniform addresses / coordinates); one elected lane issues
    {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < last_tile; ++tile, ++it) {
        const int m0 = (tile % tiles_m) * (ROWS_CTA * CG) + (int)rank * ROWS_CTA, n0 = (tile / tiles_m) * BN;
        if (lane == 0) GEMM_TRACE(8 + 2 * it);
        if constexpr (LN) {
          // the LN epilogue stages through this CTA's ring: do not refill it before that epilogue has drained
          mbar_wait(&ldone_bar[acc], acc_phase ^ 1u);
          if (++acc == ACC) acc = 0, acc_phase ^= 1u;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* a_dst = sA + stage * Cfg::A_BYTES;
          uint8_t* b_dst = sB + stage * Cfg::B_BYTES;
          if constexpr (CG == 2) {
            const uint32_t bar = mapa_cluster(smem_u32(&full_bar[stage]), 0);  // the leader's barrier
            const bool skip_w = (p.dbg & 16) && (it > 0 || kb >= STAGES);  // probe: W stays whatever the ring holds
            if (elect_one()) {
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], skip_w ? 2 * Cfg::A_BYTES : 2 * Cfg::STAGE_BYTES);
              if constexpr (LN) {  // four 32-row boxes, RQ rows apart: TMEM lane quarter q <- rows m0 + q RQ ...
#pragma unroll
                for (int q = 0; q < 4; ++q) tma_load_2d_2sm(a_dst + q * 32 * BK * 2, &tmA, bar, kb * BK, m0 + q * RQ);
              } else {
                tma_load_2d_2sm(a_dst, &tmA, bar, kb * BK, m0);
              }
#pragma unroll
              for (int h = 0; h < NH; ++h)
                if (!skip_w)
                  tma_load_2d_2sm(b_dst + h * Cfg::BOX_B * BK * 2, &tmB, bar, kb * BK, n0 + h * UN + (int)rank * Cfg::BOX_B);
            }
            // (no arrive from the peer: a remote release-arrive blocks ~1500 cycles per k-block; the peer's bytes are
            //  already accounted for by the leader's expect_tx, and its loads for the next use of a slot cannot be
            //  issued before the multicast commit that follows the completion of this phase)
          } else {
            if (elect_one()) {
              mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
              if constexpr (LN) {
#pragma unroll
                for (int q = 0; q < 4; ++q) tma_load_2d(a_dst + q * 32 * BK * 2, &tmA, &full_bar[stage], kb * BK, m0 + q * RQ);
              } else {
                tma_load_2d(a_dst, &tmA, &full_bar[stage], kb * BK, m0);
              }
#pragma unroll
              for (int h = 0; h < NH; ++h)
                tma_load_2d(b_dst + h * Cfg::BOX_B * BK * 2, &tmB, &full_bar[stage], kb * BK, n0 + h * UN);
            }
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        if (lane == 0) GEMM_TRACE(9 + 2 * it);
      }
    }
  } else if (warp == PW + 1) {
    // ===================== MMA issuer (leader CTA) =====================
    // all lanes run the loop so that the shared-memory descriptors are warp-uniform; one elected lane issues
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM * CG, UN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      int it = 0;
      for (int tile = first_tile; tile < last_tile; ++tile, ++it) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (lane == 0) {
            if (kb == 0) GEMM_TRACE(24 + 2 * it);
            if (it == 1 && kb < 8) GEMM_TRACE(56 + kb);
          }
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = (BK == 64) ? umma_desc_k_sw128(a_addr + k * 32) : umma_desc_k_sw64(a_addr + k * 32);
#pragma unroll
              for (int h = 0; h < NH; ++h) {
                const uint32_t bh = b_addr + h * Cfg::BOX_B * BK * 2 + k * 32;
                const uint64_t bd = (BK == 64) ? umma_desc_k_sw128(bh) : umma_desc_k_sw64(bh);
                if constexpr (CG == 2)
                  umma_bf16_2sm(d_tmem + h * UN, ad, bd, idesc, (kb | k) ? 1u : 0u);
                else
                  umma_bf16(d_tmem + h * UN, ad, bd, idesc, (kb | k) ? 1u : 0u);
              }
            }
            // ring slot reusable (in both CTAs) once these MMAs have read it; last k-block publishes the accumulator
            if constexpr (CG == 2) {
              umma_commit_2sm(&empty_bar[stage]);
              if (kb == num_kb - 1) umma_commit_2sm(&tfull_bar[acc]);
            } else {
              umma_commit(&empty_bar[stage]);
              if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        if (lane == 0) GEMM_TRACE(25 + 2 * it);
        if (++acc == ACC) acc = 0, acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 0..15, both CTAs) =====================
    const int lq = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter
    const int row_in_tile = lq * 32 + lane;
    constexpr int QW = BN / 4;       // columns per warp
    constexpr int CHUNKS = QW / 32;  // 32-column TMEM chunks per warp
    const uint32_t tempty_leader = (CG == 2) ? mapa_cluster(smem_u32(&tempty_bar[0]), 0) : 0u;
    uint32_t acc = 0, acc_phase = 0;
    uint32_t rpar = 0;  // LN: parity bits of the two residual-tile mbarriers of this warp
    int staged_n0 = -1;
    int it = 0;
    for (int tile = first_tile; tile < last_tile; ++tile, ++it) {
      const int m0 = (tile % tiles_m) * (ROWS_CTA * CG) + (int)rank * ROWS_CTA, n0 = (tile / tiles_m) * BN;
      const int row = m0 + lq * RQ + lane;  // dense tiles: RQ = 32
      const bool row_ok = row < p.M;
      const int grow0 = m0 + lq * RQ;  // first global row of this warp
      if constexpr (epi_tma_bf16(EPI)) {
        if (n0 != staged_n0) {  // the 64 bias values of this column quarter, shared by its 4 warps (named barrier of
                                // 128 threads instead of a CTA-wide one between tiles)
          float* wb = s_bias + cq * QW;
          const int c = n0 + cq * QW + lane;
          asm volatile("bar.sync %0, 128;" ::"r"(2 + cq) : "memory");  // the quarter has left the previous slice
          if (lq == 0) {
            wb[lane] = (p.bias && c < p.N) ? p.bias[c] : 0.f;
            wb[lane + 32] = (p.bias && c + 32 < p.N) ? p.bias[c + 32] : 0.f;
          }
          asm volatile("bar.sync %0, 128;" ::"r"(2 + cq) : "memory");
          staged_n0 = n0;
        }
      } else if constexpr (!LN) {
        if (n0 != staged_n0) {  // (re)stage the tile's bias slice; all 16 warps walk the same tile sequence
          asm volatile("bar.sync 1, 512;" ::: "memory");
          for (int i = threadIdx.x; i < BN; i += PT) s_bias[i] = (p.bias && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
          asm volatile("bar.sync 1, 512;" ::: "memory");
          staged_n0 = n0;
        }
      }
      // EPI_POSTERIOR: x_t and the step's coefficients do not depend on the accumulator -- fetch them before waiting
      // for it, so their latency hides behind the mainloop.  (Inside the store loop the compiler must keep every
      // x_t load behind the preceding x_out store, x_out may alias x_t: 25 serialised L2 round trips per thread.)
      [[maybe_unused]] float xt[32];
      [[maybe_unused]] float post_k1 = 0.f, post_k2 = 0.f, post_sg = 0.f;
      [[maybe_unused]] int pt = 0;
      [[maybe_unused]] unsigned long long post_seed = 0;
      if constexpr (EPI == EPI_POSTERIOR) {
        post_seed = p.seed_ptr ? *p.seed_ptr : p.seed;
        static_assert(EPI != EPI_POSTERIOR || CHUNKS == 1, "the posterior epilogue handles one 32-column chunk per warp");
        const int c0 = n0 + cq * QW;
        const int b = row / p.S, s = row % p.S;
#pragma unroll
        for (int j = 0; j < 32; ++j) xt[j] = 0.f;
        if (row_ok && s >= p.P0 && c0 < p.nfeat) {
          pt = p.t_ptr[b];
          if (p.x_out) {
            const float* xr = p.x_t + ((size_t)b * p.nfeat + c0) * p.T + (s - p.P0);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c0 + j < p.nfeat) xt[j] = xr[(size_t)j * p.T];
          }
          post_k1 = p.c1[pt], post_k2 = p.c2[pt], post_sg = p.sigma[pt];
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (threadIdx.x == 0) GEMM_TRACE(40 + 2 * it);
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(lq * 32) << 16) + cq * QW;

      // hand the drained accumulator stage back to the MMA warp
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2)
            mbar_arrive_cluster(tempty_leader + acc * 8);
          else
            mbar_arrive(&tempty_bar[acc]);
        }
      };
      if constexpr (LN) {
        // ---------------- x = LayerNorm(x + acc + b), two passes over TMEM ----------------
        // Lanes >= RQ carry rows that belong to the next quarter / CTA: they run the arithmetic on whatever the
        // staging buffers hold and nothing of theirs is stored (the TMA boxes are RQ rows tall).
        // Warp-private 8 KB region of the drained ring: pass 1 receives the residual as TMA tiles [RQ rows x 32 cols]
        // of the two bf16 planes (two buffers, two chunks in flight); pass 2 stages the normalised hi / lo tiles
        // (two sets) and hands them to TMA stores, so no LSU global access is left in the epilogue.
        const uint32_t wst = smem_u32(s_stage) + warp * GEMM_LN_STG_WARP;
        uint64_t* rb = rbar + warp * 2;
        const int ccol0 = cq * QW;  // first column of this warp's quarter
        auto prefetch = [&](int ck) {  // lane 0: residual tiles of chunk ck -> buffer ck & 1 (rows past M arrive as zeros)
          mbar_arrive_expect_tx(&rb[ck & 1], (uint32_t)RQ * 128u);
          tma_load_2d_u32(wst + (ck & 1) * 4096, &tmC, smem_u32(&rb[ck & 1]), ccol0 + ck * 32, grow0);
          tma_load_2d_u32(wst + (ck & 1) * 4096 + 2048, &tmX, smem_u32(&rb[ck & 1]), ccol0 + ck * 32, grow0);
        };
        if (elect_one()) {  // TMA issue on uniform operands (see elect_one())
          prefetch(0);
          if (CHUNKS > 1) prefetch(1);
        }
        // ---- pass 1: y = acc + bias + residual -> back to TMEM; row sum and sum of squares ----
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          uint32_t v[32];
          tmem_ld32(taddr + ck * 32, v);
          mbar_wait(&rb[ck & 1], (rpar >> (ck & 1)) & 1u);
          rpar ^= 1u << (ck & 1);
          tc_wait_ld();
          const int c0 = ccol0 + ck * 32;
          const uint32_t src = wst + (ck & 1) * 4096;
#pragma unroll
          for (int j = 0; j < 4; ++j) {  // 8 columns per 16-byte piece of each plane
            const uint4 h4 = lds128(src + stg64_off(lane, j));
            const uint4 l4 = lds128(src + 2048 + stg64_off(lane, j));
            const float4 ba = *reinterpret_cast<const float4*>(s_bias + c0 + j * 8);
            const float4 bb = *reinterpret_cast<const float4*>(s_bias + c0 + j * 8 + 4);
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
          tmem_st32(taddr + ck * 32, v);
          __syncwarp();  // every lane has read this residual buffer: it may be refilled
          if (ck + 2 < CHUNKS && elect_one()) prefetch(ck + 2);
        }
        tc_wait_st();
        if (threadIdx.x == 0) GEMM_TRACE(4);
        s_sum[cq * 128 + row_in_tile] = (s0 + s1) + (s2 + s3);
        s_sq[cq * 128 + row_in_tile] = (q0 + q1) + (q2 + q3);
        asm volatile("bar.sync 1, 512;" ::: "memory");
        if (threadIdx.x == 0) GEMM_TRACE(5);
        const float inv_n = 1.0f / (float)p.N;
        const float mean = ((s_sum[row_in_tile] + s_sum[128 + row_in_tile]) +
                            (s_sum[256 + row_in_tile] + s_sum[384 + row_in_tile])) * inv_n;
        const float ex2 = ((s_sq[row_in_tile] + s_sq[128 + row_in_tile]) +
                           (s_sq[256 + row_in_tile] + s_sq[384 + row_in_tile])) * inv_n;
        const float var = fmaxf(ex2 - mean * mean, 0.f);  // biased variance (F.layer_norm), fp32
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        // ---- pass 2: normalise + affine -> staging set ck & 1 -> TMA stores of the hi / lo bf16 tiles ----
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          uint32_t v[32];
          tmem_ld32(taddr + ck * 32, v);
          tc_wait_ld();
          if (ck == CHUNKS - 1) release_acc();  // last TMEM read of this tile
          if (ck >= 2) {  // the stores of chunk ck - 2 must have finished reading this staging set
            if (elect_one()) bulk_wait_read<1>();
            __syncwarp();
          }
          const int c0 = ccol0 + ck * 32;
          const uint32_t s_outh = wst + (ck & 1) * 4096, s_outl = s_outh + 2048;
          if (p.dbg & 2) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = j * 8 + e * 2;
              const float2 g2 = *reinterpret_cast<const float2*>(s_gamma + c0 + c);
              const float2 be2 = *reinterpret_cast<const float2*>(s_beta + c0 + c);
              const float y0 = fmaf(fmaf(__uint_as_float(v[c]), rstd, nmr), g2.x, be2.x);
              const float y1 = fmaf(fmaf(__uint_as_float(v[c + 1]), rstd, nmr), g2.y, be2.y);
              split_bf16x2(y0, y1, hi[e], lo[e]);
            }
            sts128(s_outh + stg64_off(lane, j), make_uint4(hi[0], hi[1], hi[2], hi[3]));
            sts128(s_outl + stg64_off(lane, j), make_uint4(lo[0], lo[1], lo[2], lo[3]));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (!(p.dbg & 1) && elect_one()) {
            if (!(p.dbg & 8)) tma_store_2d(&tmC, s_outh, c0, grow0);  // rows past M are clipped by the tensor map
            if (!(p.dbg & 4)) tma_store_2d(&tmX, s_outl, c0, grow0);
            bulk_commit();
          }
        }
        if (elect_one()) {
          bulk_wait_read<0>();          // staging lives in the ring: the producer may refill it only now
          mbar_arrive(&ldone_bar[acc]);
        }
        if (threadIdx.x == 0) GEMM_TRACE(6);
      } else if constexpr (EPI == EPI_TOKEN_OUT) {
        // ---- y = nan_to_num(acc + b) + pe[P0+tau] -> the bf16 pair Xb / Xlo at token row b*S + P0 + tau ----
        const uint32_t wst = smem_u32(s_stage) + warp * GEMM_STG_WARP;
        const int prow = lane >> 2, pc = lane & 3;  // global side: 8 rows x 4 pieces (16 fp32 columns) per pass
        // Fast path (T a multiple of 32, tmC / tmX given): the 32 rows of a warp are 32 consecutive frames of ONE
        // sequence.  The positional-encoding tile [32 frames x 32 cols] fp32 is read with full-line loads (8 lanes = one
        // 128-byte row segment) and turned to thread = row through the warp's staging tile; the hi / lo bf16 tiles of
        // the result are staged like the LayerNorm output and leave as TMA stores through the [B][T][N] view of the
        // frame tokens (make_token_out_maps).  Same arithmetic as the row-piece path below (bit-identical results);
        // that path read pe with one line per lane and row (L1 tag-rate bound, 17 k cycles per tile).
        const bool tok_tma = p.tmC != nullptr && (p.T & 31) == 0;  // grid-uniform
        if (tok_tma) {
          const int wb0 = grow0 / p.T, wtau0 = grow0 - wb0 * p.T;  // warp-uniform
          const float* pe_w = p.pe + (size_t)(p.P0 + wtau0) * p.N;
          const int gr4 = lane >> 3, gpc = lane & 7;
#pragma unroll 1
          for (int ck = 0; ck < CHUNKS; ++ck) {
            const int cl = cq * QW + ck * 32, cg = n0 + cl;
            if (cg >= p.N || grow0 >= p.M) break;  // warp-uniform
            uint32_t v[32];
            tmem_ld32(taddr + ck * 32, v);
            float4 g[8];
#pragma unroll
            for (int k = 0; k < 8; ++k)
              g[k] = *reinterpret_cast<const float4*>(pe_w + (size_t)(k * 4 + gr4) * p.N + cg + gpc * 4);
            if (elect_one()) bulk_wait_read<0>();  // the previous stores have finished reading the staging tiles
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 8; ++k)
              sts128(wst + stg128_off(k * 4 + gr4, gpc), make_uint4(__float_as_uint(g[k].x), __float_as_uint(g[k].y),
                                                                     __float_as_uint(g[k].z), __float_as_uint(g[k].w)));
            __syncwarp();
            uint4 q[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = lds128(wst + stg128_off(lane, j));
            __syncwarp();  // every lane holds its pe row: the tile may be overwritten
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 ba = *reinterpret_cast<const float4*>(s_bias + cl + j * 8);
              const float4 bb = *reinterpret_cast<const float4*>(s_bias + cl + j * 8 + 4);
              const float y0 = nan_to_num(__uint_as_float(v[8 * j]) + ba.x) + __uint_as_float(q[2 * j].x);
              const float y1 = nan_to_num(__uint_as_float(v[8 * j + 1]) + ba.y) + __uint_as_float(q[2 * j].y);
              const float y2 = nan_to_num(__uint_as_float(v[8 * j + 2]) + ba.z) + __uint_as_float(q[2 * j].z);
              const float y3 = nan_to_num(__uint_as_float(v[8 * j + 3]) + ba.w) + __uint_as_float(q[2 * j].w);
              const float y4 = nan_to_num(__uint_as_float(v[8 * j + 4]) + bb.x) + __uint_as_float(q[2 * j + 1].x);
              const float y5 = nan_to_num(__uint_as_float(v[8 * j + 5]) + bb.y) + __uint_as_float(q[2 * j + 1].y);
              const float y6 = nan_to_num(__uint_as_float(v[8 * j + 6]) + bb.z) + __uint_as_float(q[2 * j + 1].z);
              const float y7 = nan_to_num(__uint_as_float(v[8 * j + 7]) + bb.w) + __uint_as_float(q[2 * j + 1].w);
              uint32_t hi[4], lo[4];
              split_bf16x2(y0, y1, hi[0], lo[0]);
              split_bf16x2(y2, y3, hi[1], lo[1]);
              split_bf16x2(y4, y5, hi[2], lo[2]);
              split_bf16x2(y6, y7, hi[3], lo[3]);
              sts128(wst + stg64_off(lane, j), make_uint4(hi[0], hi[1], hi[2], hi[3]));
              sts128(wst + 2048 + stg64_off(lane, j), make_uint4(lo[0], lo[1], lo[2], lo[3]));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
              tma_store_3d(&tmC, wst, cg, wtau0, wb0);
              tma_store_3d(&tmX, wst + 2048, cg, wtau0, wb0);
              bulk_commit();
            }
          }
        } else {
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          const int cl = cq * QW + ck * 32, cg = n0 + cl;
          if (cg >= p.N) break;  // warp-uniform
          uint32_t v[32];
          tmem_ld32(taddr + ck * 32, v);
          // positional-encoding values of this lane's 8 (row, 4-column) pieces of the chunk, fetched up front: inside
          // the store loop every load would have to stay behind the previous iteration's stores (possible aliasing),
          // i.e. 8 serialised L2 round trips per chunk
          float4 pe4[2][4];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int gr = grow0 + i * 8 + prow;
              pe4[hh][i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (gr < p.M) {
                const int tau = gr % p.T;
                pe4[hh][i] = *reinterpret_cast<const float4*>(p.pe + (size_t)(p.P0 + tau) * p.N + cg + hh * 16 + pc * 4);
              }
            }
          }
          tc_wait_ld();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {  // 16 columns (64 B per row) per staging round
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cl + hh * 16 + j * 4);
              sts128(wst + stg64_off(lane, j),
                     make_uint4(__float_as_uint(nan_to_num(__uint_as_float(v[hh * 16 + 4 * j]) + b4.x)),
                                __float_as_uint(nan_to_num(__uint_as_float(v[hh * 16 + 4 * j + 1]) + b4.y)),
                                __float_as_uint(nan_to_num(__uint_as_float(v[hh * 16 + 4 * j + 2]) + b4.z)),
                                __float_as_uint(nan_to_num(__uint_as_float(v[hh * 16 + 4 * j + 3]) + b4.w))));
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = i * 8 + prow, gr = grow0 + r;
              if (gr < p.M) {
                const uint4 o = lds128(wst + stg64_off(r, pc));
                const int b = gr / p.T, tau = gr - b * p.T;
                const int col = cg + hh * 16 + pc * 4;
                const float4 q4 = pe4[hh][i];
                const float y0 = __uint_as_float(o.x) + q4.x, y1 = __uint_as_float(o.y) + q4.y;
                const float y2 = __uint_as_float(o.z) + q4.z, y3 = __uint_as_float(o.w) + q4.w;
                const size_t orow = (size_t)b * p.S + p.P0 + tau;
                uint32_t h0, h1, l0, l1;
                split_bf16x2(y0, y1, h0, l0);
                split_bf16x2(y2, y3, h1, l1);
                *reinterpret_cast<uint2*>(p.Xb + orow * p.N + col) = make_uint2(h0, h1);
                *reinterpret_cast<uint2*>(p.Xlo + orow * p.N + col) = make_uint2(l0, l1);
              }
            }
            __syncwarp();
          }
        }
        }
      } else if constexpr (epi_tma_bf16(EPI)) {
        // ---- bias (+ activation) -> bf16: the warp's 32 x 64 slab is staged as one 128-byte-swizzled tile and
        //      leaves the SM as a single TMA store; TMEM is released as soon as the accumulator is in registers ----
        static_assert(BN == 256, "the TMA-store bf16 epilogue is written for 64-column warp slabs (BN = 256)");
        const uint32_t wst = smem_u32(s_stage) + warp * GEMM_STG_WARP;
        const int cl = cq * QW;
        const float* wb = s_bias + cl;
        // The TMEM read port (16 B/clk per sub-partition) and the math would otherwise alternate: all four warps of a
        // sub-partition wait for their loads, then all compute.  Load the second 32 columns while the first are
        // processed, so that one warp's loads overlap another's arithmetic.
        uint32_t v0[32], v1[32];
        tmem_ld32(taddr, v0);
        tc_wait_ld_dep(v0);
        tmem_ld32(taddr + 32, v1);
        if (elect_one()) bulk_wait_read<0>();  // the previous tile's store has finished reading the staging tile
        __syncwarp();
        const bool live = n0 + cl < p.N;  // warp-uniform (N is a multiple of 64)
        auto half = [&](const uint32_t (&vv)[32], int j0) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            const uint32_t* v = vv + jj * 8;
            uint32_t o[4];
            if constexpr (EPI == EPI_BIAS_GELU_BF16) {
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
                float y0 = __uint_as_float(v[2 * e]) + b2.x, y1 = __uint_as_float(v[2 * e + 1]) + b2.y;
                if constexpr (EPI == EPI_BIAS_SILU_BF16) y0 = silu(y0), y1 = silu(y1);
                o[e] = pack_bf16x2(y0, y1);
              }
            }
            sts128(wst + stg128_off(lane, j), make_uint4(o[0], o[1], o[2], o[3]));
          }
        };
        if (live) half(v0, 0);
        tc_wait_ld_dep(v1);
        release_acc();  // the whole accumulator slab of this warp is in registers
        if (live) {
          half(v1, 4);
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one()) {
            tma_store_2d(&tmC, wst, n0 + cl, grow0);  // rows past M / columns past N are clipped
            bulk_commit();
          }
        }
      } else {
        // ---- small-N / remapped epilogues: row-per-thread stores ----
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          uint32_t v[32];
          __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
          tmem_ld32(taddr + ck * 32, v);
          tc_wait_ld();
          const int cl = cq * QW + ck * 32;  // column within the tile
          const int c0 = n0 + cl;            // global column of v[0]
          if constexpr (EPI == EPI_POSTERIOR) {
            const int b = row / p.S, s = row % p.S;
            if (row_ok && s >= p.P0 && c0 < p.nfeat) {
              const int tau = s - p.P0;
              const int t = pt;
              const float k1 = post_k1, k2 = post_k2, sg = post_sg;
              const unsigned long long frame = (unsigned long long)b * p.T + tau;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float eps[4] = {0.f, 0.f, 0.f, 0.f};
                if (p.x_out && !p.noise && c0 + j < p.nfeat)
                  philox_normal4(post_seed, (uint32_t)t, frame, (uint32_t)((c0 + j) >> 2), eps);
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                  const int f = c0 + j + e4;
                  if (f < p.nfeat) {
                    const size_t e = ((size_t)b * p.nfeat + f) * p.T + tau;
                    const float x0 = nan_to_num(__uint_as_float(v[j + e4]) + s_bias[cl + j + e4]);
                    if (p.x0_out) p.x0_out[e] = x0;
                    if (p.x_out) {
                      const float n = p.noise ? p.noise[e] : eps[e4];
                      p.x_out[e] = (k1 * x0 + k2 * xt[j + e4]) + sg * n;
                    }
                  }
                }
              }
            }
          } else if constexpr (EPI == EPI_RESIDUAL_OUT) {
            const int b = row / p.S, s = row % p.S;
            if (row_ok && s >= p.P0 && c0 < p.nfeat) {
              const size_t base = ((size_t)b * p.T + (s - p.P0)) * p.nfeat;  // x_in / out are [B,T,nfeat]
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int f = c0 + j;
                if (f < p.nfeat)
                  p.x_out[base + f] = nan_to_num(p.x_t[base + f] + (__uint_as_float(v[j]) + s_bias[cl + j]));
              }
            }
          } else {
            static_assert(EPI == EPI_F32 || EPI == EPI_POSTERIOR || EPI == EPI_RESIDUAL_OUT, "unhandled epilogue");
            if (row_ok && c0 < p.N) {
              float* o = p.out_f32 + (size_t)row * p.ld_f32 + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(
                    __uint_as_float(v[j]) + s_bias[cl + j], __uint_as_float(v[j + 1]) + s_bias[cl + j + 1],
                    __uint_as_float(v[j + 2]) + s_bias[cl + j + 2], __uint_as_float(v[j + 3]) + s_bias[cl + j + 3]);
            }
          }
        }
      }
      // accumulator stage drained -> hand it back to the MMA warp (LN and the TMA-store epilogues did so already)
      if constexpr (!LN && !epi_tma_bf16(EPI)) release_acc();
      if (threadIdx.x == 0) GEMM_TRACE(41 + 2 * it);
      if (++acc == ACC) acc = 0, acc_phase ^= 1u;
    }
    if constexpr (LN || epi_tma_bf16(EPI) || EPI == EPI_TOKEN_OUT) {
      if (elect_one()) bulk_wait<0>();  // this thread's TMA stores have been performed before the CTA retires
    }
    (void)rpar;
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();  // the peer may still signal our barriers / TMEM
  if (threadIdx.x == 0) {
    GEMM_TRACE(3);
    ktime_exit(p.ktime);
  }
  if (warp == PW + 1) {
    tc_fence_after();
    if constexpr (CG == 2)
      tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    else
      tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// host-side launcher ------------------------------------------------------------------------------
int num_sms();
bool pdl_enabled();  // TAMF_PDL=0 turns programmatic dependent launch off (debug aid)

template <int BN, int EPI, int CG>
int configure_gemm() {  // once per process, outside any stream capture
  TAMF_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GemmCfg<BN, EPI, CG>::SMEM_BYTES));
  return TAMF_OK;
}

// Rows per TMEM lane quarter of the LayerNorm tiles for an [M, N] result.  The LN epilogue cannot overlap a mainloop
// (the accumulator fills TMEM) and its second pass is bound by the ~32 B/clk one SM can store, so dense 128-row tiles
// (M = 10 560: 83 CTAs) leave 64 SMs idle while the others push 256 KB each.  With RQ = ceil(M / (4 SMs)) = 18 rows per
// quarter every SM takes part: the MMA still computes 128 rows per CTA (mainloop time unchanged, the extra rows are
// discarded) and the stores per SM shrink by RQ / 32.  Measured (profiles/r01_exp_ln_rq_ab.txt): the epilogue drops
// from 17.3 k to 13.2 k cycles as predicted, but the 1.78 x redundant tensor work on all 148 SMs pushes the board into
// its power cap (986 W, SM clock 1965 -> 1837 MHz) and the chain slows from 69.0 to 66.7 sequences/s; RQ = 24 is a
// wash (69.0, clock 1935 MHz).  So dense tiles (32) stay the default and TAMF_LN_RQ selects the others.
inline int gemm_ln_rq(int M) {
  static const int forced = getenv("TAMF_LN_RQ") ? atoi(getenv("TAMF_LN_RQ")) : 32;
  int rq = forced > 0 ? forced : (M + 4 * num_sms() - 1) / (4 * num_sms());  // 0: one wave over all SMs
  return rq < 8 ? 8 : (rq > 32 ? 32 : rq);
}

// EPI_TOKEN_OUT store maps: the frame tokens of the residual planes Xb / Xlo ([B][S][N] bf16, frames at rows P0 ..
// P0 + T - 1 of every sequence) viewed as [B][T][N] with box {32 cols, 32 frames, 1}.  The kernel takes this path when
// T is a multiple of 32: a warp's 32 GEMM rows (b*T + tau) are then exactly one box of one sequence.
inline int make_token_out_maps(CUtensorMap* hi, CUtensorMap* lo, const void* Xb, const void* Xlo, int B, int T, int S,
                               int P0, int N) {
  int rc;
  const uint64_t row = (uint64_t)N * 2;
  if ((rc = make_tmap_3d_bf16(hi, (const char*)Xb + (size_t)P0 * row, N, T, B, row, row * S, 32, 32))) return rc;
  return make_tmap_3d_bf16(lo, (const char*)Xlo + (size_t)P0 * row, N, T, B, row, row * S, 32, 32);
}

// tmA: A [M,K] with box {gemm_bk(BN,CG), 128} (EPI_RES_LN: box {.., 32}); tmB: W [N,K] with box {gemm_bk(BN,CG),
// gemm_b_box_rows(BN,CG)}.
template <int BN, int EPI, int CG>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  if (EPI == EPI_RES_LN) {
    TAMF_REQUIRE(p.ln_rq >= 1 && p.ln_rq <= 32, TAMF_E_BADARG, "gemm: LayerNorm tiles need 1 <= ln_rq <= 32");
  }
  const int rows_pair = (EPI == EPI_RES_LN ? 4 * p.ln_rq : GEMM_BM) * CG;
  const int tiles = ((p.M + rows_pair - 1) / rows_pair) * ((p.N + BN - 1) / BN);
  const int slots = num_sms() / CG;
  const int grid = (tiles < slots ? tiles : slots) * CG;
  if (epi_staged(EPI) || EPI == EPI_F32) {
    TAMF_REQUIRE(p.N % 32 == 0, TAMF_E_BADARG, "gemm: N must be a multiple of 32 for this epilogue");
  }
  if (EPI == EPI_RES_LN) {
    TAMF_REQUIRE(p.N == BN, TAMF_E_BADARG, "gemm: the LayerNorm epilogue needs the whole row in one tile (N == BN)");
  }
  TAMF_REQUIRE(p.K % 8 == 0, TAMF_E_BADARG, "gemm: K must be a multiple of 8 (16-byte TMA rows)");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GemmCfg<BN, EPI, CG>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CG == 2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2, attr[na].val.clusterDim.y = 1, attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  if (epi_is_ln(EPI)) {
    TAMF_REQUIRE(p.tmC && p.tmX, TAMF_E_BADARG, "gemm: the LayerNorm epilogue needs the Xb / X tensor maps");
  } else if (epi_tma_bf16(EPI)) {
    TAMF_REQUIRE(p.tmC, TAMF_E_BADARG, "gemm: the bf16 epilogues need the output tensor map");
    TAMF_REQUIRE(p.N % 64 == 0, TAMF_E_BADARG, "gemm: N must be a multiple of 64 for the TMA-store epilogue");
  }
  const CUtensorMap& tmC = p.tmC ? *p.tmC : tmA;  // unused maps are passed as copies of tmA
  const CUtensorMap& tmX = p.tmX ? *p.tmX : tmA;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI, CG>, tmA, tmB, tmC, tmX, p);
  count_launch();
  if (e != cudaSuccess) {
    set_error(std::string("gemm launch failed: ") + cudaGetErrorString(e));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

}  // namespace tamf
