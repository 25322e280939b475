// layer_chain.cuh -- everything between two attention kernels of the encoder stack in ONE persistent tcgen05 kernel:
//
//   LN1   X = LayerNorm(X + ATT . Wo^T + b)          out_proj + residual + norm1      (eps 1e-5, biased variance)
//   L1    H = gelu(Xb . W1^T + b)                    linear1 + exact-erf GELU
//   LN2   X = LayerNorm(X + H . W2^T + b)            linear2 + residual + norm2
//   INP   QKV = Xb . Win'^T + b'                     in_proj of the NEXT layer (absent after the last layer)
//
// Reference semantics: nn.TransformerEncoderLayer, post-norm (interaction_segment_mdm.py:63-70; segment_refine_model.py
// :88-95).  Round 1 ran these as four kernels.  M = 10 560 token rows are 41.25 row tiles of 256: an N = 512 GEMM has 84
// pair tiles for 74 CTA pairs (two rounds, the second 14 % full), and a LayerNorm epilogue that needs the whole 512-wide
// row filled TMEM (no overlap with a mainloop) while 84 SMs carried all its stores: 49 k cycles for 18 k of balanced
// tensor work (profiles/r01_gemm_cta_timelines.txt).  Here
//   * every unit of work is a 256 x 256 tile of a CTA pair (tcgen05.mma.cta_group::2, M = 256, N = 256), accumulators
//     double-buffered in TMEM, so the epilogue of unit i overlaps the mainloop of unit i + 1 for EVERY epilogue;
//   * the residual is added BY THE TENSOR CORE: after the K loop of a LayerNorm unit, four more ring stages multiply the
//     unit's own 256 x 256 block of the two residual planes (Xb, Xlo) with a 64 x 64 identity that stays in shared
//     memory (M = 256, N = 64, K = 64 per 64-column block and plane: products with 1.0 are exact, the accumulation is
//     fp32), so the epilogue never loads the residual: no small-box TMA loads (one row per ~2 cycles of the SM's TMA unit
//     whatever its length -- they delayed the mainloop's own loads), no staging traffic;
//   * a 512-wide LayerNorm row is computed by TWO pairs (column halves, pairs 2k and 2k + 1): each keeps its 128 x 256
//     slab of y = acc + b in REGISTERS (64 per thread), the halves swap per-row (sum, sum of squares) through L2 as one
//     64-bit word per row that carries the data AND is its own flag (relaxed 64-bit accesses, nothing else to order: a
//     gpu-scope fence costs 1.5 - 3 k cycles), all 148 SMs store;
//   * the four stages form a dependency chain PER ROW TILE (LN1 -> L1 -> LN2 -> INP) and row tiles are independent, so a
//     host-built static schedule (list scheduling of the DAG on a two-resource model of a pair: tensor pipe + epilogue
//     warps) interleaves the units of different row tiles: the ~10 k cycles between the last MMA of a LayerNorm unit and
//     the moment its rows are in L2 are filled with other row tiles' work instead of idling every SM twice per layer.
//     A unit waits for a per-row-tile counter.  Publishing: an epilogue warp's elected lane waits for ITS TMA stores
//     (cp.async.bulk.wait_group: the writes are then visible to that thread) and arrives on a CTA-local mbarrier; the
//     signal warp collects the 16 arrivals and performs ONE gpu-scope release (fence.acq_rel.gpu + counter update).
//     Consuming: the scout warp polls with ld.acquire.gpu, issues fence.proxy.async (the data is read by TMA = async
//     proxy) and clears the unit for the TMA producer warp through shared memory.  (A relaxed update without the fence
//     let 1 evaluation in ~270 read stale rows at the production shape: profiles/r02_exp_release_signals.txt.)
//     Every pair's list is a subsequence of ONE global topological order and all CTAs of the
//     grid are co-resident (grid <= SM count, 1 CTA / SM), so the schedule cannot deadlock; every spin is bounded (trap).
//   * the same counters replace the grid-wide dependency between this kernel and the attention kernels on both sides: an LN1
//     unit waits for the attention CTAs of the sequences overlapping its row tile, an attention CTA for the in_proj tiles
//     of its sequence (attn_tc.cuh) -- the ragged tail of one kernel overlaps the head of the next;
//   * sync state: a statistics word is reset by its single reader; the counters are cleared once per denoiser evaluation
//     and launch number l waits for l x target.
#pragma once
#include <algorithm>
#include <map>
#include <mutex>
#include <type_traits>
#include <vector>

#include "gemm.cuh"

namespace tamf {

enum ChainKind { CK_LN1 = 0, CK_L1 = 1, CK_LN2 = 2, CK_INP = 3 };

// Unit code: layer << 24 | kind << 28 is NOT used -- layout: bits 28-29 kind, 24-27 layer (stack form, else 0), 8-23 row
// tile, 0-7 column tile.
__host__ __device__ inline int chain_code(int layer, int kind, int m, int n) { return (kind << 28) | (layer << 24) | (m << 8) | n; }
__host__ __device__ inline int chain_kind(int code) { return (code >> 28) & 3; }
__host__ __device__ inline int chain_layer(int code) { return (code >> 24) & 15; }
__host__ __device__ inline int chain_m(int code) { return (code >> 8) & 0xFFFF; }
__host__ __device__ inline int chain_n(int code) { return code & 0xFF; }
constexpr int CH_MAX_STACK_LAYERS = 16;

struct LayerWeightPtrs {
  const float* bias[4];             // as LayerParams::bias
  const float *gamma[2], *beta[2];  // norm1, norm2
};

struct LayerParams {
  int M, d, ff, n_inp;                // n_inp = 3 d, or 0 when there is no next layer
  const float* bias[4];               // LN1: out_proj bias [d] | L1: linear1 bias [ff] | LN2: linear2 bias [d] | INP: [3d]
  const float *gamma[2], *beta[2];    // norm1, norm2 [d]
  const int* sched_off;               // [pairs + 1]
  const int* sched;                   // unit codes: kind << 28 | row tile << 8 | column tile
  // counters [6][tiles_m] then [B]: 0 LN1 high plane stored | 1 LN1 low plane stored | 2 H tiles stored | 3 LN2 high plane
  // stored | 4 LN2 low plane stored | 5 QKV tiles stored (each counts epilogue warps) | rA[b]: 32-row slabs of sequence
  // b the attention kernel has stored.  All are cleared once per denoiser evaluation (a memset node ahead of the first kernel) and
  // only grow afterwards: launch number `launch_idx` (1-based within the evaluation) waits for launch_idx x target.
  unsigned* ctr;
  int tiles_m;
  int launch_idx;                     // 1-based index of this layer kernel within the evaluation
  int B, S, heads;                    // sequences, tokens per sequence, attention CTAs per sequence
  unsigned target_ln, target_h;       // halves * 32 warps | (ff / 256) * 32 warps
  // statistics words [2 LN][tiles_m][halves][2 ranks][4 reader copies][128 rows]: (sum | sum of squares << 32) over the
  // CTA's 256 columns of a row; all-ones = not posted (the word is its own flag; its reader resets it)
  unsigned long long* stats;
  long long* trace;                   // debug only: [grid][GEMM_TRACE_SLOTS] clock64 stamps
  long long* ktime;                   // debug only: in-graph timing slots of this launch (common.cuh ktime_*)
  unsigned target_att;  // updates of rA[b] per layer: heads x ceil(S / 32) (attn_tc.cuh, one per stored 32-row slab)
  // Stack form (ONE launch for all layers, encoder.cu): unit codes carry a layer index; the weights of layer l come from
  // these device arrays instead of the kernel's own tensor maps / the pointers above (null in the per-layer form).
  const CUtensorMap* wmaps;     // [L][4]: Wo, W1, W2, Win(next layer) -- 64-byte aligned, written by the host at upload
  const LayerWeightPtrs* lw;    // [L]
  int ooo;   // > 0: run-time unit selection -- the leader's scout emits any READY unit among the next `ooo` of the pair's
             // list (LayerNorm units in list order); 0 (default): the list order, waiting for each unit in turn.
             // EXPERIMENTAL (TAMF_CHAIN_OOO): parity-green, but 4-6 % slower than the list order in both forms -- the
             // scan + the relay to the peer CTA sit on every unit's critical path, the scout cannot clear units ahead of
             // time and the residual-first phase is lost (profiles/r02_exp_stack_form.txt)
  int dbg;
  int grid_wait;  // debug: wait for the whole previous grid instead of relying on the per-unit dependencies alone
};

// 20 warps = 5 warpgroups: 16 epilogue warps, then one warpgroup with the TMA producer (16), the MMA issuer (17), the
// dependency scout (18) and the signal warp (19); register reallocation (setmaxnreg) always involves whole warpgroups.
constexpr int CH_THREADS = GEMM_EPI_THREADS + 128;
constexpr int CH_STAGES = 5;
constexpr int CH_BN = 256;
constexpr int CH_A_BYTES = GEMM_BM * 64 * 2;       // 128 rows x 64 k
constexpr int CH_B_BYTES = (CH_BN / 2) * 64 * 2;   // this CTA's half of the 256 W rows
constexpr int CH_STAGE_BYTES = CH_A_BYTES + CH_B_BYTES;
constexpr int CH_PIPE_BYTES = CH_STAGES * CH_STAGE_BYTES;
constexpr int CH_STG_WARP = 2048;                 // warp-private staging tile: 32 rows x 32 bf16 columns (64-byte rows, 64 B swizzle);
                                                  // a warp's 32 x 64 slab leaves as TWO TMA stores, which frees the 32 KB of a fifth ring stage
constexpr int CH_STG_BYTES = GEMM_EPI_WARPS * CH_STG_WARP;
constexpr int CH_IDENT_BYTES = 32 * 128;  // this CTA's 32 rows of the 64 x 64 bf16 identity (K-major, 128-byte swizzle)
constexpr int CH_SIG_BARS = 16;           // output-signal barriers of a CTA in flight (a warp runs < 2 units = 4 signals ahead)
constexpr int CH_RES_KB = 4;              // residual ring stages of a LayerNorm unit: 2 planes x 2 stages, each holding
                                          // TWO 64-column blocks (A slot, B slot) -- the ring is latency bound per stage
constexpr int CH_CTRL_BYTES = 1024;
// [2 LN][bias | gamma | beta] (256 floats each) | per-warp bias slice of the running L1 / INP tile [16][64] | row
// statistics [2 buffers][sum, sq][4 column quarters][128 rows]
constexpr int CH_PARAM_BYTES = 6 * CH_BN * 4 + GEMM_EPI_WARPS * 64 * 4 + 2 * 2 * 4 * 128 * 4;
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

// Bounded mbarrier wait without printf (a call site keeps many registers alive around it; the trap alone reports the
// protocol bug as a CUDA error).  All lanes re-issue try_wait.  Tried against the board's power cap and rejected: one
// waiting lane with a 20 us suspend-time hint + warp sync (5 % slower: the parked lane wakes late); __nanosleep(100)
// between the polls of the epilogue and signal warps (no change in time, clock or power).
__device__ __forceinline__ void mbar_wait_q(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// Lane 0 polls `*flag >= target` (modulo 2^32; acquire, gpu scope) with a bounded spin; the warp leaves together.
__device__ __forceinline__ void chain_wait_ge(const unsigned* flag, unsigned target) {
  if (lane_id() == 0) {
    if ((int)(ld_acquire_gpu(flag) - target) < 0) {
      const long long t0 = clock64();
      while ((int)(ld_acquire_gpu(flag) - target) < 0) {
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
// per-unit stamps: units 0..7 of a pair own slots 4..51 (6 each); slots 52..63 are the special stamps
#define CHAIN_TRACE_UNIT(k, it)                 \
  do {                                          \
    if ((it) < 8) CHAIN_TRACE(4 + (k) + 6 * (it)); \
  } while (0)

// Tensor maps.  tmATT / tmH / tmXh / tmXl: A operands [M, K] with box {64, 128} (tmXh = Xb is the A operand of L1 and INP
// AND the high residual plane, tmXl = Xlo the low one); tmWo / tmW1 / tmW2 / tmWin: weights [N, K], box {64, 128};
// tmHst / tmQst: bf16 outputs H / QKV, box {32, 32} (64 B swizzle); tmXh_st / tmXl_st: the residual planes, box {32, 32};
// tmI: the 64 x 64 identity, box {64, 32}.
__global__ void __launch_bounds__(CH_THREADS, 1)
    layer_chain_kernel(const __grid_constant__ CUtensorMap tmATT, const __grid_constant__ CUtensorMap tmWo,
                       const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
                       const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmHst,
                       const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW2,
                       const __grid_constant__ CUtensorMap tmWin, const __grid_constant__ CUtensorMap tmQst,
                       const __grid_constant__ CUtensorMap tmXh_st, const __grid_constant__ CUtensorMap tmXl_st,
                       const __grid_constant__ CUtensorMap tmI, const LayerParams p) {
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
  uint64_t* sig_bar = reinterpret_cast<uint64_t*>(s_dep + 1);  // [CH_SIG_BARS] output signals owed by this CTA (signal warp)
  // Order of execution: s_order[k & 15] = code of the k-th unit this pair runs, written by the scout BEFORE it publishes
  // 2 k + 1 / 2 k + 2 in s_dep (every other warp reads the unit from here).  s_prog[r]: units whose signals the signal
  // warp of CTA r has published (ring space); s_prod: units the producer has finished loading (late binding);
  // s_inbox: (rank 1, run-time selection) the leader scout's choices, one self-validating 64-bit word per ring entry.
  int* s_order = reinterpret_cast<int*>(sig_bar + CH_SIG_BARS);
  uint32_t* s_prog = reinterpret_cast<uint32_t*>(s_order + 16);
  uint32_t* s_prod = s_prog + 2;
  unsigned long long* s_inbox = reinterpret_cast<unsigned long long*>(s_prod + 2);  // [16] (rank 1, run-time selection)
  float* s_ln = reinterpret_cast<float*>(ctrl + CH_CTRL_BYTES);  // [2 LN][bias | gamma | beta][256]
  float* s_bias2 = s_ln + 6 * BN;           // [16 warps][64]
  float* s_stat = s_bias2 + PW * 64;        // [2 buffers][sum, sq][4 column quarters][128 rows]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader of the pair
  const int pair = blockIdx.x >> 1;
  const int halves = p.d / BN;
  const int tiles_m = p.tiles_m;
  const int u_begin = p.sched_off[pair], u_end = p.sched_off[pair + 1];  // host data (written at bind time)
  const int n_units = u_end - u_begin;
  // every warp but the scout takes the k-th unit from s_order once the scout has published it (phase 1: the residual
  // planes of a LayerNorm unit may be loaded; phase 2: everything may)
  auto dep_seen = [&]() {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(s_dep)) : "memory");
    return v;
  };
  auto wait_published = [&](uint32_t need, unsigned sleep_ns) {
    if (dep_seen() >= need) return;
    const long long t0 = clock64();
    while (dep_seen() < need) {
      if (sleep_ns) __nanosleep(sleep_ns);
      if (clock64() - t0 > 4000000000LL) __trap();
    }
  };
  auto unit_code = [&](int k) { return *reinterpret_cast<volatile int*>(s_order + (k & 15)); };

  if (threadIdx.x == 0) {
    CHAIN_TRACE(0);
    ktime_entry(p.ktime);
  }
  if (warp == PW && lane == 0) {
    tma_prefetch_desc(&tmATT);
    tma_prefetch_desc(&tmWo);
    tma_prefetch_desc(&tmXh);
    tma_prefetch_desc(&tmXl);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmHst);
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmXh_st);
    tma_prefetch_desc(&tmXl_st);
    tma_prefetch_desc(&tmI);
    if (p.n_inp > 0) {
      tma_prefetch_desc(&tmWin);
      tma_prefetch_desc(&tmQst);
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
    for (int i = 0; i < CH_SIG_BARS; ++i) mbar_init(&sig_bar[i], PW);  // one elected lane per epilogue warp of this CTA
    *s_dep = 0u;
    s_prog[0] = 0u, s_prog[1] = 0u, *s_prod = 0u;
    for (int i = 0; i < 16; ++i) s_inbox[i] = 0ull;
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
  if (p.grid_wait) pdl_wait();
  // NO grid-wide dependency wait: every unit waits for exactly the rows it reads (scout warp below).  The first units
  // (LN1) wait for the attention CTAs of the sequences that overlap their row tile, so this kernel starts on the SMs the
  // attention kernel has left while its last CTAs are still running; everything older is ordered transitively (an
  // attention CTA started only after the in_proj tiles of its sequence, those after LN2 of their row tiles, ...).
  if (threadIdx.x == 0) {
    CHAIN_TRACE(2);
    CHAIN_TRACE_NS(56);  // globaltimer (ns) when the dependency wait ended: the common time base across SMs
    ktime_ready(p.ktime);
  }

  // Registers: the block is allocated 20 warps x 96.  The producer / MMA / scout warpgroup gives up 64 per thread
  // (128 x 64 = 8192 go back to the CTA's pool), the four epilogue warpgroups take 16 more each (512 x 16 = 8192): the
  // LayerNorm epilogue keeps 64 accumulator values per thread over two passes.  Each setmaxnreg dominates exactly one
  // role's code, so ptxas allocates every role under its own limit.
  if (warp >= PW) {
   asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
   if (warp == PW) {
    // ===================== TMA producer (both CTAs of the pair) =====================
    uint32_t stage = 0, phase = 0;
    int it = 0;
    const bool stalls = p.trace && (p.dbg & 16);  // debug: accumulate wait cycles (tools/stack_stalls.py, slots 64..)
    long long st_dep = 0, st_empty = 0;
    const uint32_t full_leader = mapa_cluster(smem_u32(&full_bar[0]), 0);
    for (; it < n_units; ++it) {
      {
        const long long t0 = stalls ? clock64() : 0;
        wait_published((uint32_t)(2 * it + 1), 100);  // (sleeping: a busy spin on all SMs is paid for in clock)
        if (stalls) st_dep += clock64() - t0;
      }
      const int code = unit_code(it);
      const int kind = chain_kind(code), m = chain_m(code), n = chain_n(code), ul = chain_layer(code);
      const int m0 = m * 256 + (int)rank * GEMM_BM, n0 = n * BN;
      const CUtensorMap* ta = kind == CK_LN1 ? &tmATT : (kind == CK_LN2 ? &tmH : &tmXh);
      const CUtensorMap* tb = p.wmaps ? p.wmaps + ul * 4 + kind
                                      : (kind == CK_LN1 ? &tmWo : (kind == CK_L1 ? &tmW1 : (kind == CK_LN2 ? &tmW2 : &tmWin)));
      const int num_kb = (kind == CK_LN2 ? p.ff : p.d) / 64;
      const int res_kb = (kind & 1) ? 0 : CH_RES_KB;  // LayerNorm units: the residual stages come FIRST
      const int total_kb = num_kb + res_kb;
      auto wait_dep = [&](uint32_t need) { wait_published(need, 100); };
      if (!res_kb) {
        const long long t0 = stalls ? clock64() : 0;
        wait_dep((uint32_t)(2 * it + 2));
        if (stalls) st_dep += clock64() - t0;
      }
      if (lane == 0) CHAIN_TRACE_UNIT(0, it);
      for (int kk = 0; kk < total_kb; ++kk) {
        if (res_kb && kk == res_kb) {
          const long long t0 = stalls ? clock64() : 0;
          wait_dep((uint32_t)(2 * it + 2));
          if (stalls) st_dep += clock64() - t0;
        }
        {
          const long long t0 = stalls ? clock64() : 0;
          mbar_wait_q(&empty_bar[stage], phase ^ 1u);
          if (stalls) st_empty += clock64() - t0;
        }
        uint8_t* a_dst = sA + stage * CH_A_BYTES;
        uint8_t* b_dst = sB + stage * CH_B_BYTES;
        const uint32_t bar = full_leader + stage * 8;
        if (kk >= res_kb) {
          const int kb = kk - res_kb;
          if (elect_one()) {
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * CH_STAGE_BYTES);
            tma_load_2d_2sm(a_dst, ta, bar, kb * 64, m0);
            tma_load_2d_2sm(b_dst, tb, bar, kb * 64, n0 + (int)rank * (BN / 2));
          }
        } else {  // residual stage: 128 rows x two 64-column blocks of plane (rb / 2) at columns n0 + 128 (rb % 2)
          const int rb = kk;
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
      if (lane == 0) {
        CHAIN_TRACE_UNIT(1, it);
        asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(s_prod)), "r"((uint32_t)(it + 1)) : "memory");
      }
    }
    if (stalls && lane == 0) {
      long long* z = p.trace + (size_t)gridDim.x * GEMM_TRACE_SLOTS + (size_t)blockIdx.x * 8;
      z[0] = st_dep, z[1] = st_empty, z[7] = it;
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
      const bool stalls = p.trace && (p.dbg & 16);
      long long st_full = 0, st_tempty = 0;
      for (; it < n_units; ++it) {
        wait_published((uint32_t)(2 * it + 1), 0);
        const int code = unit_code(it);
        const int kind = chain_kind(code);
        const int num_kb = (kind == CK_LN2 ? p.ff : p.d) / 64;
        const int res_kb = (kind & 1) ? 0 : CH_RES_KB;
        const int total_kb = num_kb + res_kb;
        if (!(kind & 1) && !ident_ready) {
          mbar_wait_q(ident_bar, 0);
          ident_ready = true;
        }
        {
          const long long t0 = stalls ? clock64() : 0;
          mbar_wait_q(&tempty_bar[acc], acc_phase ^ 1u);
          if (stalls) st_tempty += clock64() - t0;
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kk = 0; kk < total_kb; ++kk) {
          {
            const long long t0 = stalls ? clock64() : 0;
            mbar_wait_q(&full_bar[stage], phase);
            if (stalls) st_full += clock64() - t0;
          }
          tc_fence_after();
          if (lane == 0 && kk == 0) CHAIN_TRACE_UNIT(2, it);
          const uint32_t a_addr = smem_u32(sA + stage * CH_A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * CH_B_BYTES);
          if (kk >= res_kb) {
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                              (kk | k) ? 1u : 0u);
              umma_commit_2sm(&empty_bar[stage]);  // ring slot reusable in both CTAs once these MMAs have read it
              if (kk == total_kb - 1) umma_commit_2sm(&tfull_bar[acc]);
            }
          } else {
            // residual stages open the accumulator: acc[:, 64 j .. 64 j + 63] (+)= X_plane block j . I64 for the stage's
            // two blocks; the high plane (stages 0, 1) initialises its columns, the low plane (2, 3) accumulates
            const uint32_t dj = d_tmem + 128u * (uint32_t)(kk & 1);
            const bool first_plane = kk < 2;
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(dj, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(i_addr + k * 32), idesc_res,
                              (first_plane && k == 0) ? 0u : 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_bf16_2sm(dj + 64u, umma_desc_k_sw128(b_addr + k * 32), umma_desc_k_sw128(i_addr + k * 32), idesc_res,
                              (first_plane && k == 0) ? 0u : 1u);
              umma_commit_2sm(&empty_bar[stage]);
            }
          }
          __syncwarp();
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        if (lane == 0) CHAIN_TRACE_UNIT(3, it);
        if (++acc == 2) acc = 0, acc_phase ^= 1u;
      }
      if (stalls && lane == 0) {
        long long* z = p.trace + (size_t)gridDim.x * GEMM_TRACE_SLOTS + (size_t)blockIdx.x * 8;
        z[2] = st_full, z[3] = st_tempty;
      }
    }
   } else if (warp == PW + 2) {
    // ===================== dependency scout =====================
    // Walks the unit list ahead of the producer: waits until the inputs of the next unit are in L2 (acquire on the row
    // tile's counters) and publishes the number of cleared units in shared memory.  The producer only reads that word:
    // polling and fencing in the producer itself drained the ring at every unit boundary (+2 k cycles per tile,
    // measured).  No proxy fence: a counter is bumped only after the TMA stores have COMPLETED in L2, which is where the
    // producer's TMA loads read (no L1 in that path, nothing can be stale).
    const unsigned* r_att = p.ctr + 6 * tiles_m;
    // ring space: entry k & 15 of s_order may be rewritten once both CTAs' signal warps are done with unit k - 16
    auto ring_space = [&](int k) {
      if (lane == 0) {
        const long long t0 = clock64();
        for (;;) {
          uint32_t a, b;
          asm volatile("ld.relaxed.cta.shared::cta.u32 %0, [%1];" : "=r"(a) : "r"(smem_u32(&s_prog[0])) : "memory");
          asm volatile("ld.relaxed.cta.shared::cta.u32 %0, [%1];" : "=r"(b) : "r"(smem_u32(&s_prog[1])) : "memory");
          if (k - (int)min(a, b) < 12) break;
          __nanosleep(100);
          if (clock64() - t0 > 4000000000LL) __trap();
        }
      }
      __syncwarp();
    };
    // the inputs were written through the async proxy (TMA stores) and will be read through it (TMA loads of the
    // producer warp): the acquire of the poll orders generic accesses only
    auto publish = [&](uint32_t v) {
      if (lane == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(s_dep)), "r"(v) : "memory");
      }
      __syncwarp();
    };
    auto set_order = [&](int k, int code) {
      if (lane == 0) *reinterpret_cast<volatile int*>(s_order + (k & 15)) = code;
      __syncwarp();
    };
    // non-blocking: are all inputs of the unit in L2?  (one thread; acquire loads at gpu scope)
    auto unit_ready = [&](int code) -> bool {
      const int kind = chain_kind(code), m = chain_m(code);
      const unsigned li = (unsigned)(p.launch_idx + chain_layer(code));
      const unsigned t_ln = li * p.target_ln, t_h = li * p.target_h;
      auto ge = [&](const unsigned* c, unsigned target) { return (int)(ld_acquire_gpu(c) - target) >= 0; };
      if (kind == CK_LN1) {
        if (!ge(p.ctr + 4 * tiles_m + m, (li - 1u) * p.target_ln)) return false;
        const int b0 = (m * 256) / p.S, b1 = min(p.B - 1, (m * 256 + 255) / p.S);
        for (int b = b0; b <= b1; ++b)
          if (!ge(r_att + b, li * p.target_att)) return false;
        return true;
      }
      if (kind == CK_L1) return ge(p.ctr + 0 * tiles_m + m, t_ln);
      if (kind == CK_LN2) return ge(p.ctr + 1 * tiles_m + m, t_ln) && ge(p.ctr + 2 * tiles_m + m, t_h);
      return ge(p.ctr + 3 * tiles_m + m, t_ln);
    };
    if (p.ooo <= 0) {
      // ---- list order: wait for each unit in turn (both CTAs of the pair walk the same list on their own) ----
      for (int it = 0; it < n_units; ++it) {
        const int code = p.sched[u_begin + it];
        const int kind = chain_kind(code), m = chain_m(code);
        const unsigned li = (unsigned)(p.launch_idx + chain_layer(code));  // 1-based layer kernel index of the evaluation
        const unsigned t_ln = li * p.target_ln, t_h = li * p.target_h;
        ring_space(it);
        set_order(it, code);
        if (kind == CK_LN1) {
          // residual: the previous layer's LN2 planes of the row tile (low plane counter: a CTA announces it after both
          // of its stores).  In the first layer the planes come from the embed kernels, ordered only through the
          // attention kernel's grid-wide wait, so there the attention counters are waited for first.
          const int b0 = (m * 256) / p.S, b1 = min(p.B - 1, (m * 256 + 255) / p.S);
          if (li == 1u)
            for (int b = b0; b <= b1; ++b) chain_wait_ge(r_att + b, li * p.target_att);
          chain_wait_ge(p.ctr + 4 * tiles_m + m, (li - 1u) * p.target_ln);
          publish((uint32_t)(2 * it + 1));
          // the attention output of every sequence that overlaps the row tile
          for (int b = b0; b <= b1; ++b) chain_wait_ge(r_att + b, li * p.target_att);
        } else if (kind == CK_L1) {
          chain_wait_ge(p.ctr + 0 * tiles_m + m, t_ln);  // LN1 high plane of the row tile
        } else if (kind == CK_LN2) {
          chain_wait_ge(p.ctr + 1 * tiles_m + m, t_ln);  // LN1 low plane (the residual of this unit)
          publish((uint32_t)(2 * it + 1));
          chain_wait_ge(p.ctr + 2 * tiles_m + m, t_h);   // every H tile of the row tile
        } else {
          chain_wait_ge(p.ctr + 3 * tiles_m + m, t_ln);  // LN2 high plane
        }
        if (kind == CK_L1 && lane == 0 && p.trace && p.trace[(size_t)blockIdx.x * GEMM_TRACE_SLOTS + 54] == 0)
          CHAIN_TRACE_NS(54);
        publish((uint32_t)(2 * it + 2));
      }
    } else if (rank == 0) {
      // ---- run-time selection (leader CTA decides for the pair): among the next p.ooo un-emitted units of the list,
      // plus the earliest un-emitted LayerNorm unit (LayerNorm units keep their list order: both pairs of a duo hold the
      // same LayerNorm sequence, and a half must never wait for a partner whose window does not reach it), emit the first
      // one whose inputs are all in L2.  Late binding: at most two units ahead of the TMA producer.  The choice goes to
      // the peer CTA as ONE 64-bit word (unit number + 1 | code) per ring entry, which its scout re-validates. ----
      int head = 0, emitted = 0;
      uint32_t mask = 0;  // bit k: unit head + k already emitted
      long long t_idle = clock64();
      while (emitted < n_units) {
        ring_space(emitted);
        uint32_t prod = 0;
        if (lane == 0) asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(prod) : "r"(smem_u32(s_prod)) : "memory");
        prod = __shfl_sync(0xffffffffu, prod, 0);
        if (emitted - (int)prod >= 2) {
          __nanosleep(60);
          if (clock64() - t_idle > 4000000000LL) __trap();
          continue;
        }
        const int k = lane;
        const bool valid = head + k < n_units && !((mask >> k) & 1u);
        const int code = valid ? p.sched[u_begin + head + k] : 0;
        const bool is_ln = valid && !(chain_kind(code) & 1);
        const uint32_t ln_bits = __ballot_sync(0xffffffffu, is_ln), valid_bits = __ballot_sync(0xffffffffu, valid);
        const int before = __popc(valid_bits & ((1u << k) - 1u));
        const bool eligible = valid && (is_ln ? (k == __ffs(ln_bits) - 1) : (before < p.ooo));
        const bool ready = eligible && unit_ready(code);
        const uint32_t rb = __ballot_sync(0xffffffffu, ready);
        if (rb == 0u) {
          __nanosleep(40);
          if (clock64() - t_idle > 4000000000LL) __trap();
          continue;
        }
        const int kk = __ffs(rb) - 1;
        const int chosen = __shfl_sync(0xffffffffu, code, kk);
        asm volatile("fence.proxy.async;" ::: "memory");  // (every lane: the ready lane's acquire is ordered before it)
        __syncwarp();
        if (lane == 0) {
          const unsigned long long w = ((unsigned long long)(uint32_t)(emitted + 1) << 32) | (uint32_t)chosen;
          asm volatile("st.relaxed.cluster.shared::cluster.u64 [%0], %1;" ::"r"(
                           mapa_cluster(smem_u32(s_inbox + (emitted & 15)), 1)), "l"(w) : "memory");
        }
        set_order(emitted, chosen);
        publish((uint32_t)(2 * emitted + 2));
        mask |= 1u << kk;
        ++emitted;
        while (mask & 1u) mask >>= 1, ++head;
        t_idle = clock64();
      }
    } else {
      // ---- run-time selection, peer CTA: take the leader's choice from the inbox, acquire the unit's counters in THIS
      // thread (they are satisfied: one pass; formally this CTA's TMA loads need their own acquire + proxy fence) ----
      for (int it = 0; it < n_units; ++it) {
        ring_space(it);
        int code = 0;
        if (lane == 0) {
          const long long t0 = clock64();
          unsigned long long w;
          for (;;) {
            asm volatile("ld.relaxed.cluster.shared::cta.u64 %0, [%1];" : "=l"(w) : "r"(smem_u32(s_inbox + (it & 15))) : "memory");
            if ((uint32_t)(w >> 32) == (uint32_t)(it + 1)) break;
            __nanosleep(40);
            if (clock64() - t0 > 4000000000LL) __trap();
          }
          code = (int)(uint32_t)w;
          while (!unit_ready(code)) {
            if (clock64() - t0 > 4000000000LL) __trap();
          }
        }
        code = __shfl_sync(0xffffffffu, code, 0);
        set_order(it, code);
        publish((uint32_t)(2 * it + 2));
      }
    }
   } else {
    // ===================== signal warp =====================
    // Walks the unit list behind the epilogue warps: per output signal (LN unit: high plane, then low plane; L1 / INP
    // unit: its tile) waits until all 16 epilogue warps of this CTA have seen their TMA stores complete
    // (cp.async.bulk.wait_group: the writes are visible to the waiting thread; its barrier arrival releases that to this
    // warp), then publishes with a gpu-scope release: fence + counter update (cumulative over the 16 warps' stores).
    // Without the fence the dependents occasionally read stale rows (1 evaluation in ~300 at the production shape).
    uint32_t seq = 0;
    const uint32_t prog_peer = mapa_cluster(smem_u32(&s_prog[rank]), rank ^ 1u);
    for (int k = 0; k < n_units; ++k) {
      wait_published((uint32_t)(2 * k + 1), 200);
      const int code = unit_code(k);
      const int kind = chain_kind(code), m = chain_m(code);
      const int nsig = (kind & 1) ? 1 : 2;
      for (int j = 0; j < nsig; ++j, ++seq) {
        const int c = kind == CK_LN1 ? j : (kind == CK_L1 ? 2 : (kind == CK_LN2 ? 3 + j : 5));
        mbar_wait_q(&sig_bar[seq & (CH_SIG_BARS - 1)], (seq / CH_SIG_BARS) & 1u);
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          red_relaxed_gpu_add(p.ctr + c * tiles_m + m, (unsigned)PW);
        }
        __syncwarp();
      }
      if (lane == 0) {  // this CTA is done with entry k of the order ring (flow control of the scouts, both CTAs)
        asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(&s_prog[rank])), "r"((uint32_t)(k + 1)) : "memory");
        asm volatile("st.relaxed.cluster.shared::cluster.u32 [%0], %1;" ::"r"(prog_peer), "r"((uint32_t)(k + 1)) : "memory");
      }
    }
   }
  } else {
    // ===================== epilogue (warps 0..15, both CTAs) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // LayerNorm parameters of this pair's column half for both norms of the layer being processed (weights).  LayerNorm
    // unit (row tile, half h) always runs on a pair with pair % halves == h (host schedule).  Stack form: restaged when
    // the unit list moves on to the next layer (all 16 warps walk the same list: two barriers per switch).
    auto weights_of = [&](int ul) -> const LayerWeightPtrs* { return p.lw ? p.lw + ul : nullptr; };
    auto stage_ln = [&](int ul) {
      const LayerWeightPtrs* lw = weights_of(ul);
      const int c0 = (pair % halves) * BN;
      for (int i = threadIdx.x; i < BN; i += PT) {
#pragma unroll
        for (int ln = 0; ln < 2; ++ln) {
          const float* b = lw ? lw->bias[ln ? CK_LN2 : CK_LN1] : p.bias[ln ? CK_LN2 : CK_LN1];
          s_ln[(ln * 3 + 0) * BN + i] = b ? b[c0 + i] : 0.f;
          s_ln[(ln * 3 + 1) * BN + i] = (lw ? lw->gamma[ln] : p.gamma[ln])[c0 + i];
          s_ln[(ln * 3 + 2) * BN + i] = (lw ? lw->beta[ln] : p.beta[ln])[c0 + i];
        }
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
    };
    int cur_layer = -1;  // staged at the first LayerNorm unit
    asm volatile("bar.sync 1, 512;" ::: "memory");
    const int lq = warp & 3, cq = warp >> 2;  // TMEM lane quarter, column quarter (64 columns)
    const int row_in_tile = lq * 32 + lane;
    const uint32_t tempty_leader = mapa_cluster(smem_u32(&tempty_bar[0]), 0);
    const uint32_t wst = smem_u32(s_stage) + warp * CH_STG_WARP;
    const int cl = cq * 64;  // first column of this warp's slab inside the 256-column tile
    uint32_t acc = 0, acc_phase = 0, ln_count = 0;
    int staged_key = -1;
    // Output signals.  Every output of a unit (LN: high plane, low plane; L1 / INP: the tile) is announced to the units
    // that read it by a counter update with RELEASE semantics at gpu scope -- a gpu-scope fence takes 1.5-3 k cycles
    // here, so the epilogue warps do not execute it: the elected lane of a warp waits for ITS TMA stores to complete and
    // arrives on the CTA's barrier of that signal (numbered in unit order, the same in all 16 warps); the signal warp
    // collects the 16 arrivals, fences and bumps the counter (below).  `pending`: a signal whose stores are still in
    // flight; it is handed over before the warp blocks on anything that may depend on it.
    bool pending = false;
    uint32_t sig_seq = 0, pending_seq = 0;
    // A warp never blocks while it owes a signal: the signal is flushed before any wait that may depend on it.
    auto flush_pending = [&]() {
      if (pending) {
        if (elect_one()) {
          bulk_wait<0>();
          mbar_arrive(&sig_bar[pending_seq & (CH_SIG_BARS - 1)]);
        }
        __syncwarp();
        pending = false;
      }
    };
    int it = 0;
    const bool stalls = p.trace && (p.dbg & 16);
    long long st_tfull = 0, st_stats = 0;
    for (; it < n_units; ++it) {
      if (dep_seen() < (uint32_t)(2 * it + 1)) {
        flush_pending();  // (a warp never blocks while it owes a signal)
        wait_published((uint32_t)(2 * it + 1), 100);
      }
      const int code = unit_code(it);
      const int kind = chain_kind(code), m = chain_m(code), n = chain_n(code), ul = chain_layer(code);
      if (!(kind & 1) && ul != cur_layer) {  // (only LayerNorm units read the staged parameters)
        asm volatile("bar.sync 1, 512;" ::: "memory");  // every warp has finished the LayerNorm units staged before
        stage_ln(ul);
        cur_layer = ul;
      }
      const int n0 = n * BN;
      const int grow0 = m * 256 + (int)rank * GEMM_BM + lq * 32;  // first global row of this warp
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(lq * 32) << 16) + cl;
      auto release_acc = [&]() {  // hand the drained accumulator stage back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + acc * 8);
      };
      if (kind & 1) {
        const int key = (ul * 4 + kind) * 256 + n;
        if (key != staged_key) {  // this warp's 64 bias values (a private slice: no barrier between the warps)
          const LayerWeightPtrs* lw = weights_of(ul);
          const float* bsrc = lw ? lw->bias[kind] : p.bias[kind];
          const int c = n0 + cl + lane;
          __syncwarp();
          s_bias2[warp * 64 + lane] = bsrc ? bsrc[c] : 0.f;
          s_bias2[warp * 64 + lane + 32] = bsrc ? bsrc[c + 32] : 0.f;
          __syncwarp();
          staged_key = key;
        }
      } else {
        flush_pending();  // a LayerNorm unit blocks on its partner's statistics: no signal may be owed while it does
      }
      if (!mbar_try_wait(&tfull_bar[acc], acc_phase)) {
        flush_pending();
        const long long t0 = stalls ? clock64() : 0;
        mbar_wait_q(&tfull_bar[acc], acc_phase);
        if (stalls) st_tfull += clock64() - t0;
      }
      tc_fence_after();
      if (threadIdx.x == 0) CHAIN_TRACE_UNIT(4, it);
      uint32_t v0[32], v1[32];
      tmem_ld32(taddr, v0);
      tc_wait_ld_dep(v0);
      tmem_ld32(taddr + 32, v1);
      if (!(kind & 1)) {
        // ---------------- x = LayerNorm(acc + b)   (acc already holds A . W^T + residual) ----------------
        const int ln = kind >> 1;
        const float* s_b1 = s_ln + (ln * 3 + 0) * BN;
        const float* s_gamma = s_ln + (ln * 3 + 1) * BN;
        const float* s_beta = s_ln + (ln * 3 + 2) * BN;
        const int col0 = n0 + cl;  // global column of the slab
        // ---- pass 1: y = acc + bias stays in registers; row sum and sum of squares (packed fp32 pairs: the epilogue is
        //      issue bound -- 16 warps on 4 schedulers) ----
        f32x2 sa = pk2(0.f, 0.f), sb = pk2(0.f, 0.f), qa = pk2(0.f, 0.f), qb = pk2(0.f, 0.f);
        auto pass1 = [&](uint32_t (&v)[32], int cbase) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            asm volatile("" ::: "memory");  // keep the parameter loads next to their use (64 live accumulator registers)
            const float4 b4 = *reinterpret_cast<const float4*>(s_b1 + cbase + j * 4);
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
          // One 64-bit word per (row, reader): the sum of squares is >= 0 (or a NaN, made positive), so a posted word
          // never equals the all-ones "not posted" pattern and the single store is its own flag (no fence).  The warp
          // with column quarter 0 posts four copies, one per reading quarter of the partner CTA; a reader resets its
          // copy after reading, which leaves the state clean for the next launch.
          const size_t base = ((size_t)(ln * tiles_m + m) * 2) * 2;
          const size_t mine = ((base + (size_t)n * 2 + rank) * 4) * 128 + row_in_tile;
          const size_t theirs = ((base + (size_t)(n ^ 1) * 2 + rank) * 4 + cq) * 128 + row_in_tile;
          if (cq == 0) {
            const unsigned long long w = (unsigned long long)__float_as_uint(tot_s) |
                                         ((unsigned long long)(__float_as_uint(tot_q) & 0x7fffffffu) << 32);
#pragma unroll
            for (int c = 0; c < 4; ++c) st_relaxed_gpu_u64(p.stats + mine + c * 128, w);
          }
          unsigned long long o = ld_relaxed_gpu_u64(p.stats + theirs);
          if (o == ~0ull) {
            const long long t0 = clock64();
            while ((o = ld_relaxed_gpu_u64(p.stats + theirs)) == ~0ull) {
              __nanosleep(20);
              if (clock64() - t0 > 4000000000LL) __trap();
            }
            if (stalls) st_stats += clock64() - t0;
          }
          st_relaxed_gpu_u64(p.stats + theirs, ~0ull);  // single reader: reset for the next launch
          tot_s += __uint_as_float((uint32_t)o), tot_q += __uint_as_float((uint32_t)(o >> 32));
          // a + b == b + a: both halves normalise with bit-identical statistics
        }
        if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(58);  // statistics complete (first unit of the pair)
        const float inv_n = 1.0f / (float)p.d;
        const float mean = tot_s * inv_n;
        const float var = fmaxf(tot_q * inv_n - mean * mean, 0.f);  // biased variance (F.layer_norm), fp32
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        const float nmr = -mean * rstd;
        // ---- pass 2: normalise + affine, hi plane -> staging -> TMA store, then lo plane ----
        if (pending) {
          flush_pending();  // (also: the previous unit's store has finished reading the staging tile)
        } else {
          if (elect_one()) bulk_wait_read<0>();
          __syncwarp();
        }
        const f32x2 rstd2 = pk2(rstd, rstd), nmr2 = pk2(nmr, nmr);
        // y = ((acc - mean) rstd) gamma + beta on packed pairs; the hi plane bf16(y) goes to the staging tile, the
        // register pair is replaced by the lo plane's input y - bf16(y) (exact in fp32)
        // The warp's staging tile holds 32 columns: the 64-column slab leaves in two halves per plane, each half staged
        // once the previous half's store has finished READING the tile (the math of the next half runs under that read).
        auto wait_tile_free = [&]() {
          if (elect_one()) bulk_wait_read<0>();
          __syncwarp();
        };
        auto store_half = [&](const CUtensorMap* tm, int col) {
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one() && !(p.dbg & 1)) {
            tma_store_2d(tm, wst, col, grow0);  // rows past M are clipped by the tensor map
            bulk_commit();
          }
        };
        auto norm_hi = [&](uint32_t (&v)[32], int cbase, bool wait_first) {
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
            if (wait_first && jj == 0) wait_tile_free();  // (after the first chunk's math)
            sts128(wst + stg64_off(lane, jj), make_uint4(hi[0], hi[1], hi[2], hi[3]));
          }
        };
        norm_hi(v0, cl, false);
        store_half(&tmXh_st, col0);
        norm_hi(v1, cl + 32, true);
        store_half(&tmXh_st, col0 + 32);
        if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(59);  // hi plane staged + stores issued
        auto stage_lo = [&](const uint32_t (&v)[32]) {  // lo = bf16(y - bf16(y)), the difference is in v
          uint32_t lo[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) lo[q] = pack_bf16x2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
          wait_tile_free();
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            sts128(wst + stg64_off(lane, jj), make_uint4(lo[4 * jj], lo[4 * jj + 1], lo[4 * jj + 2], lo[4 * jj + 3]));
        };
        stage_lo(v0);
        if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(60);
        store_half(&tmXl_st, col0);
        // The HIGH plane (the A operand of the L1 / INP units of this row tile) is announced as soon as its two stores
        // have completed: all but the newest bulk group of this thread.
        if (elect_one()) {
          if (threadIdx.x == 0 && it == 0) CHAIN_TRACE(61);
          bulk_wait<1>();
          if (threadIdx.x == 0 && it == 0) {
            CHAIN_TRACE(62);
            CHAIN_TRACE_NS(55);
          }
          if (p.trace && it == 0)  // latest warp of this CTA to have its hi plane complete (ns)
            atomicMax(reinterpret_cast<unsigned long long*>(p.trace + (size_t)blockIdx.x * GEMM_TRACE_SLOTS + 53),
                      (unsigned long long)chain_globaltimer());
          mbar_arrive(&sig_bar[sig_seq & (CH_SIG_BARS - 1)]);
        }
        __syncwarp();
        stage_lo(v1);
        store_half(&tmXl_st, col0 + 32);
        __syncwarp();
        // the low plane is the residual input of the next LayerNorm of this row tile (LN2 here, LN1 of the next layer
        // kernel): owed once the store has completed
        pending = true, pending_seq = sig_seq + 1;
        sig_seq += 2;
      } else {
        // ---------------- bias (+ GELU) -> bf16: the warp's 32 x 64 slab leaves as one TMA store ----------------
        const float* wb = s_bias2 + warp * 64;
        if (pending) {
          flush_pending();
        } else {
          if (elect_one()) bulk_wait_read<0>();  // the previous unit's store has finished reading the staging tile
          __syncwarp();
        }
        // 32 columns at a time: the second half is computed into registers while the first half's store reads the tile
        auto half = [&](const uint32_t (&vv)[32], int j0, uint32_t (&o)[16], auto gelu) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const int j = j0 + jj;
            const uint32_t* v = vv + jj * 8;
            if constexpr (decltype(gelu)::value) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {  // packed pairs: bias add + GELU as FADD2 / FFMA2 chains
                const float2 b2 = *reinterpret_cast<const float2*>(wb + j * 8 + 2 * e);
                const f32x2 y = gelu_erf2(add2(pk2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1])), pk2(b2.x, b2.y)));
                o[4 * jj + e] = pack_bf16x2(pk_lo(y), pk_hi(y));
              }
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float2 b2 = *reinterpret_cast<const float2*>(wb + j * 8 + 2 * e);
                o[4 * jj + e] = pack_bf16x2(__uint_as_float(v[2 * e]) + b2.x, __uint_as_float(v[2 * e + 1]) + b2.y);
              }
            }
          }
        };
        auto stage_store = [&](const uint32_t (&o)[16], int col) {
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            sts128(wst + stg64_off(lane, jj), make_uint4(o[4 * jj], o[4 * jj + 1], o[4 * jj + 2], o[4 * jj + 3]));
          fence_proxy_async_smem();
          __syncwarp();
          if (elect_one() && !(p.dbg & 1)) {
            tma_store_2d(kind == CK_L1 ? &tmHst : &tmQst, wst, col, grow0);  // rows past M are clipped
            bulk_commit();
          }
        };
        uint32_t o[16];
        if (kind == CK_L1)
          half(v0, 0, o, std::true_type{});
        else
          half(v0, 0, o, std::false_type{});
        stage_store(o, n0 + cl);
        tc_wait_ld_dep(v1);
        release_acc();
        if (kind == CK_L1)
          half(v1, 4, o, std::true_type{});
        else
          half(v1, 4, o, std::false_type{});
        if (elect_one()) bulk_wait_read<0>();  // the first half's store has finished reading the tile
        __syncwarp();
        stage_store(o, n0 + cl + 32);
        // an H tile is an input of this row tile's LN2 units, a QKV tile of the attention CTAs of the next layer: owed
        // once the store has completed
        pending = true, pending_seq = sig_seq;
        sig_seq += 1;
      }
      if (threadIdx.x == 0) CHAIN_TRACE_UNIT(5, it);
      if (++acc == 2) acc = 0, acc_phase ^= 1u;
    }
    flush_pending();
    if (elect_one()) bulk_wait<0>();  // this thread's TMA stores have been performed before the CTA retires
    if (stalls && threadIdx.x == 0) {
      long long* z = p.trace + (size_t)gridDim.x * GEMM_TRACE_SLOTS + (size_t)blockIdx.x * 8;
      z[4] = st_tfull, z[5] = st_stats;
    }
  }
  tc_fence_before();
  cluster_sync_all();  // the peer may still signal our barriers / read our TMEM half
  if (threadIdx.x == 0) {
    CHAIN_TRACE(3);
    ktime_exit(p.ktime);
    if (p.trace && (p.dbg & 16)) p.trace[(size_t)gridDim.x * GEMM_TRACE_SLOTS + (size_t)blockIdx.x * 8 + 6] = clock64();  // (slot 0 holds the start)
  }
  if (warp == PW + 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ---- host side ------------------------------------------------------------------------------------
// Static schedule of one layer kernel: per CTA pair the list of unit codes (kind << 28 | row tile << 8 | column tile).
struct LayerSchedule {
  int pairs = 0, tiles_m = 0, halves = 0;
  std::vector<int> off, units;
  double makespan = 0.0;
};

// Cost estimates in cycles (profiles/r02_chain_*): mainloop time of a unit, epilogue time, the offset inside the
// epilogue (LN: from its start; L1: from its end) at which the unit's output is usable by a dependent unit.
struct LayerCosts {
  double kb = 625.0;        // one 64-deep k-block of a 256 x 256 pair tile (ring-latency bound with 4 stages; 512 at the MMA rate)
  double res = 2000.0;      // the 4 residual stages of a LayerNorm unit
  double epi_ln = 9500.0, ln_ready = 7000.0;   // LayerNorm epilogue; high plane complete this long after its start
  double epi_gelu = 5300.0, epi_bias = 2000.0;
  double signal = 2500.0;   // store completion + counter update + scout poll + first TMA loads of the dependent unit
  double slack = 8000.0;    // a unit may start this much later than on the earliest pair if that pair is less loaded
                            // (tools/sched_sweep.py, profiles/r02_sched_sweep.txt: the step time moves by < 2 % over wide
                            // ranges of every constant; only signal = 0 -- dependents scheduled too early -- costs 12 %)
};

// List scheduling of the per-row-tile chains LN1 -> L1 (ff / 256 tiles) -> LN2 -> INP (n_inp / 256 tiles) on `slots`
// pairs.  A pair is modelled as two resources (tensor pipe, epilogue warps) with two accumulator stages; a LayerNorm row
// tile is a job for a DUO of neighbouring pairs (2k, 2k + 1: the two column halves run at the same list position and
// exchange statistics).  Units are placed in order of their simulated start time, so every pair's list is a subsequence
// of one global topological order.
// Attention side of the stack form: `ctas` persistent attention CTAs, `unit` cycles per (sequence, head) unit.
struct AttnModel {
  int ctas = 0, S = 1, heads = 1;
  double unit = 9600.0;
};

// List scheduling of `L` layers of the per-row-tile chain LN1 -> L1 -> LN2 -> INP on `slots` CTA pairs.  L == 1 is the
// per-layer form (one launch per layer; `last_has_inp` says whether the launch ends with the next layer's in_proj, all row
// tiles are ready at t = 0).  L > 1 is the stack form (ONE launch for all layers; only the last layer has no INP): the
// attention of a row tile between INP of layer l and LN1 of layer l + 1 runs on other SMs and is modelled as a server
// with `att.ctas` units in flight.  Unit codes carry the layer (chain_code).
inline LayerSchedule build_stack_schedule(int M, int d, int ff, int L, bool last_has_inp, int slots, const LayerCosts& c,
                                          const AttnModel& att) {
  LayerSchedule s;
  s.tiles_m = (M + 255) / 256;
  s.halves = d / CH_BN;
  const int T = s.tiles_m, H = s.halves, n1 = ff / CH_BN, n3 = 3 * d / CH_BN;
  int pairs = slots;
  const long units_full = (long)T * (2 * H + n1 + n3), units_last = (long)T * (2 * H + n1 + (last_has_inp ? n3 : 0));
  const long total_units = units_full * (L - 1) + units_last;
  if ((long)pairs > total_units) pairs = (int)total_units;
  if (H == 2) pairs &= ~1;
  if (pairs < H) pairs = H;
  s.pairs = pairs;
  struct Pair {
    double mma_free = 0, epi_free = 0, epi_prev = 0, epi_last = 0;  // epi_prev: end of the epilogue before the last one
  };
  std::vector<Pair> ps(pairs);
  std::vector<std::vector<int>> lists(pairs);
  // place a unit on pair pr with inputs ready at `r`: returns {epilogue start, epilogue end}
  auto place = [&](int pr, double r, double mma, double epi, bool commit, double* epi_start_out) {
    Pair q = ps[pr];
    const double mma_start = std::max(std::max(q.mma_free, r), q.epi_prev);  // needs a free accumulator stage
    const double mma_end = mma_start + mma;
    const double epi_start = std::max(mma_end, q.epi_free);
    const double epi_end = epi_start + epi;
    if (commit) {
      q.mma_free = mma_end, q.epi_prev = q.epi_last, q.epi_last = epi_end, q.epi_free = epi_end;
      ps[pr] = q;
    }
    if (epi_start_out) *epi_start_out = epi_start;
    return epi_end;
  };
  const double mma_ln1 = c.kb * d / 64 + c.res, mma_ln2 = c.kb * ff / 64 + c.res, mma_t = c.kb * d / 64;
  // per row tile: stage = 4 * layer + (0 LN1 pending, 1 L1 tiles, 2 LN2 pending, 3 INP tiles); 4 L = done
  const int done_stage = 4 * L;
  std::vector<int> stage(T, 0), next_n(T, 0);
  std::vector<double> ready(T, 0.0), acc_ready(T, 0.0);
  // attention server: time per row tile with all CTAs busy, and the latency of one unit
  const double att_tile = att.ctas > 0 ? att.unit * att.heads * 256.0 / att.S / att.ctas : 0.0;
  double att_free = 0.0;
  auto attention = [&](double inputs_ready) {  // -> time at which the row tile's attention output is complete
    if (att.ctas <= 0) return inputs_ready;
    const double start = std::max(inputs_ready, att_free);
    att_free = start + att_tile;
    return att_free + att.unit;
  };
  if (L > 1)
    for (int m = 0; m < T; ++m) ready[m] = attention(0.0);  // the first layer's attention, row tiles in order
  long remaining = total_units;
  std::vector<double> load(pairs, 0.0);  // tensor-pipe work assigned to a pair so far
  while (remaining > 0) {
    // candidate = the next unit of each row tile; pick the row tile that can start earliest (ties: the deeper stage
    // first -- finish chains that are under way -- then the lower row tile), and for it, among the pairs (duos) that can
    // start within `slack` of the earliest one, the least loaded: start times stay near-optimal while the long
    // LayerNorm units spread over all duos instead of piling up on those that happened to be idle first.
    int best_m = -1, best_pr = -1;
    double best_start = 1e300;
    for (int m = 0; m < T; ++m) {
      if (stage[m] >= done_stage) continue;
      const int k4 = stage[m] & 3;
      const bool ln = (k4 == 0 || k4 == 2);
      const int step = ln ? H : 1;
      double st_min = 1e300;
      for (int k = 0; k + step <= pairs; k += step) {
        double st = 0;
        for (int h = 0; h < step; ++h) {
          const Pair& q = ps[k + h];
          st = std::max(st, std::max(std::max(q.mma_free, ready[m]), q.epi_prev));
        }
        st_min = std::min(st_min, st);
      }
      const double key = st_min - 1e-3 * stage[m];  // deeper stage wins ties
      if (key < best_start - 1e-9) best_start = key, best_m = m;
    }
    {
      const int m = best_m;
      const int k4 = stage[m] & 3;
      const bool ln = (k4 == 0 || k4 == 2);
      const int step = ln ? H : 1;
      const double limit = best_start + 1e-3 * stage[m] + c.slack;
      double best_load = 1e300;
      for (int k = 0; k + step <= pairs; k += step) {
        double st = 0, ld = 0;
        for (int h = 0; h < step; ++h) {
          const Pair& q = ps[k + h];
          st = std::max(st, std::max(std::max(q.mma_free, ready[m]), q.epi_prev));
          ld = std::max(ld, load[k + h]);
        }
        if (st <= limit && ld < best_load - 1e-9) best_load = ld, best_pr = k;
      }
    }
    const int m = best_m;
    const int st = stage[m] & 3, lay = stage[m] >> 2;
    const bool has_inp = lay + 1 < L || last_has_inp;
    if (st == 0 || st == 2) {
      double done = 0, duo_start = ready[m];
      for (int h = 0; h < H; ++h) {
        const Pair& q = ps[best_pr + h];
        duo_start = std::max(duo_start, std::max(q.mma_free, q.epi_prev));
      }
      for (int h = 0; h < H; ++h) {
        double es = 0;
        // both halves start together: feed the common start time as the ready time
        place(best_pr + h, duo_start, st == 0 ? mma_ln1 : mma_ln2, c.epi_ln, true, &es);
        done = std::max(done, es + c.ln_ready + c.signal);
        lists[best_pr + h].push_back(chain_code(lay, st == 0 ? CK_LN1 : CK_LN2, m, h));
        load[best_pr + h] += st == 0 ? mma_ln1 : mma_ln2;
        --remaining;
      }
      ready[m] = done, acc_ready[m] = 0.0;
      ++stage[m], next_n[m] = 0;
      if (st == 2 && !has_inp) stage[m] = 4 * (lay + 1);  // (only the last layer: the row tile is finished)
    } else {
      const int nt = st == 1 ? n1 : n3;
      const double end = place(best_pr, ready[m], mma_t, st == 1 ? c.epi_gelu : c.epi_bias, true, nullptr);
      lists[best_pr].push_back(chain_code(lay, st == 1 ? CK_L1 : CK_INP, m, next_n[m]));
      load[best_pr] += mma_t;
      --remaining;
      acc_ready[m] = std::max(acc_ready[m], end + c.signal);
      if (++next_n[m] == nt) {
        ++stage[m];
        // LN2 needs every H tile of the row tile; the next layer's LN1 the attention over the row tile's in_proj tiles
        ready[m] = st == 1 ? acc_ready[m] : attention(acc_ready[m]);
      }
    }
  }
  s.off.assign(pairs + 1, 0);
  for (int pr = 0; pr < pairs; ++pr) {
    s.off[pr + 1] = s.off[pr] + (int)lists[pr].size();
    s.units.insert(s.units.end(), lists[pr].begin(), lists[pr].end());
    s.makespan = std::max(s.makespan, ps[pr].epi_free);
  }
  return s;
}

// One layer per launch (every row tile ready at t = 0); n_inp = 0: the last layer (no next in_proj).
inline LayerSchedule build_layer_schedule(int M, int d, int ff, int n_inp, int slots, const LayerCosts& c) {
  return build_stack_schedule(M, d, ff, 1, n_inp > 0, slots, c, AttnModel{});
}

inline int configure_layer_chain() {
  TAMF_CUDA_CHECK(cudaFuncSetAttribute(layer_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES));
  return TAMF_OK;
}

struct LayerMaps {
  const CUtensorMap *ATT, *Wo, *Xh, *Xl, *W1, *Hst, *H, *W2, *Win, *Qst, *Xh_st, *Xl_st, *I;  // Win / Qst null: no INP
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

inline int launch_layer_chain(const LayerMaps& tm, const LayerParams& p, int pairs, cudaStream_t stream) {
  TAMF_REQUIRE(p.d == 256 || p.d == 512, TAMF_E_BADARG, "layer_chain: LayerNorm width must be 256 or 512");
  TAMF_REQUIRE(p.ff > 0 && p.ff % 256 == 0 && p.n_inp % 256 == 0, TAMF_E_BADARG,
               "layer_chain: ff and the in_proj width must be multiples of 256");
  TAMF_REQUIRE(tm.ATT && tm.Wo && tm.Xh && tm.Xl && tm.W1 && tm.Hst && tm.H && tm.W2 && tm.Xh_st && tm.Xl_st && tm.I &&
                   (p.n_inp == 0 || (tm.Win && tm.Qst)),
               TAMF_E_BADARG, "layer_chain: missing tensor map");
  TAMF_REQUIRE(pairs >= 1 && 2 * pairs <= num_sms(), TAMF_E_BADARG, "layer_chain: the grid must be co-resident");
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
  const CUtensorMap& win = tm.Win ? *tm.Win : *tm.W1;  // unused maps are passed as copies
  const CUtensorMap& qst = tm.Qst ? *tm.Qst : *tm.Hst;
  cudaError_t e = cudaLaunchKernelEx(&cfg, layer_chain_kernel, *tm.ATT, *tm.Wo, *tm.Xh, *tm.Xl, *tm.W1, *tm.Hst, *tm.H,
                                     *tm.W2, win, qst, *tm.Xh_st, *tm.Xl_st, *tm.I, p);
  count_launch();
  if (e != cudaSuccess) {
    set_error(std::string("layer_chain launch failed: ") + cudaGetErrorString(e));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

}  // namespace tamf
