// gemm.cuh -- persistent warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   C[M,N] = A[M,K] (bf16, K-major) . W[N,K]^T (bf16, K-major = nn.Linear layout), fp32 accumulation in TMEM.
//
// One CTA per SM, 320 threads:
//   warps 0-7  epilogue: tcgen05.ld the 128 x BN fp32 accumulator (warp w owns TMEM lanes 32*(w%4).., column half
//              w/4) and apply the fused epilogue.  TMEM hands every thread one accumulator ROW, so results are
//              transposed through a warp-private XOR-swizzled shared-memory tile and leave the SM as full sectors;
//              the fp32 residual of the LayerNorm epilogue arrives the same way (cp.async, one chunk ahead).
//   warp 8     TMA producer: cp.async.bulk.tensor 2D loads of the 128x64 A tile and BNx64 W tile (128B swizzle)
//              into a STAGES-deep shared-memory ring, signalled through mbarriers (expect_tx / complete_tx).
//   warp 9     MMA issuer: one thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16) on
//              shared-memory descriptors; tcgen05.commit releases ring slots and publishes accumulators.
// The accumulator is double-buffered in TMEM when 2*BN <= 512 columns, so the epilogue of tile i overlaps the
// mainloop of tile i+1 (persistent static round-robin tile schedule).  Launched with programmatic dependent launch:
// barrier init, TMEM allocation, descriptor prefetch and the staging of bias/gamma/beta overlap the previous
// kernel's tail; griddepcontrol.wait precedes every access to activations.
//
// Epilogues (what the reference runs as separate ATen kernels, SURVEY.md 2.4 K1,K3-K6,K9):
//   EPI_BIAS_BF16      y = acc + b                                   -> bf16            (QKV in_proj)
//   EPI_BIAS_GELU_BF16 y = gelu_erf(acc + b)                         -> bf16            (linear1 + F.gelu)
//   EPI_ADD_SILU_BF16  y = silu(acc + addmat[m,n])                   -> bf16            (input_merge.0, hand half)
//   EPI_TOKEN_OUT      y = nan_to_num(acc + b) + pe[P0+tau]          -> fp32 + bf16 token rows (input_merge.2)
//   EPI_RES_LN         x = LayerNorm(x + acc + b) (eps 1e-5, biased var) in place -> fp32 + bf16 (BN == N == d)
//   EPI_POSTERIOR      x0 = nan_to_num(acc + b); x_{t-1} = c1 x0 + c2 x_t + sigma eps  -> [B,99,1,T] fp32
//   EPI_RESIDUAL_OUT   y = nan_to_num(x_in + acc + b) -> [B,T,99] fp32 (MF-MDM R: segment_refine_model.py:215-217)
//   EPI_F32            y = acc + b -> fp32 (self-test)
#pragma once
#include "common.cuh"

namespace tamf {

enum GemmEpi {
  EPI_BIAS_BF16 = 0,
  EPI_BIAS_GELU_BF16 = 1,
  EPI_ADD_SILU_BF16 = 2,
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
  // EPI_ADD_SILU_BF16
  const float* addmat;  // [M,N]
  // EPI_TOKEN_OUT: GEMM row m = b*T + tau  ->  token row b*S + P0 + tau
  const float* pe;  // [rows, N]
  int T, S, P0;
  // EPI_RES_LN: X fp32 [M,N] residual in / normalised out; Xb bf16 copy
  float* X;
  __nv_bfloat16* Xb;
  const float* gamma;
  const float* beta;
  // EPI_POSTERIOR / EPI_RESIDUAL_OUT: GEMM row m = b*S + s (token); frame tau = s - P0
  const float* x_t;    // [B,nfeat,1,T] (POSTERIOR) | x_in [B,T,nfeat] (RESIDUAL_OUT)
  float* x_out;        // x_{t-1}, may alias x_t; null -> forward only
  float* x0_out;       // may be null
  const float* noise;  // null -> Philox
  const int* t_ptr;    // [B]
  const float *c1, *c2, *sigma;  // [num_steps] fp32
  unsigned long long seed;
  int nfeat;
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_EPI_THREADS = 256;
constexpr int GEMM_CTRL_BYTES = 256;      // mbarriers + TMEM slot
constexpr int GEMM_STG_WARP = 2048;       // bf16 epilogues: warp-private staging tile, 32 rows x 64 B
constexpr int GEMM_LN_STG_WARP = 14336;   // LN epilogue, per warp: 2 x 4 KB residual in | 4 KB fp32 out | 2 KB bf16 out

constexpr bool epi_is_ln(int e) { return e == EPI_RES_LN; }
constexpr bool epi_staged_bf16(int e) { return e == EPI_BIAS_BF16 || e == EPI_BIAS_GELU_BF16 || e == EPI_ADD_SILU_BF16; }

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 128) ? 6 : (BN == 256 ? 4 : 2);
  static constexpr int ACC_STAGES = (2 * BN <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = (ACC_STAGES * BN <= 256) ? 256 : 512;
  static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
  static constexpr int B_BYTES = BN * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static_assert(PIPE_BYTES >= 8 * GEMM_LN_STG_WARP, "LN staging must fit in the pipeline ring");
};

// Shared-memory map: [pipeline ring | control | parameters | staging].  LN epilogues stage through the drained
// ring (one tile in flight); the bf16 epilogues own a staging area because they overlap the next tile's mainloop.
template <int BN, int EPI>
struct GemmSmem {
  using Cfg = GemmCfg<BN>;
  static constexpr int PARAM_BYTES = epi_is_ln(EPI) ? (3 * BN * 4 + 1024) : (2 * BN * 4);  // LN: b|gamma|beta|stats
  static constexpr int STG_BYTES = epi_staged_bf16(EPI) ? 8 * GEMM_STG_WARP : 0;
  static constexpr int BYTES = 1024 /*align slack*/ + Cfg::PIPE_BYTES + GEMM_CTRL_BYTES + PARAM_BYTES + STG_BYTES;
  static_assert(BYTES <= 232448, "shared memory budget (227 KB) exceeded");
};

// exact-erf GELU (F.gelu default, the reference's activation="gelu")
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

// Programmatic dependent launch (PDL) controls
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
    gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  using Sm = GemmSmem<BN, EPI>;
  constexpr int STAGES = Cfg::STAGES, ACC = Cfg::ACC_STAGES;
  constexpr bool LN = epi_is_ln(EPI);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  uint8_t* smem = smem_raw + pad;
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint8_t* ctrl = smem + Cfg::PIPE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ctrl);  // [STAGES]
  uint64_t* empty_bar = full_bar + STAGES;                  // [STAGES]
  uint64_t* tfull_bar = empty_bar + STAGES;                 // [ACC]
  uint64_t* tempty_bar = tfull_bar + ACC;                   // [ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + ACC);
  float* s_par = reinterpret_cast<float*>(ctrl + GEMM_CTRL_BYTES);
  // LN: bias | gamma | beta | row statistics [2][128];  otherwise: bias slices of two consecutive tiles
  float* s_gamma = s_par + BN;
  float* s_beta = s_par + 2 * BN;
  float* s_red = s_par + 3 * BN;
  uint8_t* s_stage = LN ? smem : (ctrl + GEMM_CTRL_BYTES + Sm::PARAM_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int tiles_n = (p.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 8);  // one elected lane per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  if constexpr (LN) {
    if (warp < 8) {  // weights only (constant over the chain): staged before the dependency wait
      for (int i = threadIdx.x; i < BN; i += GEMM_EPI_THREADS) {
        s_par[i] = p.bias ? p.bias[i] : 0.f;
        s_gamma[i] = p.gamma[i];
        s_beta[i] = p.beta[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();  // the next kernel may run its own prologue on SMs this grid has already left
  pdl_wait();               // everything the previous kernel wrote is visible from here on

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * GEMM_BM, n0 = (tile % tiles_n) * BN;
        if constexpr (LN) {
          // the LN epilogue stages through the pipeline ring: do not refill it before that epilogue has drained
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
          if (++acc == ACC) acc = 0, acc_phase ^= 1u;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full_bar[stage], kb * GEMM_BK, m0);
#pragma unroll
          for (int h = 0; h < (BN + 255) / 256; ++h)  // TMA box rows <= 256
            tma_load_2d(sB + stage * Cfg::B_BYTES + h * 256 * GEMM_BK * 2, &tmB, &full_bar[stage], kb * GEMM_BK,
                        n0 + h * 256);
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr int UN = (BN > 256) ? 256 : BN;  // N per instruction
      constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, UN);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            const uint64_t ad = umma_desc_k_sw128(a_addr + k * 32);
#pragma unroll
            for (int h = 0; h < BN / UN; ++h) {
              const uint64_t bd = umma_desc_k_sw128(b_addr + h * UN * GEMM_BK * 2 + k * 32);
              umma_bf16(d_tmem + h * UN, ad, bd, idesc, (kb | k) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);  // slot reusable once these MMAs have read it
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == STAGES) stage = 0, phase ^= 1u;
        }
        if (++acc == ACC) acc = 0, acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue (warps 0..7) =====================
    const int lq = warp & 3, ch = warp >> 2;
    const int row_in_tile = lq * 32 + lane;
    constexpr int HALF = BN / 2;
    constexpr int CHUNKS = HALF / 32;
    uint32_t acc = 0, acc_phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m0 = (tile / tiles_n) * GEMM_BM, n0 = (tile % tiles_n) * BN;
      const int row = m0 + row_in_tile;
      const bool row_ok = row < p.M;
      const int grow0 = m0 + lq * 32;  // first global row of this warp
      const float* s_bias = s_par;
      if constexpr (!LN) {
        // this tile's bias slice -> buffer (it & 1).  One barrier per tile is enough: a warp that reaches it has
        // finished tile it-1, so nobody still reads the buffer being overwritten for tile it+1.
        float* sb = s_par + (it & 1) * BN;
        for (int i = threadIdx.x; i < BN; i += GEMM_EPI_THREADS) sb[i] = (p.bias && n0 + i < p.N) ? p.bias[n0 + i] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        s_bias = sb;
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(lq * 32) << 16) + ch * HALF;

      if constexpr (LN) {
        const uint32_t wst = smem_u32(s_stage) + warp * GEMM_LN_STG_WARP;
        const uint32_t s_in0 = wst, s_outf = wst + 8192, s_outb = wst + 12288;
        const int crow = lane >> 3, cpiece = lane & 7;  // global side of the fp32 tiles: 4 rows x 8 pieces per pass
        const float* xin = p.X + (size_t)ch * HALF;     // this warp's column half
        auto prefetch = [&](int ck) {                   // residual chunk ck -> s_in[ck & 1], full 128-byte lines
          const uint32_t dst = s_in0 + (ck & 1) * 4096;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + crow;
            const int gr = min(grow0 + r, p.M - 1);  // rows past M read a valid row; they are never stored
            cp_async_16(dst + stg128_off(r, cpiece), xin + (size_t)gr * p.N + ck * 32 + cpiece * 4);
          }
          cp_async_commit();
        };
        // ---- pass 1: v = acc + bias + residual -> back to TMEM; row sum ----
        prefetch(0);
        float sum = 0.f;
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          if (ck + 1 < CHUNKS) {
            prefetch(ck + 1);
            cp_async_wait<1>();
          } else {
            cp_async_wait<0>();
          }
          __syncwarp();
          uint32_t v[32];
          tmem_ld32(taddr + ck * 32, v);
          tc_wait_ld();
          const int c0 = ch * HALF + ck * 32;
          const uint32_t src = s_in0 + (ck & 1) * 4096;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 r4 = lds128(src + stg128_off(lane, j));
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c0 + j * 4);
            const float y0 = __uint_as_float(v[4 * j]) + b4.x + __uint_as_float(r4.x);
            const float y1 = __uint_as_float(v[4 * j + 1]) + b4.y + __uint_as_float(r4.y);
            const float y2 = __uint_as_float(v[4 * j + 2]) + b4.z + __uint_as_float(r4.z);
            const float y3 = __uint_as_float(v[4 * j + 3]) + b4.w + __uint_as_float(r4.w);
            sum += (y0 + y1) + (y2 + y3);
            v[4 * j] = __float_as_uint(y0), v[4 * j + 1] = __float_as_uint(y1);
            v[4 * j + 2] = __float_as_uint(y2), v[4 * j + 3] = __float_as_uint(y3);
          }
          tmem_st32(taddr + ck * 32, v);
          __syncwarp();  // the buffer just read is refilled by the next iteration's prefetch
        }
        tc_wait_st();
        s_red[ch * 128 + row_in_tile] = sum;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float mean = (s_red[row_in_tile] + s_red[128 + row_in_tile]) / (float)p.N;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        // ---- pass 2: centred second moment (TMEM only) ----
        float sq = 0.f;
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          uint32_t v[32];
          tmem_ld32(taddr + ck * 32, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = __uint_as_float(v[j]) - mean;
            sq = fmaf(d, d, sq);
          }
        }
        s_red[ch * 128 + row_in_tile] = sq;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float var = (s_red[row_in_tile] + s_red[128 + row_in_tile]) / (float)p.N;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        // ---- pass 3: normalise + affine; fp32 residual stream and bf16 operand copy leave as full sectors ----
        float* xo = p.X + (size_t)ch * HALF;
        __nv_bfloat16* xbo = p.Xb + (size_t)ch * HALF;
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          uint32_t v[32];
          tmem_ld32(taddr + ck * 32, v);
          tc_wait_ld();
          const int c0 = ch * HALF + ck * 32;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 g4 = *reinterpret_cast<const float4*>(s_gamma + c0 + j * 4);
            const float4 be4 = *reinterpret_cast<const float4*>(s_beta + c0 + j * 4);
            const float y0 = (__uint_as_float(v[4 * j]) - mean) * rstd * g4.x + be4.x;
            const float y1 = (__uint_as_float(v[4 * j + 1]) - mean) * rstd * g4.y + be4.y;
            const float y2 = (__uint_as_float(v[4 * j + 2]) - mean) * rstd * g4.z + be4.z;
            const float y3 = (__uint_as_float(v[4 * j + 3]) - mean) * rstd * g4.w + be4.w;
            sts128(s_outf + stg128_off(lane, j),
                   make_uint4(__float_as_uint(y0), __float_as_uint(y1), __float_as_uint(y2), __float_as_uint(y3)));
            v[2 * j] = pack_bf16x2(y0, y1), v[2 * j + 1] = pack_bf16x2(y2, y3);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(s_outb + stg64_off(lane, j), make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // fp32: 4 rows x 128 B per pass
            const int r = i * 4 + crow;
            const uint4 o = lds128(s_outf + stg128_off(r, cpiece));
            if (grow0 + r < p.M) *reinterpret_cast<uint4*>(xo + (size_t)(grow0 + r) * p.N + ck * 32 + cpiece * 4) = o;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // bf16: 8 rows x 64 B per pass
            const int r = i * 8 + (lane >> 2), pc = lane & 3;
            const uint4 o = lds128(s_outb + stg64_off(r, pc));
            if (grow0 + r < p.M) *reinterpret_cast<uint4*>(xbo + (size_t)(grow0 + r) * p.N + ck * 32 + pc * 8) = o;
          }
          __syncwarp();
        }
      } else if constexpr (epi_staged_bf16(EPI)) {
        // ---- bias (+ activation) -> bf16, transposed through the warp-private staging tile ----
        const uint32_t wst = smem_u32(s_stage) + warp * GEMM_STG_WARP;
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          const int cl = ch * HALF + ck * 32;  // column within the tile
          const int cg = n0 + cl;              // global column
          if (cg >= p.N) break;                // warp-uniform (N is a multiple of 64)
          uint32_t v[32];
          tmem_ld32(taddr + ck * 32, v);
          tc_wait_ld();
          float y[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b4 = *reinterpret_cast<const float4*>(s_bias + cl + j * 4);
            y[4 * j] = __uint_as_float(v[4 * j]) + b4.x, y[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b4.y;
            y[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b4.z, y[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b4.w;
          }
          if constexpr (EPI == EPI_ADD_SILU_BF16) {
            const float* ar = p.addmat + (size_t)(row_ok ? row : 0) * p.N + cg;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 a4 = *reinterpret_cast<const float4*>(ar + j);
              y[j] = silu(y[j] + a4.x), y[j + 1] = silu(y[j + 1] + a4.y);
              y[j + 2] = silu(y[j + 2] + a4.z), y[j + 3] = silu(y[j + 3] + a4.w);
            }
          } else if constexpr (EPI == EPI_BIAS_GELU_BF16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = gelu_erf(y[j]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j)
            sts128(wst + stg64_off(lane, j),
                   make_uint4(pack_bf16x2(y[8 * j], y[8 * j + 1]), pack_bf16x2(y[8 * j + 2], y[8 * j + 3]),
                              pack_bf16x2(y[8 * j + 4], y[8 * j + 5]), pack_bf16x2(y[8 * j + 6], y[8 * j + 7])));
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {  // 8 rows x 64 B per pass
            const int r = i * 8 + (lane >> 2), pc = lane & 3;
            const uint4 o = lds128(wst + stg64_off(r, pc));
            if (grow0 + r < p.M) *reinterpret_cast<uint4*>(p.out_bf16 + (size_t)(grow0 + r) * p.ld_bf16 + cg + pc * 8) = o;
          }
          __syncwarp();
        }
      } else {
        // ---- small-N / remapped epilogues: row-per-thread stores ----
#pragma unroll 1
        for (int ck = 0; ck < CHUNKS; ++ck) {
          uint32_t v[32];
          __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores below
          tmem_ld32(taddr + ck * 32, v);
          tc_wait_ld();
          const int cl = ch * HALF + ck * 32;  // column within the tile
          const int c0 = n0 + cl;              // global column of v[0]
          if constexpr (EPI == EPI_POSTERIOR) {
            const int b = row / p.S, s = row % p.S;
            if (row_ok && s >= p.P0 && c0 < p.nfeat) {
              const int tau = s - p.P0;
              const int t = p.t_ptr[b];
              const float k1 = p.c1[t], k2 = p.c2[t], sg = p.sigma[t];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int f = c0 + j;
                if (f < p.nfeat) {
                  const size_t e = ((size_t)b * p.nfeat + f) * p.T + tau;
                  const float x0 = nan_to_num(__uint_as_float(v[j]) + s_bias[cl + j]);
                  if (p.x0_out) p.x0_out[e] = x0;
                  if (p.x_out) {
                    const float eps = p.noise ? p.noise[e] : philox_normal(p.seed, (uint32_t)t, e);
                    p.x_out[e] = (k1 * x0 + k2 * p.x_t[e]) + sg * eps;
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
          } else if (row_ok && c0 < p.N) {
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(v[j]) + s_bias[cl + j];
            if constexpr (EPI == EPI_F32) {
              float* o = p.out_f32 + (size_t)row * p.ld_f32 + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
            } else {
              static_assert(EPI == EPI_F32 || EPI == EPI_TOKEN_OUT || EPI == EPI_POSTERIOR || EPI == EPI_RESIDUAL_OUT,
                            "unhandled epilogue");
              const int b = row / p.T, tau = row % p.T;
              const size_t orow = (size_t)b * p.S + p.P0 + tau;
              const float* per = p.pe + (size_t)(p.P0 + tau) * p.N + c0;
              float* xo = p.X + orow * p.N + c0;
              __nv_bfloat16* xb = p.Xb + orow * p.N + c0;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
#pragma unroll
                for (int e = 0; e < 8; e += 4) {
                  const float4 pe4 = *reinterpret_cast<const float4*>(per + j + e);
                  y[j + e + 0] = nan_to_num(y[j + e + 0]) + pe4.x, y[j + e + 1] = nan_to_num(y[j + e + 1]) + pe4.y;
                  y[j + e + 2] = nan_to_num(y[j + e + 2]) + pe4.z, y[j + e + 3] = nan_to_num(y[j + e + 3]) + pe4.w;
                  *reinterpret_cast<float4*>(xo + j + e) = make_float4(y[j + e], y[j + e + 1], y[j + e + 2], y[j + e + 3]);
                }
                *reinterpret_cast<uint4*>(xb + j) =
                    make_uint4(pack_bf16x2(y[j], y[j + 1]), pack_bf16x2(y[j + 2], y[j + 3]),
                               pack_bf16x2(y[j + 4], y[j + 5]), pack_bf16x2(y[j + 6], y[j + 7]));
              }
            }
          }
        }
      }
      // accumulator stage drained -> hand it back to the MMA warp (and, for LN, the ring to the producer)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == ACC) acc = 0, acc_phase ^= 1u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// host-side launcher ------------------------------------------------------------------------------
int num_sms();
bool pdl_enabled();  // TAMF_PDL=0 turns programmatic dependent launch off (debug aid)

template <int BN, int EPI>
int configure_gemm() {  // once per process, outside any stream capture
  TAMF_CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GemmSmem<BN, EPI>::BYTES));
  return TAMF_OK;
}

template <int BN, int EPI>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
  const int tiles = ((p.M + GEMM_BM - 1) / GEMM_BM) * ((p.N + BN - 1) / BN);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (epi_staged_bf16(EPI)) {
    TAMF_REQUIRE(p.N % 64 == 0, TAMF_E_BADARG, "gemm: N must be a multiple of 64 for the bf16 epilogues");
  }
  if (EPI == EPI_RES_LN) {
    TAMF_REQUIRE(p.N == BN, TAMF_E_BADARG, "gemm: the LayerNorm epilogue needs the whole row in one tile (N == BN)");
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = GemmSmem<BN, EPI>::BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, EPI>, tmA, tmB, p);
  count_launch();
  if (e != cudaSuccess) {
    set_error(std::string("gemm launch failed: ") + cudaGetErrorString(e));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

}  // namespace tamf
