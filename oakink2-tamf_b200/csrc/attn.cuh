// attn.cuh -- fused frame-sequence self-attention, softmax(Q K^T / sqrt(hd)) V, no mask (the reference applies none:
// nn.TransformerEncoderLayer built at interaction_segment_mdm.py:63-70, applied :171).
//
// S <= 176 tokens (5 prefix + 160 frames), so a whole (sequence, head) fits one CTA's shared memory and the softmax is
// single-pass.  Grid (q-tiles of 64 rows, heads, batch); 4 warps, each owns 16 query rows:
//   K, V (176 x hd) and the Q tile (64 x hd) arrive as TMA boxes of a [B][S][3d] view of the packed in_proj output
//   (one elected thread, one mbarrier; tokens >= S are zero-filled by the tensor map, so no guard code and no LSU
//   traffic -- the cp.async version of this kernel was bound by the ~27 B/clk/SM of the LSU global-load path),
//   128-byte-swizzled 64-column blocks = conflict-free ldmatrix;
//   scores 16 x 176 in registers (fp32) from bf16 mma.m16n8k16, fp32 max / exp2 / sum with quad shuffles,
//   probabilities re-packed in registers as the A operand of P.V, V fragments by ldmatrix.trans;
//   the normalised output is staged in the warp's own (dead) Q rows and leaves as TMA stores of a [B][S][d] view
//   (rows >= S clipped).
// Launched with programmatic dependent launch: descriptor prefetch / barrier init overlap the in_proj tail.
#pragma once
#include "common.cuh"

namespace tamf {

constexpr int ATT_KP = 176;  // padded key count (multiple of 16)
constexpr int ATT_QT = 64;   // query rows per CTA
constexpr int ATT_NT = ATT_KP / 8;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Byte offset of 16-byte chunk c (0 .. hd/8-1) of row r inside an operand stored as hd/64 blocks of [ROWS][128 B],
// each block written by TMA with CU_TENSOR_MAP_SWIZZLE_128B (chunk index XOR row & 7).
template <int ROWS>
__device__ __forceinline__ uint32_t att_off(int r, int c) {
  return (uint32_t)((c >> 3) * (ROWS * 128) + r * 128 + (((c & 7) ^ (r & 7)) << 4));
}

template <int HD>
struct AttnCfg {
  static constexpr int NB = HD / 64;  // 64-column blocks per operand
  static constexpr int TX_BYTES = (2 * ATT_KP + ATT_QT) * HD * 2;
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + TX_BYTES;
};

// tmKV: QKV viewed [B][S][3d], box {64, ATT_KP, 1}; tmQ: same view, box {64, ATT_QT, 1}; tmO: ATT viewed [B][S][d],
// box {64, 16, 1}.
template <int HD>
__global__ void __launch_bounds__(128)
    attn_kernel(const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmQ,
                const __grid_constant__ CUtensorMap tmO, int S, int d) {
  constexpr int NB = AttnCfg<HD>::NB;
  constexpr int CPR = HD / 8;  // 16-byte chunks per row
  extern __shared__ uint8_t att_raw[];
  __shared__ uint64_t bar;
  const uint32_t sK = (smem_u32(att_raw) + 1023u) & ~1023u;
  const uint32_t sV = sK + ATT_KP * HD * 2;
  const uint32_t sQ = sV + ATT_KP * HD * 2;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    tma_prefetch_desc(&tmKV);
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmO);
    mbar_init(&bar, 1);
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  __syncthreads();
  pdl_launch_dependents();
  if (tid == 0) {
    pdl_wait();  // QKV is the previous kernel's output
    mbar_arrive_expect_tx(&bar, AttnCfg<HD>::TX_BYTES);
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      tma_load_3d_u32(sQ + j * (ATT_QT * 128), &tmQ, &bar, h * HD + j * 64, qt * ATT_QT, b);
      tma_load_3d_u32(sK + j * (ATT_KP * 128), &tmKV, &bar, d + h * HD + j * 64, 0, b);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) tma_load_3d_u32(sV + j * (ATT_KP * 128), &tmKV, &bar, 2 * d + h * HD + j * 64, 0, b);
  }
  if (qt * ATT_QT + warp * 16 >= S) return;  // warp owns no valid query row (no block-wide sync below)
  mbar_wait(&bar, 0);

  // ---- scores = Q K^T ----
  float sc[ATT_NT][4];
#pragma unroll
  for (int n = 0; n < ATT_NT; ++n) sc[n][0] = sc[n][1] = sc[n][2] = sc[n][3] = 0.f;
  const int mi = lane >> 3, r8 = lane & 7;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t a0, a1, a2, a3;
    {
      const int row = warp * 16 + (lane & 15), c = kk * 2 + (lane >> 4);
      ldsm_x4(sQ + att_off<ATT_QT>(row, c), a0, a1, a2, a3);
    }
#pragma unroll
    for (int np = 0; np < ATT_NT / 2; ++np) {
      const int key = np * 16 + (mi >> 1) * 8 + r8, c = kk * 2 + (mi & 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sK + att_off<ATT_KP>(key, c), b0, b1, b2, b3);
      mma_bf16_16816(sc[2 * np], a0, a1, a2, a3, b0, b1);
      mma_bf16_16816(sc[2 * np + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  // ---- softmax over keys (fp32), rows g = lane/4 and g+8 ----
  const float scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
  const int t2 = (lane & 3) * 2;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int n = 0; n < ATT_NT; ++n) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int key = n * 8 + t2 + (e & 1);
      if (key >= S) sc[n][e] = -INFINITY;
    }
    mx0 = fmaxf(mx0, fmaxf(sc[n][0], sc[n][1]));
    mx1 = fmaxf(mx1, fmaxf(sc[n][2], sc[n][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int n = 0; n < ATT_NT; ++n) {
    sc[n][0] = exp2f((sc[n][0] - mx0) * scale_log2);
    sc[n][1] = exp2f((sc[n][1] - mx0) * scale_log2);
    sc[n][2] = exp2f((sc[n][2] - mx1) * scale_log2);
    sc[n][3] = exp2f((sc[n][3] - mx1) * scale_log2);
    sum0 += sc[n][0] + sc[n][1];
    sum1 += sc[n][2] + sc[n][3];
  }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

  // ---- O = P V ----
  float o[HD / 8][4];
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
  for (int kt = 0; kt < ATT_KP / 16; ++kt) {
    const uint32_t a0 = pack_bf16x2(sc[2 * kt][0], sc[2 * kt][1]), a1 = pack_bf16x2(sc[2 * kt][2], sc[2 * kt][3]);
    const uint32_t a2 = pack_bf16x2(sc[2 * kt + 1][0], sc[2 * kt + 1][1]),
                   a3 = pack_bf16x2(sc[2 * kt + 1][2], sc[2 * kt + 1][3]);
#pragma unroll
    for (int hp = 0; hp < HD / 16; ++hp) {
      const int key = kt * 16 + (mi & 1) * 8 + r8, c = hp * 2 + (mi >> 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(sV + att_off<ATT_KP>(key, c), b0, b1, b2, b3);
      mma_bf16_16816(o[2 * hp], a0, a1, a2, a3, b0, b1);
      mma_bf16_16816(o[2 * hp + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  // ---- normalise -> bf16 into this warp's own Q rows (its last Q read was the final ldmatrix above) -> TMA store ----
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  const int g = lane >> 2;
  const int r0 = warp * 16 + g, r1 = r0 + 8;
  __syncwarp();
#pragma unroll
  for (int n = 0; n < CPR; ++n) {
    const uint32_t v0 = pack_bf16x2(o[n][0] * inv0, o[n][1] * inv0), v1 = pack_bf16x2(o[n][2] * inv1, o[n][3] * inv1);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + att_off<ATT_QT>(r0, n) + (lane & 3) * 4), "r"(v0) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(sQ + att_off<ATT_QT>(r1, n) + (lane & 3) * 4), "r"(v1) : "memory");
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NB; ++j)
      tma_store_3d(&tmO, sQ + j * (ATT_QT * 128) + warp * (16 * 128), h * HD + j * 64, qt * ATT_QT + warp * 16, b);
    bulk_commit();
    bulk_wait<0>();
  }
}

struct AttnMaps {
  CUtensorMap kv, q, o;
};
// qkv bf16 [B*S, 3d] (packed in_proj output), out bf16 [B*S, d]
inline int make_attn_maps(AttnMaps* m, const void* qkv, const void* out, int B, int S, int d) {
  int rc;
  const uint64_t ld = (uint64_t)3 * d * 2;
  if ((rc = make_tmap_3d_bf16(&m->kv, qkv, 3 * (uint64_t)d, S, B, ld, ld * S, ATT_KP))) return rc;
  if ((rc = make_tmap_3d_bf16(&m->q, qkv, 3 * (uint64_t)d, S, B, ld, ld * S, ATT_QT))) return rc;
  if ((rc = make_tmap_3d_bf16(&m->o, out, d, S, B, (uint64_t)d * 2, (uint64_t)d * 2 * S, 16))) return rc;
  return TAMF_OK;
}

template <int HD>
int configure_attn() {
  TAMF_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       AttnCfg<HD>::SMEM_BYTES));
  return TAMF_OK;
}

template <int HD>
int launch_attn(const AttnMaps& m, int B, int S, int H, int d, cudaStream_t stream) {
  TAMF_REQUIRE(S <= ATT_KP, TAMF_E_BADARG, "attention: at most 176 tokens per sequence");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((S + ATT_QT - 1) / ATT_QT, H, B);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = AttnCfg<HD>::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, attn_kernel<HD>, m.kv, m.q, m.o, S, d);
  count_launch();
  if (e != cudaSuccess) {
    set_error(std::string("attention launch failed: ") + cudaGetErrorString(e));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

}  // namespace tamf
