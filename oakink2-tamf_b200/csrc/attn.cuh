// attn.cuh -- fused frame-sequence self-attention, softmax(Q K^T / sqrt(hd)) V, no mask (the reference applies none:
// nn.TransformerEncoderLayer built at interaction_segment_mdm.py:63-70, applied :171).
//
// S <= 176 tokens (5 prefix + 160 frames), so a whole (sequence, head) fits one CTA's shared memory and the softmax is
// single-pass.  Grid (q-tiles of 64 rows, heads, batch); 4 warps, each owns 16 query rows:
//   scores 16 x 176 in registers (fp32) from bf16 mma.m16n8k16 over ldmatrix fragments of Q and K,
//   fp32 max / exp2 / sum with quad shuffles, probabilities re-packed in registers as the A operand of P.V,
//   V fragments by ldmatrix.trans.  Q/K/V rows are read straight out of the packed in_proj output [M, 3d]
//   with 16-byte cp.async into XOR-swizzled shared memory (conflict-free ldmatrix).
// Tensor work here is 0.6 % of a layer's FLOPs; it stays on the warp-level MMA path.
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int HD>
__global__ void __launch_bounds__(128) attn_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                                   int S, int d) {
  constexpr int ROWB = HD * 2;      // bytes per smem row
  constexpr int CPR = ROWB / 16;    // 16-byte chunks per row
  extern __shared__ __align__(128) uint8_t att_smem[];
  const uint32_t sK = smem_u32(att_smem);
  const uint32_t sV = sK + ATT_KP * ROWB;
  const uint32_t sQ = sV + ATT_KP * ROWB;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const size_t ld = (size_t)3 * d;
  const __nv_bfloat16* base = qkv + (size_t)b * S * ld + (size_t)h * HD;

  // ---- stage Q tile, K, V (rows >= S are zero) ----
  for (int i = tid; i < ATT_KP * CPR; i += 128) {
    const int r = i / CPR, c = i % CPR;
    const uint32_t off = r * ROWB + ((c ^ (r & 7)) << 4);
    if (r < S) {
      cp_async16(sK + off, base + (size_t)r * ld + d + c * 8);
      cp_async16(sV + off, base + (size_t)r * ld + 2 * d + c * 8);
    } else {
      *reinterpret_cast<uint4*>(att_smem + off) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(att_smem + ATT_KP * ROWB + off) = make_uint4(0, 0, 0, 0);
    }
  }
  for (int i = tid; i < ATT_QT * CPR; i += 128) {
    const int r = i / CPR, c = i % CPR, gr = qt * ATT_QT + r;
    const uint32_t off = r * ROWB + ((c ^ (r & 7)) << 4);
    if (gr < S)
      cp_async16(sQ + off, base + (size_t)gr * ld + c * 8);
    else
      *reinterpret_cast<uint4*>(att_smem + 2 * ATT_KP * ROWB + off) = make_uint4(0, 0, 0, 0);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (qt * ATT_QT + warp * 16 >= S) return;  // warp owns no valid query row (no block-wide sync below)

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
      ldsm_x4(sQ + row * ROWB + ((c ^ (row & 7)) << 4), a0, a1, a2, a3);
    }
#pragma unroll
    for (int np = 0; np < ATT_NT / 2; ++np) {
      const int key = np * 16 + (mi >> 1) * 8 + r8, c = kk * 2 + (mi & 1);
      uint32_t b0, b1, b2, b3;
      ldsm_x4(sK + key * ROWB + ((c ^ (key & 7)) << 4), b0, b1, b2, b3);
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
      ldsm_x4_t(sV + key * ROWB + ((c ^ (key & 7)) << 4), b0, b1, b2, b3);
      mma_bf16_16816(o[2 * hp], a0, a1, a2, a3, b0, b1);
      mma_bf16_16816(o[2 * hp + 1], a0, a1, a2, a3, b2, b3);
    }
  }
  // ---- normalise + store bf16 [M, d] ----
  const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
  const int g = lane >> 2;
  const int row0 = qt * ATT_QT + warp * 16 + g, row1 = row0 + 8;
  __nv_bfloat16* ob = out + (size_t)b * S * d + (size_t)h * HD;
#pragma unroll
  for (int n = 0; n < HD / 8; ++n) {
    const int col = n * 8 + t2;
    if (row0 < S) *reinterpret_cast<uint32_t*>(ob + (size_t)row0 * d + col) = pack_bf16x2(o[n][0] * inv0, o[n][1] * inv0);
    if (row1 < S) *reinterpret_cast<uint32_t*>(ob + (size_t)row1 * d + col) = pack_bf16x2(o[n][2] * inv1, o[n][3] * inv1);
  }
}

template <int HD>
int configure_attn() {
  TAMF_CUDA_CHECK(cudaFuncSetAttribute(attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (2 * ATT_KP + ATT_QT) * HD * 2));
  return TAMF_OK;
}

template <int HD>
int launch_attn(const __nv_bfloat16* qkv, __nv_bfloat16* out, int B, int S, int H, int d, cudaStream_t stream) {
  constexpr int SMEM = (2 * ATT_KP + ATT_QT) * HD * 2;
  dim3 grid((S + ATT_QT - 1) / ATT_QT, H, B);
  attn_kernel<HD><<<grid, 128, SMEM, stream>>>(qkv, out, S, d);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

}  // namespace tamf
