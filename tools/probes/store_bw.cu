// store_bw.cu -- microbenchmark (not product code): chip / per-SM global WRITE bandwidth into an L2-resident region on
// B200, (a) TMA stores of [32 x 128 B] swizzled tiles from shared memory, (b) coalesced st.global.v4 from registers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bw store_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(512, 1) tma_store_probe(const __grid_constant__ CUtensorMap tm, int rows, int cols, int iters, int box_bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024 - (s32(raw) & 1023)) & 1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // each warp owns 2 staging tiles of 4 KB and streams [32 rows x 32 fp32] boxes over its own row range
  const int row_tiles = rows / 32, col_tiles = cols / 32;
  int rt = (blockIdx.x * 16 + warp) % row_tiles, ct = 0;
  for (int i = 0; i < iters; ++i) {
    if (lane == 0) {
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)&tm),
                   "r"(s32(sm + warp * 8192 + (i & 1) * 4096)), "r"(ct * 32), "r"(rt * 32) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (++ct == col_tiles) ct = 0, rt = (rt + gridDim.x * 16) % row_tiles;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void __launch_bounds__(512, 1) stg_probe(float4* out, size_t n4, int iters) {
  // every warp writes 4 full 128-byte lines per instruction, grid-stride
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float4 v = make_float4(1.f, 2.f, 3.f, (float)threadIdx.x);
  for (int it = 0; it < iters; ++it) {
    out[i % n4] = v;
    i += stride;
  }
}
__global__ void __launch_bounds__(512, 1) ldg_probe(const float4* in, size_t n4, int iters, float* sink) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (int it = 0; it < iters; it += 4) {
    float4 a = in[i % n4], b = in[(i + stride) % n4], c = in[(i + 2 * stride) % n4], d = in[(i + 3 * stride) % n4];
    acc += a.x + b.y + c.z + d.w;
    i += 4 * stride;
  }
  if (acc == 1.2345f) *sink = acc;
}
typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Enc enc = (Enc)fn;
  cudaFuncSetAttribute(tma_store_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  const int cols = 512, rows = 16384;  // 32 MB fp32, L2 resident
  float* d;
  cudaMalloc(&d, (size_t)rows * cols * 4);
  cudaMemset(d, 0, (size_t)rows * cols * 4);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t str[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
  enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  cudaEvent_t a, b;
  cudaEventCreate(&a), cudaEventCreate(&b);
  float ms;
  for (int grid : {1, 16, 42, 84, 148}) {
    int iters = 2000;
    tma_store_probe<<<grid, 512, 1024 + 16 * 8192>>>(tm, rows, cols, 100, 4096);
    cudaEventRecord(a);
    tma_store_probe<<<grid, 512, 1024 + 16 * 8192>>>(tm, rows, cols, iters, 4096);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    double bytes = (double)grid * 16 * iters * 4096;
    printf("TMA store  grid %3d: %7.1f GB/s per SM, %8.1f GB/s chip (%s)\n", grid, bytes / ms / 1e6 / grid, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    iters = 4000;
    size_t n4 = (size_t)rows * cols / 4;
    stg_probe<<<grid, 512>>>((float4*)d, n4, 100);
    cudaEventRecord(a);
    stg_probe<<<grid, 512>>>((float4*)d, n4, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    bytes = (double)grid * 512 * iters * 16;
    printf("st.global  grid %3d: %7.1f GB/s per SM, %8.1f GB/s chip (%s)\n", grid, bytes / ms / 1e6 / grid, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    ldg_probe<<<grid, 512>>>((const float4*)d, n4, 100, d);
    cudaEventRecord(a);
    ldg_probe<<<grid, 512>>>((const float4*)d, n4, iters, d);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    printf("ld.global  grid %3d: %7.1f GB/s per SM, %8.1f GB/s chip (%s)\n", grid, bytes / ms / 1e6 / grid, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
