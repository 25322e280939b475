// tma_bw.cu -- microbenchmark (not product code): per-SM TMA ingest bandwidth from L2-resident data on B200.
// Each CTA streams [128 x 64] bf16 boxes (16 KB, 128B swizzle) from a small L2-resident tensor into a ring of
// `stages` x `boxes_per_stage` buffers; one thread issues, one thread waits.  Reports GB/s per SM and per chip.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_bw tma_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W;\n}" ::"r"(s32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)),
               "l"((uint64_t)m), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int rows, int kcols, int stages, int boxes, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024 - (s32(raw) & 1023)) & 1023);
  uint64_t* full = (uint64_t*)(sm + stages * boxes * 16384);
  uint64_t* empty = full + stages;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int row_tiles = rows / 128, ktiles = kcols / 64;
  if (threadIdx.x == 0) {  // producer
    uint32_t st = 0, ph = 0;
    int r = (blockIdx.x * 7) % row_tiles, k = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&empty[st], ph ^ 1u);
      mbar_expect(&full[st], boxes * 16384);
      for (int b = 0; b < boxes; ++b) {
        tma2d(sm + (st * boxes + b) * 16384, &tm, &full[st], k * 64, r * 128);
        if (++k == ktiles) k = 0, r = (r + 1) % row_tiles;
      }
      if (++st == stages) st = 0, ph ^= 1u;
    }
  } else if (threadIdx.x == 32) {  // consumer
    uint32_t st = 0, ph = 0;
    for (int i = 0; i < iters; ++i) {
      mbar_wait(&full[st], ph);
      mbar_arrive(&empty[st]);
      if (++st == stages) st = 0, ph ^= 1u;
    }
  }
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  Enc enc = (Enc)fn;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  for (int mb : {8, 64}) {  // tensor size in MB (L2 resident: 126 MB L2)
    int kcols = 512, rows = mb * 1024 * 1024 / (kcols * 2);
    void* d;
    cudaMalloc(&d, (size_t)rows * kcols * 2);
    cudaMemset(d, 1, (size_t)rows * kcols * 2);
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows};
    cuuint64_t str[1] = {(cuuint64_t)kcols * 2};
    cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int grid : {1, 8, 74, 83, 148}) {
      for (int stages : {2, 4}) {
        for (int boxes : {1, 3, 5}) {
          if (stages * boxes * 16384 > 200 * 1024) continue;
          int iters = 4000 / boxes;
          size_t smem = 1024 + (size_t)stages * boxes * 16384 + 256;
          probe<<<grid, 64, smem>>>(tm, rows, kcols, stages, boxes, 200);
          cudaEvent_t a, b;
          cudaEventCreate(&a), cudaEventCreate(&b);
          cudaEventRecord(a);
          probe<<<grid, 64, smem>>>(tm, rows, kcols, stages, boxes, iters);
          cudaEventRecord(b);
          cudaEventSynchronize(b);
          float ms;
          cudaEventElapsedTime(&ms, a, b);
          double bytes = (double)grid * iters * boxes * 16384;
          printf("tensor %3d MB grid %3d stages %d x %d boxes (%3d KB in flight): %7.1f GB/s per SM, %8.1f GB/s chip  (%s)\n", mb, grid,
                 stages, boxes, stages * boxes * 16, bytes / ms / 1e6 / grid, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
      }
    }
    cudaFree(d);
  }
  return 0;
}
