// common.cuh -- error handling, launch accounting and sm_100a PTX wrappers (mbarrier, TMA, tcgen05/TMEM).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/tamf_b200.h"

namespace tamf {

// ---------------------------------------------------------------------------------------------
// host side: error state + launch counter
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define TAMF_CUDA_CHECK(expr)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::tamf::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
      return TAMF_E_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

#define TAMF_REQUIRE(cond, code, msg)                                                              \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      ::tamf::set_error(std::string(msg) + " [" #cond "]");                                        \
      return (code);                                                                               \
    }                                                                                              \
  } while (0)

#define TAMF_LAUNCH_CHECK()                                                                        \
  do {                                                                                             \
    ::tamf::count_launch();                                                                        \
    cudaError_t _e = cudaGetLastError();                                                           \
    if (_e != cudaSuccess) {                                                                       \
      ::tamf::set_error(std::string("kernel launch failed: ") + cudaGetErrorString(_e));           \
      return TAMF_E_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

int check_device();  // TAMF_E_ARCH unless the current device is sm_100
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// cuTensorMapEncodeTiled resolved through the runtime (no link-time libcuda dependency)
int make_tmap_2d_bf16(CUtensorMap* out, const void* gptr, uint64_t inner /*elements*/, uint64_t rows,
                      uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_rows);
// fp32 row-major tensor, box [box_rows][32 floats] (128-byte rows, 128-byte swizzle): residual-stream tiles
int make_tmap_2d_f32(CUtensorMap* out, const void* gptr, uint64_t inner /*elements*/, uint64_t rows,
                     uint64_t row_pitch_bytes, uint32_t box_rows);

// bf16 [outer][mid][inner] tensor (strides in bytes), box {64, box_mid, 1}, 128-byte swizzle
int make_tmap_3d_bf16(CUtensorMap* out, const void* gptr, uint64_t inner, uint64_t mid, uint64_t outer,
                      uint64_t mid_pitch_bytes, uint64_t outer_pitch_bytes, uint32_t box_mid, uint32_t box_inner = 64);
bool pdl_enabled();  // TAMF_PDL=0 turns programmatic dependent launch off (debug aid)
void pdl_block(int delta);
struct PdlBlock {  // launches inside the scope carry no programmatic-serialization attribute (api.cu)
  bool on;
  explicit PdlBlock(bool enable) : on(enable) { if (on) pdl_block(+1); }
  ~PdlBlock() { if (on) pdl_block(-1); }
};

// ---------------------------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

// torch.nan_to_num defaults: nan -> 0, +-inf -> +-FLT_MAX
__device__ __forceinline__ float nan_to_num(float v) {
  if (v != v) return 0.f;
  if (v > 3.402823466e+38f) return 3.402823466e+38f;
  if (v < -3.402823466e+38f) return -3.402823466e+38f;
  return v;
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tamf: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// One lane of a converged warp (elect.sync).  The warp-specialised issue loops run on ALL lanes so that shared-memory
// addresses, descriptors and coordinates stay warp-uniform (uniform registers); only the TMA / MMA instruction itself
// sits behind the elected lane.  A loop entered by a single lane (`if (lane == 0)`) makes every operand per-thread,
// and each UTCHMMA / UTMALDG is then wrapped in an ELECT + R2UR waterfall of ~20 dependent instructions.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- TMA ----
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c_inner, int c_row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_row)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_u32(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar, int c_inner, int c_row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c_inner), "r"(c_row)
      : "memory");
}
// TMA store of a shared-memory tile (bulk async-group completion); rows / columns outside the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c_inner, int c_row) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c_inner), "r"(c_row)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the N most recent bulk groups of this thread have finished READING their shared-memory source
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// ... have completed (writes performed)
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 3D variants (attention: [batch][token][feature] views; tokens past S are zero-filled on load, clipped on store)
__device__ __forceinline__ void tma_load_3d_u32(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// In-graph kernel timing (debug, tamf_denoiser_profile_graph): kt[0] = earliest CTA entry, kt[1] = earliest end of a CTA's
// dependency wait, kt[2] = latest CTA exit, all in globaltimer nanoseconds (one clock for every SM).  Null in product.
__device__ __forceinline__ unsigned long long tamf_globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void ktime_entry(long long* kt) {
  if (kt) atomicMin(reinterpret_cast<unsigned long long*>(kt), tamf_globaltimer());
}
__device__ __forceinline__ void ktime_ready(long long* kt) {
  if (kt) atomicMin(reinterpret_cast<unsigned long long*>(kt + 1), tamf_globaltimer());
}
__device__ __forceinline__ void ktime_exit(long long* kt) {
  if (kt) atomicMax(reinterpret_cast<unsigned long long*>(kt + 2), tamf_globaltimer());
}

// Programmatic dependent launch (PDL) controls
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.wait::ld that also names the destination registers of the load it completes, so that the compiler cannot
// schedule a use of them above the wait when other work sits between the load and the wait.
__device__ __forceinline__ void tc_wait_ld_dep(uint32_t (&v)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]),
                 "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]),
                 "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
               :
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T ; bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- 2-CTA (cta_group::2) variants: the CTA pair of a cluster issues one 256-row MMA from the leader CTA ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_cluster(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // relaxed: a release at cluster scope compiles to MEMBAR.ALL.GPU (waits for every outstanding global access of the
  // thread).  The only hand-off through this barrier is TMEM (reads completed by tcgen05.wait::ld + tcgen05.fence).
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of a CTA pair: data lands in the issuing CTA's shared memory, bytes are credited to `bar_cluster_addr`
// (the leader CTA's mbarrier).
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c_inner,
                                                int c_row) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c_inner), "r"(c_row)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows per CTA] * B[N/2 rows per CTA]^T, M = 256; issued by ONE thread of the leader.
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs once all prior MMAs of this thread completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// Instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major (cute/arch/mma_sm100_desc.hpp InstrDescriptor).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Same with B MN-major (bit 16): B tile stored [K rows][N contiguous].
__host__ __device__ constexpr uint32_t umma_idesc_bf16_bmn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 16); }

// Shared-memory matrix descriptor for a K-major tile whose rows are exactly one 128-byte swizzle row
// (64 bf16), written by TMA with CU_TENSOR_MAP_SWIZZLE_128B: 8-row groups are 1024 B apart (SBO),
// LBO unused for swizzled K-major, version=1 (Blackwell), layout_type=2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <- lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,"
      "%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,"
      "%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Residual-stream element -> its two bf16 planes: hi = bf16(v), lo = bf16(v - hi)  (see gemm.cuh, EPI_RES_LN)
__device__ __forceinline__ void store_hilo(__nv_bfloat16* hi_plane, __nv_bfloat16* lo_plane, size_t i, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi_plane[i] = h;
  lo_plane[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// Philox4x32-10 (Salmon et al. 2011) -- counter-based, used for the sampler's eps.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}
// Standard normal for flat element index e at diffusion step t: element e uses word (e & 3) pairs of the
// Philox block with counter (e >> 2, t, 0, 0), key = seed; Box-Muller on (u0,u1) / (u2,u3).
__device__ __forceinline__ float philox_normal(uint64_t seed, uint32_t t, uint64_t e) {
  uint64_t blk = e >> 2;
  uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), t, 0u),
                          make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  uint32_t sel = (uint32_t)(e & 3u);
  uint32_t a = (sel < 2) ? r.x : r.z, b = (sel < 2) ? r.y : r.w;
  float u0 = ((float)a + 0.5f) * 2.3283064365386963e-10f;  // (0,1)
  float u1 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  float rad = sqrtf(-2.0f * __logf(u0));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  return (sel & 1u) ? rad * s : rad * c;
}

// Sampler noise inside the fused posterior epilogue: one Philox block = 4 standard normals for the 4 consecutive
// features 4g..4g+3 of frame (b*T + tau) at diffusion step t.  Counter (frame.lo, frame.hi, t, g | 0x80000000),
// key = seed (the high bit of the last word keeps this stream disjoint from philox_normal's).
__device__ __forceinline__ void philox_normal4(uint64_t seed, uint32_t t, uint64_t frame, uint32_t g, float (&n)[4]) {
  const uint4 r = philox4x32_10(make_uint4((uint32_t)frame, (uint32_t)(frame >> 32), t, g | 0x80000000u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const float u0 = ((float)r.x + 0.5f) * 2.3283064365386963e-10f, u1 = ((float)r.y + 0.5f) * 2.3283064365386963e-10f;
  const float u2 = ((float)r.z + 0.5f) * 2.3283064365386963e-10f, u3 = ((float)r.w + 0.5f) * 2.3283064365386963e-10f;
  const float ra = sqrtf(-2.0f * __logf(u0)), rb = sqrtf(-2.0f * __logf(u2));
  float s, c;
  __sincosf(6.283185307179586f * u1, &s, &c);
  n[0] = ra * c, n[1] = ra * s;
  __sincosf(6.283185307179586f * u3, &s, &c);
  n[2] = rb * c, n[3] = rb * s;
}

#endif  // __CUDACC__
}  // namespace tamf
