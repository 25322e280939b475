// api.cu -- library-wide state: version, thread-local error, device check, TMA descriptor encoding,
// GEMM self-test and Philox test entry points.
#include <stdlib.h>

#include <mutex>

#include "attn_tc.cuh"
#include "gemm.cuh"

namespace tamf {

static thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& msg) { g_err = msg; }

int check_device() {
  static thread_local int cached_dev = -1, cached_rc = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    set_error("no CUDA device: libtamf_b200 has no CPU fallback");
    return TAMF_E_CUDA;
  }
  if (dev == cached_dev) return cached_rc;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
    set_error("cudaDeviceGetAttribute failed");
    return TAMF_E_CUDA;
  }
  cached_dev = dev;
  cached_rc = TAMF_OK;
  if (major != 10) {
    set_error("libtamf_b200 is built for sm_100a (B200) only; device has compute capability major " +
              std::to_string(major));
    cached_rc = TAMF_E_ARCH;
  }
  return cached_rc;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// Programmatic dependent launch is on unless TAMF_PDL=0, or the caller holds a PdlBlock (common.cuh): a kernel that must
// not become resident before everything older has finished (the kernel after the two co-resident persistent kernels of
// the encoder's stack form: parked early on free SMs it would starve the one that is launched last).
static thread_local int g_pdl_blocked = 0;
void pdl_block(int delta) { g_pdl_blocked += delta; }
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TAMF_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1 && g_pdl_blocked == 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2D bf16 row-major tensor [rows][inner], TMA box [box_rows][box_inner]: box_inner = 64 -> 128-byte swizzle,
// box_inner = 32 -> 64-byte swizzle (the two K-major operand layouts gemm.cuh builds UMMA descriptors for).
int make_tmap_2d_bf16(CUtensorMap* out, const void* gptr, uint64_t inner, uint64_t rows, uint64_t row_pitch_bytes,
                      uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  TAMF_REQUIRE(enc != nullptr, TAMF_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  TAMF_REQUIRE(aligned16(gptr) && (row_pitch_bytes % 16) == 0, TAMF_E_ALIGN, "TMA tensor must be 16-byte aligned");
  TAMF_REQUIRE((box_inner == 64 || box_inner == 32) && box_rows <= 256, TAMF_E_BADARG,
               "TMA box must be 64 or 32 bf16 x <=256 rows");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* gptr, uint64_t inner, uint64_t mid, uint64_t outer,
                      uint64_t mid_pitch_bytes, uint64_t outer_pitch_bytes, uint32_t box_mid, uint32_t box_inner) {
  EncodeTiledFn enc = get_encode();
  TAMF_REQUIRE(enc != nullptr, TAMF_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  TAMF_REQUIRE(aligned16(gptr) && (mid_pitch_bytes % 16) == 0 && (outer_pitch_bytes % 16) == 0, TAMF_E_ALIGN,
               "TMA tensor must be 16-byte aligned");
  TAMF_REQUIRE(box_mid >= 1 && box_mid <= 256 && (box_inner == 64 || box_inner == 32), TAMF_E_BADARG,
               "TMA box must be 64 or 32 bf16 x <=256 rows");
  cuuint64_t dims[3] = {inner, mid, outer};
  cuuint64_t strides[2] = {mid_pitch_bytes, outer_pitch_bytes};
  cuuint32_t box[3] = {box_inner, box_mid, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3d) failed with CUresult " + std::to_string((int)r));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

int make_tmap_2d_f32(CUtensorMap* out, const void* gptr, uint64_t inner, uint64_t rows, uint64_t row_pitch_bytes,
                     uint32_t box_rows) {
  EncodeTiledFn enc = get_encode();
  TAMF_REQUIRE(enc != nullptr, TAMF_E_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  TAMF_REQUIRE(aligned16(gptr) && (row_pitch_bytes % 16) == 0, TAMF_E_ALIGN, "TMA tensor must be 16-byte aligned");
  TAMF_REQUIRE(box_rows <= 256, TAMF_E_BADARG, "TMA box must be <= 256 rows");
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(gptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult " + std::to_string((int)r));
    return TAMF_E_CUDA;
  }
  return TAMF_OK;
}

__global__ void philox_fill_kernel(float* out, size_t n, unsigned long long seed, uint32_t t) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = philox_normal(seed, t, i);
}

int philox_fill(float* out, size_t n, uint64_t seed, uint32_t t, cudaStream_t stream) {
  philox_fill_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(out, n, seed, t);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

}  // namespace tamf

using namespace tamf;

extern "C" int tamf_version(void) { return 100; }
extern "C" const char* tamf_last_error(void) { return g_err.c_str(); }
extern "C" uint64_t tamf_kernel_launch_count(void) { return g_launches.load(); }

extern "C" int tamf_philox_normal(float* out, size_t n, uint64_t seed, uint32_t t, void* stream) {
  TAMF_REQUIRE(out != nullptr, TAMF_E_BADARG, "tamf_philox_normal: null pointer");
  int rc = check_device();
  if (rc) return rc;
  if (n == 0) return TAMF_OK;
  return philox_fill(out, n, seed, t, (cudaStream_t)stream);
}

namespace tamf {
template <int BN, int CG>
static int selftest_run(const uint16_t* a, const uint16_t* w, const GemmParams& p, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d_bf16(&tmA, a, p.K, p.M, (uint64_t)p.K * 2, gemm_bk(BN, CG), 128);
  if (rc) return rc;
  rc = make_tmap_2d_bf16(&tmB, w, p.K, p.N, (uint64_t)p.K * 2, gemm_bk(BN, CG), gemm_b_box_rows(BN, CG));
  if (rc) return rc;
  if ((rc = configure_gemm<BN, EPI_F32, CG>())) return rc;
  return launch_gemm<BN, EPI_F32, CG>(tmA, tmB, p, stream);
}
}  // namespace tamf

extern "C" int tamf_gemm_selftest(const uint16_t* a, const uint16_t* w, const float* bias, float* c, int M, int N, int K,
                                  int tile_n, int cta_group, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_device();
  if (rc) return rc;
  TAMF_REQUIRE(a && w && c, TAMF_E_BADARG, "tamf_gemm_selftest: null pointer");
  TAMF_REQUIRE(M > 0 && N > 0 && K > 0 && (K % 8) == 0 && (N % 32) == 0, TAMF_E_BADARG,
               "tamf_gemm_selftest: need K % 8 == 0 and N % 32 == 0");
  TAMF_REQUIRE(tile_n == 128 || tile_n == 256 || tile_n == 512, TAMF_E_BADARG, "tamf_gemm_selftest: tile_n");
  TAMF_REQUIRE(cta_group == 1 || cta_group == 2, TAMF_E_BADARG, "tamf_gemm_selftest: cta_group must be 1 or 2");
  GemmParams p{};
  p.M = M, p.N = N, p.K = K, p.bias = bias, p.out_f32 = c, p.ld_f32 = N;
  if (cta_group == 1) {
    if (tile_n == 128) return selftest_run<128, 1>(a, w, p, stream);
    if (tile_n == 256) return selftest_run<256, 1>(a, w, p, stream);
    return selftest_run<512, 1>(a, w, p, stream);
  }
  if (tile_n == 128) return selftest_run<128, 2>(a, w, p, stream);
  if (tile_n == 256) return selftest_run<256, 2>(a, w, p, stream);
  return selftest_run<512, 2>(a, w, p, stream);
}

extern "C" int tamf_attn_selftest(const uint16_t* qkv, uint16_t* out, int B, int S, int H, int d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_device();
  if (rc) return rc;
  TAMF_REQUIRE(qkv && out, TAMF_E_BADARG, "tamf_attn_selftest: null pointer");
  TAMF_REQUIRE(B > 0 && S > 0 && S <= ATC_KP && H > 0 && d % H == 0 && (d / H == 64 || d / H == 128), TAMF_E_BADARG,
               "tamf_attn_selftest: need 0 < S <= 176 and head_dim 64 or 128");
  AttnTcMaps at;
  if ((rc = make_attn_tc_maps(&at, qkv, out, B, S, d))) return rc;
  if (d / H == 128) {
    if ((rc = configure_attn_tc<128>())) return rc;
    return launch_attn_tc<128>(at, B, S, H, d, stream);
  }
  if ((rc = configure_attn_tc<64>())) return rc;
  return launch_attn_tc<64>(at, B, S, H, d, stream);
}

// Debug aid (tools/attn_trace.py): the tcgen05 attention kernel with per-CTA clock64 stamps, trace [H*B][16] int64.
extern "C" int tamf_attn_trace(const uint16_t* qkv, uint16_t* out, int B, int S, int H, int d, long long* trace,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_device();
  if (rc) return rc;
  TAMF_REQUIRE(qkv && out && trace, TAMF_E_BADARG, "tamf_attn_trace: null pointer");
  TAMF_REQUIRE(B > 0 && S > 0 && S <= ATC_KP && H > 0 && d % H == 0 && (d / H == 64 || d / H == 128), TAMF_E_BADARG,
               "tamf_attn_trace: need 0 < S <= 176 and head_dim 64 or 128");
  AttnTcMaps at;
  if ((rc = make_attn_tc_maps(&at, qkv, out, B, S, d))) return rc;
  if (d / H == 128) {
    if ((rc = configure_attn_tc<128>())) return rc;
    return launch_attn_tc<128>(at, B, S, H, d, stream, trace);
  }
  if ((rc = configure_attn_tc<64>())) return rc;
  return launch_attn_tc<64>(at, B, S, H, d, stream, trace);
}

// Debug aid (tools/gemm_trace.py): one launch of the hot-path GEMM shape `which` on caller data with per-CTA
// clock64 timestamps.  which: 0 in_proj-like <256,BIAS_BF16>, 1 linear1-like <256,GELU>, 2 LN <512,RES_LN>.
// trace [148][64] int64 device.  out: bf16 [M,N] (which 0/1); X fp32 + Xb bf16 [M,N] (which 2).
extern "C" int tamf_gemm_trace(int which, const uint16_t* a, const uint16_t* w, const float* bias, void* out, float* X,
                               int M, int N, int K, long long* trace, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_device();
  if (rc) return rc;
  CUtensorMap tmA, tmB;
  if ((rc = make_tmap_2d_bf16(&tmA, a, K, M, (uint64_t)K * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tmB, w, K, N, (uint64_t)K * 2, 64, gemm_b_box_rows(256, 2)))) return rc;
  GemmParams p{};
  p.M = M, p.N = N, p.K = K, p.bias = bias, p.trace = trace;
  p.dbg = getenv("TAMF_GEMM_DBG") ? atoi(getenv("TAMF_GEMM_DBG")) : 0;
  CUtensorMap tmC, tmX;
  if (which == 2) {
    p.ln_rq = gemm_ln_rq(M);
    if ((rc = make_tmap_2d_bf16(&tmA, a, K, M, (uint64_t)K * 2, 64, 32))) return rc;  // one box per TMEM lane quarter
    if ((rc = make_tmap_2d_bf16(&tmC, out, N, M, (uint64_t)N * 2, 32, (uint32_t)p.ln_rq))) return rc;
    if ((rc = make_tmap_2d_bf16(&tmX, X, N, M, (uint64_t)N * 2, 32, (uint32_t)p.ln_rq))) return rc;  // low plane (bf16 view of the buffer)
    p.tmC = &tmC, p.tmX = &tmX;
    p.Xlo = (__nv_bfloat16*)X, p.Xb = (__nv_bfloat16*)out, p.gamma = bias, p.beta = bias;
    if ((rc = configure_gemm<512, EPI_RES_LN, 2>())) return rc;
    return launch_gemm<512, EPI_RES_LN, 2>(tmA, tmB, p, stream);
  }
  if (which == 4) {  // embed-b shape: token epilogue, M = 64 * 160 frame rows -> token rows b*165 + 5 + tau of the pair
                     // out (Xb, [64*165, N] bf16) / X + 4 MB (Xlo); positional encoding read from the head of X
    TAMF_REQUIRE(M == 64 * 160, TAMF_E_BADARG, "tamf_gemm_trace: which = 4 expects M = 10240");
    p.pe = X, p.T = 160, p.S = 165, p.P0 = 5;
    p.Xb = (__nv_bfloat16*)out, p.Xlo = (__nv_bfloat16*)((char*)X + (4u << 20));
    if ((rc = make_token_out_maps(&tmC, &tmX, p.Xb, p.Xlo, 64, p.T, p.S, p.P0, N))) return rc;
    p.tmC = &tmC, p.tmX = &tmX;
    if ((rc = configure_gemm<256, EPI_TOKEN_OUT, 2>())) return rc;
    return launch_gemm<256, EPI_TOKEN_OUT, 2>(tmA, tmB, p, stream);
  }
  p.out_bf16 = (__nv_bfloat16*)out, p.ld_bf16 = N;
  if ((rc = make_tmap_2d_bf16(&tmC, out, N, M, (uint64_t)N * 2, 64, 32))) return rc;
  p.tmC = &tmC;
  if (which == 5) {  // embed-a shape: bias + SiLU -> bf16
    if ((rc = configure_gemm<256, EPI_BIAS_SILU_BF16, 2>())) return rc;
    return launch_gemm<256, EPI_BIAS_SILU_BF16, 2>(tmA, tmB, p, stream);
  }
  if (which == 10) {  // single-CTA form of the in_proj shape (comparison only)
    if ((rc = make_tmap_2d_bf16(&tmB, w, K, N, (uint64_t)K * 2, 64, gemm_b_box_rows(256, 1)))) return rc;
    if ((rc = configure_gemm<256, EPI_BIAS_BF16, 1>())) return rc;
    return launch_gemm<256, EPI_BIAS_BF16, 1>(tmA, tmB, p, stream);
  }
  if (which == 1) {
    if ((rc = configure_gemm<256, EPI_BIAS_GELU_BF16, 2>())) return rc;
    return launch_gemm<256, EPI_BIAS_GELU_BF16, 2>(tmA, tmB, p, stream);
  }
  if ((rc = configure_gemm<256, EPI_BIAS_BF16, 2>())) return rc;
  return launch_gemm<256, EPI_BIAS_BF16, 2>(tmA, tmB, p, stream);
}
