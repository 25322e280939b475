// encoder.cu -- shared transformer encoder stack (see encoder.cuh) + fp32 conditioning helpers.
#include "encoder.cuh"

#include "attn_tc.cuh"
#include "gemm.cuh"
#include "layer_chain.cuh"

namespace tamf {

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------

// fp32 SIMT linear: 64x64 tile, 16-wide k slab, 4x4 outputs per thread.
__global__ void __launch_bounds__(256)
    linear_f32_kernel(const float* __restrict__ in, int ld_in, const float* __restrict__ W, int ld_w,
                      const float* __restrict__ bias, float* __restrict__ out, int ld_out, int R, int N, int K, int post,
                      const float* __restrict__ add, int ld_add) {
  __shared__ float sI[16][65], sW[16][65];
  const int r0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      const int rr = i / 16, kk = i % 16;
      sI[kk][rr] = (r0 + rr < R && k0 + kk < K) ? in[(size_t)(r0 + rr) * ld_in + k0 + kk] : 0.f;
      sW[kk][rr] = (n0 + rr < N && k0 + kk < K) ? W[(size_t)(n0 + rr) * ld_w + k0 + kk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sI[kk][ty * 4 + i], b[i] = sW[kk][tx * 4 + i];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty * 4 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (post == 1) v = v / (1.0f + expf(-v));
      if (post == 2) v = nan_to_num(v) + add[(size_t)r * ld_add + n];
      out[(size_t)r * ld_out + n] = v;
    }
  }
}

int linear_f32(const float* in, int ld_in, const float* W, int ld_w, const float* bias, float* out, int ld_out, int R,
               int N, int K, int post, const float* add, int ld_add, cudaStream_t s) {
  dim3 grid((N + 63) / 64, (R + 63) / 64);
  linear_f32_kernel<<<grid, 256, 0, s>>>(in, ld_in, W, ld_w, bias, out, ld_out, R, N, K, post, add, ld_add);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

__global__ void mean_axis_kernel(const float* __restrict__ in, float* __restrict__ out, int outer, int red, int inner) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)outer * inner) return;
  const int o = (int)(i / inner), c = (int)(i % inner);
  float acc = 0.f;
  for (int r = 0; r < red; ++r) acc += in[((size_t)o * red + r) * inner + c];
  out[i] = acc / (float)red;
}

int mean_axis(const float* in, float* out, int outer, int red, int inner, cudaStream_t s) {
  const size_t n = (size_t)outer * inner;
  mean_axis_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, out, outer, red, inner);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

__global__ void traj_mean_kernel(const float* __restrict__ traj, float* __restrict__ out, int B, int nobj, int T) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * T * 9) return;
  const int c = (int)(i % 9), tau = (int)((i / 9) % T), b = (int)(i / (9 * (size_t)T));
  float acc = 0.f;
  for (int o = 0; o < nobj; ++o) acc += traj[(((size_t)b * nobj + o) * T + tau) * 9 + c];
  out[i] = acc / (float)nobj;
}

int traj_mean(const float* traj, float* out, int B, int nobj, int T, cudaStream_t s) {
  const size_t n = (size_t)B * T * 9;
  traj_mean_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(traj, out, B, nobj, T);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

__global__ void to_bf16_pad_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows, int cols,
                                   int ld_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * ld_out) return;
  const int c = (int)(i % ld_out), r = (int)(i / ld_out);
  out[i] = __float2bfloat16_rn(c < cols ? in[(size_t)r * cols + c] : 0.f);
}

__global__ void fill_int_kernel(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

int fill_int(int* p, int n, int v, cudaStream_t s) {
  fill_int_kernel<<<(n + 255) / 256, 256, 0, s>>>(p, n, v);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

// ------------------------------------------------------------------------------------------------
// DevPool
// ------------------------------------------------------------------------------------------------
int DevPool::alloc(void** p, size_t bytes) {
  TAMF_CUDA_CHECK(cudaMalloc(p, bytes));
  owned.push_back(*p);
  return TAMF_OK;
}

int DevPool::upload_f32(float** dst, const float* src, size_t n) {
  TAMF_REQUIRE(src != nullptr, TAMF_E_BADARG, "create: null weight pointer");
  int rc = alloc((void**)dst, n * sizeof(float));
  if (rc) return rc;
  TAMF_CUDA_CHECK(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
  return TAMF_OK;
}

int DevPool::upload_bf16(__nv_bfloat16** dst, const float* src, int rows, int cols, int ld) {
  TAMF_REQUIRE(src != nullptr, TAMF_E_BADARG, "create: null weight pointer");
  float* tmp = nullptr;
  TAMF_CUDA_CHECK(cudaMalloc(&tmp, (size_t)rows * cols * sizeof(float)));
  cudaError_t e = cudaMemcpy(tmp, src, (size_t)rows * cols * sizeof(float), cudaMemcpyHostToDevice);
  int rc = (e == cudaSuccess) ? alloc((void**)dst, (size_t)rows * ld * sizeof(__nv_bfloat16)) : TAMF_E_CUDA;
  if (rc == TAMF_OK) {
    const size_t n = (size_t)rows * ld;
    to_bf16_pad_kernel<<<(unsigned)((n + 255) / 256), 256>>>(tmp, *dst, rows, cols, ld);
    count_launch();
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = TAMF_E_CUDA;
  }
  cudaFree(tmp);
  if (rc == TAMF_E_CUDA) set_error(std::string("upload_bf16: ") + cudaGetErrorString(cudaGetLastError()));
  return rc;
}

void DevPool::free_all() {
  for (void* p : owned) cudaFree(p);
  owned.clear();
}

// ------------------------------------------------------------------------------------------------
// encoder stack
// ------------------------------------------------------------------------------------------------
int EncoderStack::upload(DevPool& pool, const tamf_layer_weights* w, int d_, int ff_, int L_, int H_) {
  TAMF_REQUIRE(d_ == 256 || d_ == 512, TAMF_E_BADARG, "latent_dim must be 256 or 512 (arch_mdm / arch_mdm_l / arch_refine)");
  TAMF_REQUIRE(H_ > 0 && d_ % H_ == 0 && (d_ / H_ == 64 || d_ / H_ == 128), TAMF_E_BADARG, "head_dim must be 64 or 128");
  TAMF_REQUIRE(ff_ % 256 == 0 && ff_ >= 256, TAMF_E_BADARG, "ff_size must be a multiple of 256");
  TAMF_REQUIRE(L_ > 0 && L_ <= 64 && w, TAMF_E_BADARG, "bad num_layers");
  d = d_, ff = ff_, L = L_, H = H_;
  layers.resize(L);
  int rc;
#define TRY(x) \
  if ((rc = (x)) != TAMF_OK) return rc;
  for (int l = 0; l < L; ++l) {
    const tamf_layer_weights& s = w[l];
    LayerDev& o = layers[l];
    TRY(pool.upload_bf16(&o.w_in, s.in_proj_w, 3 * d, d, d));
    TRY(pool.upload_bf16(&o.w_out, s.out_proj_w, d, d, d));
    TRY(pool.upload_bf16(&o.w1, s.lin1_w, ff, d, d));
    TRY(pool.upload_bf16(&o.w2, s.lin2_w, d, ff, ff));
    TRY(pool.upload_f32(&o.b_in, s.in_proj_b, 3 * d));
    TRY(pool.upload_f32(&o.b_out, s.out_proj_b, d));
    TRY(pool.upload_f32(&o.b1, s.lin1_b, ff));
    TRY(pool.upload_f32(&o.b2, s.lin2_b, d));
    TRY(pool.upload_f32(&o.g1, s.norm1_w, d));
    TRY(pool.upload_f32(&o.be1, s.norm1_b, d));
    TRY(pool.upload_f32(&o.g2, s.norm2_w, d));
    TRY(pool.upload_f32(&o.be2, s.norm2_b, d));
    // every encoder GEMM runs as a CTA pair (CG = 2): BK = 64, each CTA stages UN/2 = 128 rows of W per box
    const uint32_t wbox = (uint32_t)gemm_b_box_rows(256, 2);
    TRY(make_tmap_2d_bf16(&o.tm_in, o.w_in, d, 3 * d, (uint64_t)d * 2, 64, wbox));
    TRY(make_tmap_2d_bf16(&o.tm_out, o.w_out, d, d, (uint64_t)d * 2, 64, wbox));
    TRY(make_tmap_2d_bf16(&o.tm_w1, o.w1, d, ff, (uint64_t)d * 2, 64, wbox));
    TRY(make_tmap_2d_bf16(&o.tm_w2, o.w2, ff, d, (uint64_t)ff * 2, 64, wbox));
  }
  if (L <= CH_MAX_STACK_LAYERS) {  // device copies for the stack form of the layer kernel
    std::vector<CUtensorMap> maps((size_t)L * 4);
    std::vector<LayerWeightPtrs> ptrs(L);
    for (int l = 0; l < L; ++l) {
      const LayerDev& o = layers[l];
      const LayerDev& nx = layers[l + 1 < L ? l + 1 : l];  // (the last layer has no INP units)
      maps[l * 4 + CK_LN1] = o.tm_out, maps[l * 4 + CK_L1] = o.tm_w1, maps[l * 4 + CK_LN2] = o.tm_w2, maps[l * 4 + CK_INP] = nx.tm_in;
      ptrs[l].bias[CK_LN1] = o.b_out, ptrs[l].bias[CK_L1] = o.b1, ptrs[l].bias[CK_LN2] = o.b2, ptrs[l].bias[CK_INP] = nx.b_in;
      ptrs[l].gamma[0] = o.g1, ptrs[l].beta[0] = o.be1, ptrs[l].gamma[1] = o.g2, ptrs[l].beta[1] = o.be2;
    }
    TRY(pool.alloc((void**)&d_wmaps, maps.size() * sizeof(CUtensorMap)));
    TRY(pool.alloc((void**)&d_lw, ptrs.size() * sizeof(LayerWeightPtrs)));
    TAMF_CUDA_CHECK(cudaMemcpy(d_wmaps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    TAMF_CUDA_CHECK(cudaMemcpy(d_lw, ptrs.data(), ptrs.size() * sizeof(LayerWeightPtrs), cudaMemcpyHostToDevice));
  }
#undef TRY
  return TAMF_OK;
}

// aux layout: [counters | statistics words | schedule (with next in_proj) | schedule (last layer)], 256-byte aligned
struct ChainLayout {
  int M, tiles_m, halves;
  size_t ctr_words, stats_words, off_ctr, off_stats, off_sched, off_schedL, off_schedS, sched_bytes, schedS_bytes, total;
};
static ChainLayout chain_layout(int M, int d, int ff) {
  ChainLayout L{};
  L.M = M;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  L.tiles_m = (M + 255) / 256;
  L.halves = d / CH_BN;
  L.ctr_words = (size_t)6 * L.tiles_m + (size_t)L.tiles_m * 256 + 8;  // 6 per-row-tile counters + one word per sequence (<= rows)
  L.stats_words = (size_t)2 * L.tiles_m * L.halves * 2 * 4 * 128;
  L.sched_bytes = al(((size_t)num_sms() / 2 + 1 + (size_t)L.tiles_m * (2 * L.halves + (ff + 3 * d) / CH_BN)) * 4);
  size_t o = 0;
  L.off_ctr = o, o += al(L.ctr_words * 4);
  L.off_stats = o, o += al(L.stats_words * 8);
  L.off_sched = o, o += L.sched_bytes;
  L.off_schedL = o, o += L.sched_bytes;
  L.schedS_bytes = al(((size_t)num_sms() / 2 + 1 +
                       (size_t)CH_MAX_STACK_LAYERS * L.tiles_m * (2 * L.halves + (ff + 3 * d) / CH_BN)) * 4);
  L.off_schedS = o, o += L.schedS_bytes;
  L.total = o;
  return L;
}

static LayerCosts layer_costs_from_env() {
  // unit cost estimates in cycles (profiles/r02_*_timelines.txt), overridable for schedule experiments
  LayerCosts c;
  auto env_d = [](const char* k, double v) { return getenv(k) ? atof(getenv(k)) : v; };
  c.kb = env_d("TAMF_CHAIN_KB", c.kb), c.res = env_d("TAMF_CHAIN_RES", c.res);
  c.epi_ln = env_d("TAMF_CHAIN_EPI_LN", c.epi_ln), c.ln_ready = env_d("TAMF_CHAIN_LN_READY", c.ln_ready);
  c.epi_gelu = env_d("TAMF_CHAIN_EPI_GELU", c.epi_gelu), c.epi_bias = env_d("TAMF_CHAIN_EPI_BIAS", c.epi_bias);
  c.signal = env_d("TAMF_CHAIN_SIGNAL", c.signal), c.slack = env_d("TAMF_CHAIN_SLACK", c.slack);
  return c;
}

static int upload_schedule(int* dst, const LayerSchedule& sc, size_t cap_bytes, cudaStream_t stream) {
  TAMF_REQUIRE((size_t)(sc.pairs + 1 + sc.units.size()) * 4 <= cap_bytes, TAMF_E_BADARG, "layer schedule overflow");
  TAMF_CUDA_CHECK(cudaMemcpyAsync(dst, sc.off.data(), (size_t)(sc.pairs + 1) * 4, cudaMemcpyHostToDevice, stream));
  TAMF_CUDA_CHECK(cudaMemcpyAsync(dst + sc.pairs + 1, sc.units.data(), sc.units.size() * 4, cudaMemcpyHostToDevice, stream));
  TAMF_CUDA_CHECK(cudaStreamSynchronize(stream));  // the host vectors may go out of scope
  return TAMF_OK;
}

size_t encoder_aux_bytes(int M, int d, int ff) { return chain_layout(M, d, ff).total; }

int EncoderBuffers::make_maps(int d, int ff, int layers, int heads) {
  int rc;
  if ((rc = make_tmap_2d_bf16(&tm_Xb, Xb, d, M, (uint64_t)d * 2, 64, 128))) return rc;
  // A operands of the two LayerNorm GEMMs: 32-row boxes (one per TMEM lane quarter, ln_rq rows apart; gemm_ln_rq())
  ln_rq = gemm_ln_rq(M);
  if ((rc = make_tmap_2d_bf16(&tm_ATT, ATT, d, M, (uint64_t)d * 2, 64, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_H, Hb, ff, M, (uint64_t)ff * 2, 64, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_QKV_st, QKV, 3 * d, M, (uint64_t)3 * d * 2, 64, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_H_st, Hb, ff, M, (uint64_t)ff * 2, 64, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_Xb_st, Xb, d, M, (uint64_t)d * 2, 32, (uint32_t)ln_rq))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_Xlo, Xlo, d, M, (uint64_t)d * 2, 32, (uint32_t)ln_rq))) return rc;
  AttnTcMaps at;
  if ((rc = make_attn_tc_maps(&at, QKV, ATT, B, S, d))) return rc;
  tm_att_kv = at.kv, tm_att_o = at.o;
  // ---- layer kernel ----
  static const bool env_chain = !(getenv("TAMF_CHAIN") && getenv("TAMF_CHAIN")[0] == '0');
  chain = env_chain && aux != nullptr;
  if (!chain) return TAMF_OK;
  if ((rc = make_tmap_2d_bf16(&tm_ATT128, ATT, d, M, (uint64_t)d * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_H128, Hb, ff, M, (uint64_t)ff * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_Xlo128, Xlo, d, M, (uint64_t)d * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_Xh_st, Xb, d, M, (uint64_t)d * 2, 32, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_Xl_st, Xlo, d, M, (uint64_t)d * 2, 32, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_H_st32, Hb, ff, M, (uint64_t)ff * 2, 32, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tm_QKV_st32, QKV, 3 * d, M, (uint64_t)3 * d * 2, 32, 32))) return rc;
  if ((rc = chain_identity_map(&tm_ident))) return rc;
  const ChainLayout lay = chain_layout(M, d, ff);
  tiles_m = lay.tiles_m, halves = lay.halves;
  uint8_t* a = static_cast<uint8_t*>(aux);
  ctr = reinterpret_cast<unsigned*>(a + lay.off_ctr);
  stats = reinterpret_cast<unsigned long long*>(a + lay.off_stats);
  sched = reinterpret_cast<int*>(a + lay.off_sched);
  schedL = reinterpret_cast<int*>(a + lay.off_schedL);
  const LayerCosts costs = layer_costs_from_env();
  const int slots = num_sms() / 2;
  const LayerSchedule sc = build_layer_schedule(M, d, ff, 3 * d, slots, costs);
  const LayerSchedule sl = build_layer_schedule(M, d, ff, 0, slots, costs);
  pairs = sc.pairs, pairsL = sl.pairs;
  if ((rc = upload_schedule(sched, sc, lay.sched_bytes, nullptr)) || (rc = upload_schedule(schedL, sl, lay.sched_bytes, nullptr)))
    return rc;
  // ---- stack form: one launch for all layers + one persistent attention launch ----
  static const bool env_stack = getenv("TAMF_CHAIN") && getenv("TAMF_CHAIN")[0] == '2';
  stack = env_stack && layers >= 2 && layers <= CH_MAX_STACK_LAYERS && heads > 0 && num_sms() >= 16;
  if (stack) {
    static const int env_att = getenv("TAMF_STACK_ATT") ? atoi(getenv("TAMF_STACK_ATT")) : 28;
    // (splits with more than 28 attention CTAs fail with a launch failure on the B200s of this pool -- not understood; the
    // form is experimental and off by default, so the range is clamped to what has been validated: 20, 24, 28)
    att_ctas = std::max(8, std::min(env_att & ~1, 28));
    static const double env_unit = getenv("TAMF_STACK_ATT_UNIT") ? atof(getenv("TAMF_STACK_ATT_UNIT")) : 9600.0;
    AttnModel am;
    am.ctas = att_ctas, am.S = S, am.heads = heads, am.unit = env_unit;
    const LayerSchedule ss = build_stack_schedule(M, d, ff, layers, false, (num_sms() - att_ctas) / 2, costs, am);
    schedS = reinterpret_cast<int*>(a + lay.off_schedS);
    pairsS = ss.pairs;
    if ((rc = upload_schedule(schedS, ss, lay.schedS_bytes, nullptr))) return rc;
    if (!side) TAMF_CUDA_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    if (!ev_fork) TAMF_CUDA_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    if (!ev_join) TAMF_CUDA_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
  }
  TAMF_CUDA_CHECK(cudaMemset(ctr, 0, lay.ctr_words * 4));          // (cleared again at the start of every evaluation)
  ctr_bytes = lay.ctr_words * 4;
  TAMF_CUDA_CHECK(cudaMemset(stats, 0xFF, lay.stats_words * 8));   // "not posted"; every reader resets its word
  return TAMF_OK;
}

void EncoderBuffers::release() {
  if (side) cudaStreamDestroy(side);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  side = nullptr, ev_fork = nullptr, ev_join = nullptr;
}

int configure_encoder_kernels() {
  int rc;
  if ((rc = configure_gemm<256, EPI_BIAS_BF16, 2>())) return rc;
  if ((rc = configure_gemm<256, EPI_BIAS_GELU_BF16, 2>())) return rc;
  if ((rc = configure_gemm<256, EPI_RES_LN, 2>())) return rc;
  if ((rc = configure_gemm<512, EPI_RES_LN, 2>())) return rc;
  if ((rc = configure_attn_tc<64>())) return rc;
  if ((rc = configure_attn_tc<128>())) return rc;
  if ((rc = configure_layer_chain())) return rc;
  return TAMF_OK;
}

// debug only (tools/chain_trace_model.py): per-CTA timeline of the layer kernel of one layer inside a real step
static long long* g_dbg_trace = nullptr;
static int g_dbg_trace_layer = -1;

static void fill_layer_params(LayerParams& p, const EncoderStack& enc, const EncoderBuffers& buf, int l) {
  const LayerDev& w = enc.layers[l];
  const bool last = l + 1 == enc.L;
  p.M = buf.M, p.d = enc.d, p.ff = enc.ff, p.n_inp = last ? 0 : 3 * enc.d;
  p.bias[CK_LN1] = w.b_out, p.bias[CK_L1] = w.b1, p.bias[CK_LN2] = w.b2;
  p.bias[CK_INP] = last ? nullptr : enc.layers[l + 1].b_in;
  p.gamma[0] = w.g1, p.beta[0] = w.be1, p.gamma[1] = w.g2, p.beta[1] = w.be2;
  const int* sc = last ? buf.schedL : buf.sched;
  const int pairs = last ? buf.pairsL : buf.pairs;
  p.sched_off = sc, p.sched = sc + pairs + 1;
  p.ctr = buf.ctr, p.stats = buf.stats, p.tiles_m = buf.tiles_m;
  p.launch_idx = l + 1, p.B = buf.B, p.S = buf.S, p.heads = enc.H;
  p.target_ln = (unsigned)(buf.halves * 2 * GEMM_EPI_WARPS);
  p.target_h = (unsigned)((enc.ff / CH_BN) * 2 * GEMM_EPI_WARPS);
  p.target_att = (unsigned)(enc.H * ((buf.S + 31) / 32));
  static const int env_ooo = getenv("TAMF_CHAIN_OOO") ? atoi(getenv("TAMF_CHAIN_OOO")) : 0;
  p.ooo = std::max(0, std::min(env_ooo, 24));  // run-time unit selection window (0: list order)
}

// The dependency counters of an evaluation start from zero.  Called ahead of the FIRST kernel of the evaluation (so that the
// programmatic launch chain between the kernels stays unbroken); everything older has completed (stream order).
int encoder_begin_evaluation(const EncoderBuffers& buf, cudaStream_t s) {
  if (buf.chain) TAMF_CUDA_CHECK(cudaMemsetAsync(buf.ctr, 0, buf.ctr_bytes, s));
  return TAMF_OK;
}

// Layer-kernel form of the stack (layer_chain.cuh): in_proj(0), then per layer  attention | everything up to the next
// layer's in_proj  = 1 + 2 L kernels.
static int enqueue_encoder_chain(const EncoderStack& enc, const EncoderBuffers& buf, cudaStream_t s,
                                 std::vector<cudaEvent_t>* marks, long long* ktime, int* kidx) {
  const int d = enc.d, M = buf.M;
  int rc;
  auto kt = [&]() -> long long* { return ktime ? ktime + 4 * ((*kidx)++) : nullptr; };
  unsigned* r_q = buf.ctr + 5 * buf.tiles_m;   // QKV tiles of a row tile stored (layer kernel, INP units)
  unsigned* r_att = buf.ctr + 6 * buf.tiles_m;  // attention CTAs of a sequence done
  const unsigned t_q = (unsigned)((3 * d / CH_BN) * 2 * GEMM_EPI_WARPS);
  {
    const LayerDev& w = enc.layers[0];
    GemmParams p{};
    p.M = M, p.N = 3 * d, p.K = d, p.bias = w.b_in, p.out_bf16 = buf.QKV, p.ld_bf16 = 3 * d, p.tmC = &buf.tm_QKV_st;
    p.ktime = kt();
    if ((rc = launch_gemm<256, EPI_BIAS_BF16, 2>(buf.tm_Xb, w.tm_in, p, s))) return rc;
    mark_event(marks, s);
  }
  static const int dbg = getenv("TAMF_CHAIN_DBG") ? atoi(getenv("TAMF_CHAIN_DBG")) : 0;
  if (buf.stack && enc.d_wmaps) {
    // ---- stack form: ONE persistent attention kernel (att_ctas CTAs, second stream) next to ONE persistent layer kernel
    // (pairsS CTA pairs) for all layers; units of both walk the layers in order and meet only through the counters.
    // Both grids are launched as clusters of 2 (whole TPCs) and together hold <= every SM once, so all CTAs of both are
    // co-resident whatever the placement order: every agent's unit list is a subsequence of one topological order. ----
    TAMF_CUDA_CHECK(cudaEventRecord(buf.ev_fork, s));
    TAMF_CUDA_CHECK(cudaStreamWaitEvent(buf.side, buf.ev_fork, 0));
    {
      AttnTcMaps at;
      at.kv = buf.tm_att_kv, at.o = buf.tm_att_o;
      long long* k = kt();
      rc = (d / enc.H == 128) ? launch_attn_tc<128>(at, buf.B, buf.S, enc.H, d, buf.side, nullptr, k, r_q, t_q, r_att, enc.L, buf.att_ctas)
                              : launch_attn_tc<64>(at, buf.B, buf.S, enc.H, d, buf.side, nullptr, k, r_q, t_q, r_att, enc.L, buf.att_ctas);
      if (rc) return rc;
      mark_event(marks, s);  // (profiling: keeps one mark per kernel; the attention kernel runs on the side stream)
    }
    {
      LayerParams p{};
      fill_layer_params(p, enc, buf, 0);
      p.n_inp = 3 * d;
      p.sched_off = buf.schedS, p.sched = buf.schedS + buf.pairsS + 1;
      p.wmaps = enc.d_wmaps, p.lw = enc.d_lw;
      p.dbg = dbg;
      p.ktime = kt();
      if (g_dbg_trace_layer >= 0) p.trace = g_dbg_trace;
      const LayerDev& w = enc.layers[0];
      LayerMaps tm{&buf.tm_ATT128, &w.tm_out, &buf.tm_Xb, &buf.tm_Xlo128, &w.tm_w1, &buf.tm_H_st32, &buf.tm_H128, &w.tm_w2,
                   &enc.layers[1].tm_in, &buf.tm_QKV_st32, &buf.tm_Xh_st, &buf.tm_Xl_st, &buf.tm_ident};
      if ((rc = launch_layer_chain(tm, p, buf.pairsS, s))) return rc;
    }
    TAMF_CUDA_CHECK(cudaEventRecord(buf.ev_join, buf.side));
    TAMF_CUDA_CHECK(cudaStreamWaitEvent(s, buf.ev_join, 0));
    mark_event(marks, s);
    return TAMF_OK;
  }
  // debug: TAMF_FINE bit 0 = attention waits per sequence (else for the whole previous grid), bit 1 = the layer kernel
  // starts without a grid-wide wait
  static const int fine = getenv("TAMF_FINE") ? atoi(getenv("TAMF_FINE")) : 3;
  for (int l = 0; l < enc.L; ++l) {
    const LayerDev& w = enc.layers[l];
    {
      AttnTcMaps at;
      at.kv = buf.tm_att_kv, at.o = buf.tm_att_o;
      long long* k = kt();
      // layer 0 follows the plain in_proj GEMM (grid-wide dependency); later layers wait per sequence for the in_proj
      // tiles the previous layer kernel stores (l of them per row tile so far)
      const unsigned* rq = (l && (fine & 1)) ? r_q : nullptr;
      const unsigned tq = (unsigned)l * t_q;
      rc = (d / enc.H == 128) ? launch_attn_tc<128>(at, buf.B, buf.S, enc.H, d, s, nullptr, k, rq, tq, r_att)
                              : launch_attn_tc<64>(at, buf.B, buf.S, enc.H, d, s, nullptr, k, rq, tq, r_att);
      if (rc) return rc;
      mark_event(marks, s);
    }
    const bool last = l + 1 == enc.L;
    LayerParams p{};
    fill_layer_params(p, enc, buf, l);
    p.dbg = dbg;
    p.grid_wait = !(fine & 2);
    p.ktime = kt();
    if (l == g_dbg_trace_layer) p.trace = g_dbg_trace;
    LayerMaps tm{&buf.tm_ATT128, &w.tm_out, &buf.tm_Xb, &buf.tm_Xlo128, &w.tm_w1, &buf.tm_H_st32, &buf.tm_H128, &w.tm_w2,
                 last ? nullptr : &enc.layers[l + 1].tm_in, last ? nullptr : &buf.tm_QKV_st32, &buf.tm_Xh_st, &buf.tm_Xl_st,
                 &buf.tm_ident};
    if ((rc = launch_layer_chain(tm, p, last ? buf.pairsL : buf.pairs, s))) return rc;
    mark_event(marks, s);
  }
  return TAMF_OK;
}

int enqueue_encoder(const EncoderStack& enc, const EncoderBuffers& buf, cudaStream_t s,
                    std::vector<cudaEvent_t>* marks, long long* ktime, int* kidx) {
  if (buf.chain) return enqueue_encoder_chain(enc, buf, s, marks, ktime, kidx);
  auto kt = [&]() -> long long* { return ktime ? ktime + 4 * ((*kidx)++) : nullptr; };
  const int d = enc.d, ff = enc.ff, M = buf.M;
  int rc;
  for (int l = 0; l < enc.L; ++l) {
    const LayerDev& w = enc.layers[l];
    {
      GemmParams p{};
      p.M = M, p.N = 3 * d, p.K = d, p.bias = w.b_in, p.out_bf16 = buf.QKV, p.ld_bf16 = 3 * d, p.tmC = &buf.tm_QKV_st;
      p.ktime = kt();
      if ((rc = launch_gemm<256, EPI_BIAS_BF16, 2>(buf.tm_Xb, w.tm_in, p, s))) return rc;
      mark_event(marks, s);
    }
    {
      AttnTcMaps at;
      at.kv = buf.tm_att_kv, at.o = buf.tm_att_o;
      long long* k = kt();
      rc = (d / enc.H == 128) ? launch_attn_tc<128>(at, buf.B, buf.S, enc.H, d, s, nullptr, k)
                              : launch_attn_tc<64>(at, buf.B, buf.S, enc.H, d, s, nullptr, k);
      if (rc) return rc;
    }
    mark_event(marks, s);
    {
      GemmParams p{};
      p.M = M, p.N = d, p.K = d, p.bias = w.b_out, p.Xlo = buf.Xlo, p.Xb = buf.Xb, p.gamma = w.g1, p.beta = w.be1;
      p.tmC = &buf.tm_Xb_st, p.tmX = &buf.tm_Xlo, p.ln_rq = buf.ln_rq;
      p.ktime = kt();
      rc = (d == 512) ? launch_gemm<512, EPI_RES_LN, 2>(buf.tm_ATT, w.tm_out, p, s)
                      : launch_gemm<256, EPI_RES_LN, 2>(buf.tm_ATT, w.tm_out, p, s);
      if (rc) return rc;
      mark_event(marks, s);
    }
    {
      GemmParams p{};
      p.M = M, p.N = ff, p.K = d, p.bias = w.b1, p.out_bf16 = buf.Hb, p.ld_bf16 = ff, p.tmC = &buf.tm_H_st;
      p.ktime = kt();
      if ((rc = launch_gemm<256, EPI_BIAS_GELU_BF16, 2>(buf.tm_Xb, w.tm_w1, p, s))) return rc;
      mark_event(marks, s);
    }
    {
      GemmParams p{};
      p.M = M, p.N = d, p.K = ff, p.bias = w.b2, p.Xlo = buf.Xlo, p.Xb = buf.Xb, p.gamma = w.g2, p.beta = w.be2;
      p.tmC = &buf.tm_Xb_st, p.tmX = &buf.tm_Xlo, p.ln_rq = buf.ln_rq;
      p.ktime = kt();
      rc = (d == 512) ? launch_gemm<512, EPI_RES_LN, 2>(buf.tm_H, w.tm_w2, p, s)
                      : launch_gemm<256, EPI_RES_LN, 2>(buf.tm_H, w.tm_w2, p, s);
      if (rc) return rc;
      mark_event(marks, s);
    }
  }
  return TAMF_OK;
}

}  // namespace tamf

// Debug aids / self-test (tools/chain_trace_model.py, tools/layer_trace.py, tests/test_gemm_gpu.py)
extern "C" int tamf_debug_chain_trace(long long* trace, int layer) {
  tamf::g_dbg_trace = trace;
  tamf::g_dbg_trace_layer = trace ? layer : -1;
  return TAMF_OK;
}

extern "C" size_t tamf_layer_aux_bytes(int M, int d, int ff) { return tamf::encoder_aux_bytes(M, d, ff); }

// ONE launch of the layer kernel on caller data: X (two bf16 planes, in place), att [M,d], weights as nn.Linear stores
// them (bf16), ln_params = [b_out | g1 | be1 | b2 | g2 | be2] (6 d floats), b1 [ff], b_in [3d] (n_inp = 0: no INP).
extern "C" int tamf_layer_run(const uint16_t* att, const uint16_t* w_out, const uint16_t* w1, const uint16_t* w2,
                              const uint16_t* w_in, const float* ln_params, const float* b1, const float* b_in,
                              uint16_t* Xh, uint16_t* Xl, uint16_t* Hbuf, uint16_t* qkv, int M, int d, int ff, int n_inp,
                              void* aux, size_t aux_bytes, long long* trace, void* stream_) {
  using namespace tamf;
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = check_device();
  if (rc) return rc;
  TAMF_REQUIRE(att && w_out && w1 && w2 && ln_params && b1 && Xh && Xl && Hbuf && aux, TAMF_E_BADARG,
               "tamf_layer_run: null pointer");
  TAMF_REQUIRE((d == 256 || d == 512) && ff % 256 == 0 && ff > 0 && (n_inp == 0 || n_inp == 3 * d), TAMF_E_BADARG,
               "tamf_layer_run: bad shape");
  TAMF_REQUIRE(n_inp == 0 || (w_in && b_in && qkv), TAMF_E_BADARG, "tamf_layer_run: the in_proj stage needs w_in / b_in / qkv");
  const ChainLayout lay = chain_layout(M, d, ff);
  TAMF_REQUIRE(aux_bytes >= lay.total, TAMF_E_BADARG, "tamf_layer_run: aux too small (tamf_layer_aux_bytes)");
  if ((rc = configure_layer_chain())) return rc;
  CUtensorMap tATT, tWo, tXh, tXl, tW1, tHst, tH, tW2, tWin, tQst, tXhs, tXls, tI;
  const uint32_t wbox = (uint32_t)gemm_b_box_rows(256, 2);
  const uint64_t pd = (uint64_t)d * 2, pf = (uint64_t)ff * 2;
  if ((rc = make_tmap_2d_bf16(&tATT, att, d, M, pd, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tWo, w_out, d, d, pd, 64, wbox))) return rc;
  if ((rc = make_tmap_2d_bf16(&tXh, Xh, d, M, pd, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tXl, Xl, d, M, pd, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tW1, w1, d, ff, pd, 64, wbox))) return rc;
  if ((rc = make_tmap_2d_bf16(&tHst, Hbuf, ff, M, pf, 32, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tH, Hbuf, ff, M, pf, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&tW2, w2, ff, d, pf, 64, wbox))) return rc;
  if ((rc = make_tmap_2d_bf16(&tXhs, Xh, d, M, pd, 32, 32))) return rc;
  if ((rc = make_tmap_2d_bf16(&tXls, Xl, d, M, pd, 32, 32))) return rc;
  if ((rc = chain_identity_map(&tI))) return rc;
  if (n_inp) {
    if ((rc = make_tmap_2d_bf16(&tWin, w_in, d, n_inp, pd, 64, wbox))) return rc;
    if ((rc = make_tmap_2d_bf16(&tQst, qkv, n_inp, M, (uint64_t)n_inp * 2, 32, 32))) return rc;
  }
  const LayerSchedule sc = build_layer_schedule(M, d, ff, n_inp, num_sms() / 2, layer_costs_from_env());
  uint8_t* a = static_cast<uint8_t*>(aux);
  int* sched = reinterpret_cast<int*>(a + lay.off_sched);
  if ((rc = upload_schedule(sched, sc, lay.sched_bytes, stream))) return rc;
  unsigned* ctr = reinterpret_cast<unsigned*>(a + lay.off_ctr);
  unsigned long long* stats = reinterpret_cast<unsigned long long*>(a + lay.off_stats);
  TAMF_CUDA_CHECK(cudaMemsetAsync(ctr, 0, lay.ctr_words * 4, stream));
  TAMF_CUDA_CHECK(cudaMemsetAsync(stats, 0xFF, lay.stats_words * 8, stream));
  const unsigned one = 1u;  // the attention output is given: one "sequence" covering every row, announced by one "CTA"
  TAMF_CUDA_CHECK(cudaMemcpyAsync(ctr + 6 * lay.tiles_m, &one, 4, cudaMemcpyHostToDevice, stream));
  TAMF_CUDA_CHECK(cudaStreamSynchronize(stream));
  LayerParams p{};
  p.M = M, p.d = d, p.ff = ff, p.n_inp = n_inp;
  p.launch_idx = 1, p.B = 1, p.S = M > 0 ? M : 1, p.heads = 1;
  p.bias[CK_LN1] = ln_params, p.gamma[0] = ln_params + d, p.beta[0] = ln_params + 2 * d;
  p.bias[CK_LN2] = ln_params + 3 * d, p.gamma[1] = ln_params + 4 * d, p.beta[1] = ln_params + 5 * d;
  p.bias[CK_L1] = b1, p.bias[CK_INP] = b_in;
  p.sched_off = sched, p.sched = sched + sc.pairs + 1;
  p.ctr = ctr, p.stats = stats, p.tiles_m = lay.tiles_m;
  p.target_ln = (unsigned)(lay.halves * 2 * GEMM_EPI_WARPS), p.target_h = (unsigned)((ff / CH_BN) * 2 * GEMM_EPI_WARPS);
  p.target_att = 1u;
  p.trace = trace;
  LayerMaps tm{&tATT, &tWo, &tXh, &tXl, &tW1, &tHst, &tH, &tW2, n_inp ? &tWin : nullptr, n_inp ? &tQst : nullptr, &tXhs,
               &tXls, &tI};
  return launch_layer_chain(tm, p, sc.pairs, stream);
}

// Host-only: the static schedule the layer kernel would run for an [M, d] problem on `slots` CTA pairs (no GPU needed;
// tests/test_host_logic.py checks coverage, duo placement and the global topological order).  off_out [slots + 1],
// units_out [cap] unit codes kind << 28 | row tile << 8 | column tile; returns the pair count (< 0: error).
extern "C" int tamf_layer_schedule(int M, int d, int ff, int n_inp, int slots, int* off_out, int* units_out, int cap,
                                   double* makespan_out) {
  using namespace tamf;
  TAMF_REQUIRE(M > 0 && (d == 256 || d == 512) && ff > 0 && ff % 256 == 0 && n_inp % 256 == 0 && slots >= 1 && off_out &&
                   units_out,
               TAMF_E_BADARG, "tamf_layer_schedule: bad argument");
  const LayerSchedule sc = build_layer_schedule(M, d, ff, n_inp, slots, layer_costs_from_env());
  TAMF_REQUIRE((int)sc.units.size() <= cap, TAMF_E_BADARG, "tamf_layer_schedule: units_out too small");
  for (int i = 0; i <= sc.pairs; ++i) off_out[i] = sc.off[i];
  for (size_t i = 0; i < sc.units.size(); ++i) units_out[i] = sc.units[i];
  if (makespan_out) *makespan_out = sc.makespan;
  return sc.pairs;
}

// Host-only: the schedule of the STACK form (all `layers` layers in one launch on `slots` pairs, attention on `att_ctas`
// CTAs with `att_unit` cycles per (sequence, head) unit).  Unit codes: chain_code (kind << 28 | layer << 24 | m << 8 | n).
extern "C" int tamf_stack_schedule(int M, int d, int ff, int layers, int slots, int att_ctas, int S, int heads,
                                   double att_unit, int* off_out, int* units_out, int cap, double* makespan_out) {
  using namespace tamf;
  TAMF_REQUIRE(M > 0 && (d == 256 || d == 512) && ff > 0 && ff % 256 == 0 && layers >= 1 && layers <= CH_MAX_STACK_LAYERS &&
                   slots >= 1 && off_out && units_out && S > 0 && heads > 0,
               TAMF_E_BADARG, "tamf_stack_schedule: bad argument");
  AttnModel am;
  am.ctas = att_ctas, am.S = S, am.heads = heads, am.unit = att_unit > 0 ? att_unit : am.unit;
  const LayerSchedule sc = build_stack_schedule(M, d, ff, layers, false, slots, layer_costs_from_env(), am);
  TAMF_REQUIRE((int)sc.units.size() <= cap, TAMF_E_BADARG, "tamf_stack_schedule: units_out too small");
  for (int i = 0; i <= sc.pairs; ++i) off_out[i] = sc.off[i];
  for (size_t i = 0; i < sc.units.size(); ++i) units_out[i] = sc.units[i];
  if (makespan_out) *makespan_out = sc.makespan;
  return sc.pairs;
}
