// denoiser.cu -- MF-MDM G denoiser + DDPM ancestral sampler: weights, workspace, conditioning, per-step launch
// sequence and the CUDA-graph chain.
//
// Replaces InterationSegmentMDM.forward (src/oakink2_tamf/model/interaction_segment_mdm.py:134-174) and
// GaussianDiffusion.p_sample / p_sample_loop_progressive (model/diffusion/gaussian_diffusion.py:412-460, 573-640).
//
// Per step (everything else is hoisted, SURVEY.md 8a'):
//   prep        x_t [B,99,1,T] fp32 -> A0[:, 0:100] bf16 (frame-major); token 0 = ttab[t]+pe[0]; tokens 1..4 = prefix
//   embed-a     H0 = silu(A0 . Wf^T + bf)   A0 = [x_t (99) | 0 | mean_obj(obj_traj) (9) | 0..] (K = 128), Wf =
//               [merge0[:, :d] . poseEmbedding | 0 | merge0[:, d:] . objPoseEmbedding | 0..]  (both linear maps of
//               input_merge.0 folded; the trajectory columns of A0 are written once per sample)
//   embed-b     tok[5+tau] = nan_to_num(H0 . merge2^T + b) + pe[5+tau]     -> Xb / Xlo, rows b*S+5+tau
//   8 x layer   QKV = Xb . Win^T + b ; ATT = softmax(QK^T/sqrt(hd)) V ; X = LN1(X + ATT . Wo^T + b)
//               H = gelu(Xb . W1^T + b) ; X = LN2(X + H . W2^T + b)
//   final       x0 = nan_to_num(Xb . Wf^T + b) ; x_{t-1} = c1[t] x0 + c2[t] x_t + sigma[t] eps   (fused epilogue)
// = 2 + 5*L + 2 = 44 kernel launches for L = 8, captured once in a CUDA graph and replayed per step.
//
// HBM layout (row = token index b*S + s, S = 5 + T; all row-major):
//   Xb bf16 [M,d] + Xlo bf16 [M,d] residual stream x = Xb + Xlo (Xb doubles as the GEMM operand) | QKV bf16 [M,3d] |
//   ATT bf16 [M,d] | H bf16 [M,ff]
//   A0 bf16 [B*T,128] | H0 bf16 [B*T,d] | prefix fp32 [B,4,d] | ttab fp32 [steps,d]
#include <vector>

#include "attn_tc.cuh"
#include "encoder.cuh"
#include "gemm.cuh"

namespace tamf {

int philox_fill(float* out, size_t n, uint64_t seed, uint32_t t, cudaStream_t stream);

constexpr int KPAD = 128;      // K of embed-a: 99 pose features | 0 | 9 trajectory features at TRAJ_COL | 0...
constexpr int TRAJ_COL = 100;  // first column of mean_obj(obj_traj) in A0 / Wfold (even: prep writes bf16 pairs)
constexpr int MAX_NOBJ = 8;    // staging capacity of tamf_p_sample_loop_host

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------

// prefix[b,1,:] = hand-side token (rh -> 0, lh -> e0), then nan_to_num(prefix) + pe[1..4]
__global__ void prefix_finish_kernel(float* __restrict__ prefix, const int* __restrict__ hand_side,
                                     const float* __restrict__ pe, int B, int d) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * 4 * d) return;
  const int c = (int)(i % d), s = (int)((i / d) % 4), b = (int)(i / (4 * (size_t)d));
  float v = prefix[i];
  if (s == 1) v = (hand_side[b] == 1 && c == 0) ? 1.f : 0.f;
  prefix[i] = nan_to_num(v) + pe[(size_t)(1 + s) * d + c];
}

// A0[r, TRAJ_COL .. KPAD) = [mean_obj(obj_traj)[r, 0:9], 0 ...] (bf16), once per sample batch
__global__ void a0_cond_kernel(const float* __restrict__ trajmean, __nv_bfloat16* __restrict__ A0, int rows) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * (KPAD - TRAJ_COL)) return;
  const int r = (int)(i / (KPAD - TRAJ_COL)), c = (int)(i % (KPAD - TRAJ_COL));
  A0[(size_t)r * KPAD + TRAJ_COL + c] = __float2bfloat16_rn(c < 9 ? trajmean[(size_t)r * 9 + c] : 0.f);
}

// rows[i] = src[map[i]]  (timestep tokens of a respaced schedule)
__global__ void gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ map, float* __restrict__ dst,
                                   int d) {
  const float* s = src + (size_t)map[blockIdx.x] * d;
  float* o = dst + (size_t)blockIdx.x * d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) o[c] = s[c];
}

// prep: grid (ceil(T/32), B), 256 threads.
//  (a) A0[b*T+tau, k] = bf16(x[b,k,0,tau]) (k < nfeat), 0 for the K padding -- transposed through shared memory so
//      both the read (along tau) and the write (along k) are coalesced.
//  (b) blockIdx.x == 0 also writes the 5 prefix token rows of sequence b into the residual planes Xb / Xlo.
//  (c) chain mode (t_cur != null): the same CTA publishes this step's timestep as t_cur[b] (read by the posterior
//      epilogue at the end of the step) and advances the chain counter t_ptr[b] to t - 1 for the next replay of the
//      step graph -- no other CTA of the step reads t_ptr.
__global__ void __launch_bounds__(256)
    prep_kernel(const float* __restrict__ x, int* __restrict__ t_ptr, int* __restrict__ t_cur,
                const float* __restrict__ ttab, const float* __restrict__ pe, const float* __restrict__ prefix,
                __nv_bfloat16* __restrict__ A0, __nv_bfloat16* __restrict__ Xlo, __nv_bfloat16* __restrict__ Xb, int T,
                int S, int d, int nfeat, long long* ktime) {
  __shared__ float tile[KPAD][33];
  if (threadIdx.x == 0) {
    ktime_entry(ktime);
    ktime_ready(ktime);
  }
  const int b = blockIdx.y, tau0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int k = w; k < nfeat; k += 8) {
    const int tau = tau0 + lane;
    tile[k][lane] = (tau < T) ? x[((size_t)b * nfeat + k) * T + tau] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * (TRAJ_COL / 2); i += 256) {  // columns >= TRAJ_COL belong to set_cond
    const int j = i / (TRAJ_COL / 2), kp = (i % (TRAJ_COL / 2)) * 2;
    const int tau = tau0 + j;
    if (tau < T) {
      const float v0 = kp < nfeat ? tile[kp][j] : 0.f, v1 = (kp + 1) < nfeat ? tile[kp + 1][j] : 0.f;
      *reinterpret_cast<uint32_t*>(A0 + ((size_t)b * T + tau) * KPAD + kp) = pack_bf16x2(v0, v1);
    }
  }
  if (blockIdx.x == 0) {
    const int t = t_ptr[b];
    if (t_cur) {
      __syncthreads();  // every thread of the CTA has read t_ptr[b]
      if (threadIdx.x == 0) t_cur[b] = t, t_ptr[b] = t - 1;
    }
    for (int i = threadIdx.x; i < 5 * d / 2; i += 256) {  // two columns per thread: 4-byte stores into both planes
      const int s = i / (d / 2), c = (i % (d / 2)) * 2;
      float v0, v1;
      if (s == 0) {  // interaction_segment_mdm.py:142,158,170
        const float2 tt = *reinterpret_cast<const float2*>(ttab + (size_t)t * d + c);
        const float2 pp = *reinterpret_cast<const float2*>(pe + c);
        v0 = nan_to_num(tt.x) + pp.x, v1 = nan_to_num(tt.y) + pp.y;
      } else {
        const float2 pf = *reinterpret_cast<const float2*>(prefix + ((size_t)b * 4 + (s - 1)) * d + c);
        v0 = pf.x, v1 = pf.y;
      }
      const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - __bfloat162float(h.x), v1 - __bfloat162float(h.y));
      const size_t o = ((size_t)b * S + s) * d + c;
      *reinterpret_cast<__nv_bfloat162*>(Xb + o) = h;
      *reinterpret_cast<__nv_bfloat162*>(Xlo + o) = l;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) ktime_exit(ktime);
}

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
}  // namespace tamf

using namespace tamf;

struct tamf_denoiser {
  tamf_cfg cfg{};
  int d = 0, ff = 0, L = 0, H = 0, nfeat = 0;
  DevPool pool;      // device allocations freed by destroy
  EncoderStack enc;  // the 8 post-norm layers (encoder.cuh)
  EncoderBuffers buf;
  // fp32 conditioning weights (exact fp32 SIMT path, once per sample)
  float *shape_w, *shape_b, *objemb_w, *objemb_b, *merge_bias /* b1 + W1a.bp + W1b.bo */, *text_w, *text_b, *pe, *ttab;
  int pe_rows = 0;
  // bf16 hot-path weights
  __nv_bfloat16 *wfold /*[d,128]*/, *wm2 /*[d,d]*/, *wfin /*[99,d]*/;
  float *b_m2, *b_fin;
  CUtensorMap tm_wfold, tm_wm2, tm_wfin;
  float *c1, *c2, *sigma;  // [num_steps] ancestral rule of tamf_denoiser_create
  // update rule the sampler entries run (tamf_denoiser_set_sampler): K steps, x_{i-1} = k1[i] x0 + k2[i] x_i + ks[i] eps,
  // timestep token of step i = ttab[timestep_map[i]] gathered into k_ttab.  Defaults to the ancestral rule.
  float *k1 = nullptr, *k2 = nullptr, *ks = nullptr, *k_ttab = nullptr;
  float *alt1 = nullptr, *alt2 = nullptr, *alts = nullptr, *alt_ttab = nullptr;
  int* alt_map = nullptr;
  int K = 0;
  // bound workspace
  int B = 0, T = 0, S = 0, M = 0, Mf = 0;
  bool bound = false, cond_set = false;
  float *prefix, *trajmean, *shapemean, *embmean, *xbuf;
  float *st_text, *st_shape, *st_traj, *st_emb;  // host-API staging
  int *st_side, *t_dev, *t_cur;
  unsigned long long* seed_dev;  // Philox seed of the running chain (read by the captured step: one graph for every seed)
  __nv_bfloat16 *A0, *H0;
  CUtensorMap tm_tok_hi, tm_tok_lo;  // EPI_TOKEN_OUT stores (make_token_out_maps)
  CUtensorMap tm_A0, tm_H0, tm_H0_st, tm_Xb_fin;
  // cached step graph
  cudaGraphExec_t graph_exec = nullptr;
  float* graph_x = nullptr;
  cudaStream_t graph_stream = nullptr;
  int graph_kernels = 0;
};

namespace tamf {

static int dev_alloc(tamf_denoiser* h, void** p, size_t bytes) { return h->pool.alloc(p, bytes); }
static int upload_f32(tamf_denoiser* h, float** dst, const float* src, size_t n) { return h->pool.upload_f32(dst, src, n); }
static int upload_bf16(tamf_denoiser* h, __nv_bfloat16** dst, const float* src, int rows, int cols, int ld) {
  return h->pool.upload_bf16(dst, src, rows, cols, ld);
}

// One denoiser evaluation (+ optional posterior update) enqueued on `s`.
// `marks` (profiling only): one event is recorded after every kernel of the step, in launch order.
// `model_level`: t_ptr holds ORIGINAL timesteps (model(x, t) of the reference); otherwise indices of the installed
// K-step rule (what SpacedDiffusion hands to _WrappedModel, respace.py:114-119).
// `advance` (chain graph): t_ptr is the chain counter h->t_dev; prep copies it to h->t_cur for the rest of the step and
// decrements it for the next replay.
static int enqueue_step(tamf_denoiser* h, const float* x_t, int* t_ptr, float* x_out, float* x0_out,
                        const float* noise, uint64_t seed, cudaStream_t s, std::vector<cudaEvent_t>* marks = nullptr,
                        bool model_level = false, bool advance = false, long long* ktime = nullptr) {
  const int d = h->d, B = h->B, T = h->T, S = h->S, M = h->M, Mf = h->Mf;
  int rc;
  int kidx = 0;  // profiling only: 4 int64 timing slots per kernel, in launch order
  auto kt = [&]() -> long long* { return ktime ? ktime + 4 * (kidx++) : nullptr; };
  if ((rc = encoder_begin_evaluation(h->buf, s))) return rc;
  mark_event(marks, s);
  prep_kernel<<<dim3((T + 31) / 32, B), 256, 0, s>>>(x_t, t_ptr, advance ? h->t_cur : nullptr,
                                                     model_level ? h->ttab : h->k_ttab, h->pe, h->prefix, h->A0,
                                                     h->buf.Xlo, h->buf.Xb, T, S, d, h->nfeat, kt());
  TAMF_LAUNCH_CHECK();
  mark_event(marks, s);
  {  // embed-a
    GemmParams p{};
    p.M = Mf, p.N = d, p.K = KPAD, p.bias = h->merge_bias, p.out_bf16 = h->H0, p.ld_bf16 = d, p.tmC = &h->tm_H0_st;
    p.ktime = kt();
    if ((rc = launch_gemm<256, EPI_BIAS_SILU_BF16, 2>(h->tm_A0, h->tm_wfold, p, s))) return rc;
    mark_event(marks, s);
  }
  {  // embed-b
    GemmParams p{};
    p.M = Mf, p.N = d, p.K = d, p.bias = h->b_m2, p.pe = h->pe, p.T = T, p.S = S, p.P0 = 5, p.Xlo = h->buf.Xlo,
    p.Xb = h->buf.Xb;
    p.tmC = &h->tm_tok_hi, p.tmX = &h->tm_tok_lo;
    p.ktime = kt();
    if ((rc = launch_gemm<256, EPI_TOKEN_OUT, 2>(h->tm_H0, h->tm_wm2, p, s))) return rc;
    mark_event(marks, s);
  }
  if ((rc = enqueue_encoder(h->enc, h->buf, s, marks, ktime, &kidx))) return rc;
  {  // final projection + DDPM posterior
    PdlBlock no_early_launch(h->buf.stack);  // stack form: must not be parked on the SMs the attention kernel is waiting for
    GemmParams p{};
    p.M = M, p.N = h->nfeat, p.K = d, p.bias = h->b_fin, p.T = T, p.S = S, p.P0 = 5, p.nfeat = h->nfeat;
    p.x_t = x_t, p.x_out = x_out, p.x0_out = x0_out, p.noise = noise, p.t_ptr = advance ? h->t_cur : t_ptr, p.c1 = h->k1, p.c2 = h->k2,
    p.sigma = h->ks, p.seed = seed;
    p.seed_ptr = advance ? h->seed_dev : nullptr;  // chain graph: the seed lives in device memory
    p.ktime = kt();
    if ((rc = launch_gemm<128, EPI_POSTERIOR, 2>(h->tm_Xb_fin, h->tm_wfin, p, s))) return rc;
    mark_event(marks, s);
  }
  return TAMF_OK;
}

static void drop_graph(tamf_denoiser* h) {
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  h->graph_exec = nullptr;
}

}  // namespace tamf

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" int tamf_denoiser_destroy(tamf_denoiser* h) {
  if (!h) return TAMF_OK;
  drop_graph(h);
  h->buf.release();
  h->pool.free_all();
  delete h;
  return TAMF_OK;
}

extern "C" int tamf_denoiser_create(const tamf_cfg* cfg, const tamf_g_weights* w, tamf_denoiser** out) {
  TAMF_REQUIRE(cfg && w && out, TAMF_E_BADARG, "tamf_denoiser_create: null argument");
  int rc = check_device();
  if (rc) return rc;
  const int d = cfg->latent_dim, ff = cfg->ff_size, L = cfg->num_layers, H = cfg->num_heads, nf = cfg->input_dim;
  TAMF_REQUIRE(d == 256 || d == 512, TAMF_E_BADARG, "latent_dim must be 256 or 512 (arch_mdm / arch_mdm_l)");
  TAMF_REQUIRE(nf > 0 && nf < TRAJ_COL, TAMF_E_BADARG, "input_dim must be < 100");
  TAMF_REQUIRE(cfg->obj_input_dim == 9, TAMF_E_BADARG, "obj_input_dim must be 9");
  TAMF_REQUIRE(L > 0 && L <= 64 && w->layers, TAMF_E_BADARG, "bad num_layers");
  TAMF_REQUIRE(cfg->num_steps > 0 && w->pe && w->pe_rows >= cfg->num_steps, TAMF_E_BADARG,
               "pe table must cover num_steps rows");
  TAMF_REQUIRE(w->posterior_mean_coef1 && w->posterior_mean_coef2 && w->posterior_log_variance_clipped, TAMF_E_BADARG,
               "missing schedule tables");
  tamf_denoiser* h = new tamf_denoiser();
  h->cfg = *cfg, h->d = d, h->ff = ff, h->L = L, h->H = H, h->nfeat = nf;
#define TRY(x)                  \
  if ((rc = (x)) != TAMF_OK) {  \
    tamf_denoiser_destroy(h);   \
    return rc;                  \
  }
  TRY(upload_f32(h, &h->shape_w, w->shape_w, (size_t)d * cfg->hand_shape_dim));
  TRY(upload_f32(h, &h->shape_b, w->shape_b, d));
  TRY(upload_f32(h, &h->objemb_w, w->objemb_w, (size_t)d * cfg->obj_embed_dim));
  TRY(upload_f32(h, &h->objemb_b, w->objemb_b, d));
  TRY(upload_f32(h, &h->text_w, w->text_w, (size_t)d * cfg->clip_dim));
  TRY(upload_f32(h, &h->text_b, w->text_b, d));
  h->pe_rows = w->pe_rows;
  TRY(upload_f32(h, &h->pe, w->pe, (size_t)w->pe_rows * d));
  TRY(upload_f32(h, &h->b_m2, w->merge2_b, d));
  TRY(upload_f32(h, &h->b_fin, w->final_b, nf));
  {
    // fold both linear maps feeding input_merge.0 through it (exact in real arithmetic; double accumulation,
    // rounded once):  W1 [Wp x + bp ; mean_o(Wo traj_o + bo)] + b1
    //              = (W1a Wp) x + (W1b Wo) mean_o(traj_o) + (b1 + W1a bp + W1b bo)
    // (the mean over the zero-padded object axis commutes with the linear map, interaction_segment_mdm.py:243-246)
    TAMF_REQUIRE(w->pose_w && w->pose_b && w->objtraj_w && w->objtraj_b && w->merge0_w && w->merge0_b, TAMF_E_BADARG,
                 "null weight pointer");
    std::vector<float> wf((size_t)d * KPAD, 0.f), mb(d);
    std::vector<double> row(KPAD);
    for (int n = 0; n < d; ++n) {
      std::fill(row.begin(), row.end(), 0.0);
      double bacc = (double)w->merge0_b[n];
      const float* w1 = w->merge0_w + (size_t)n * 2 * d;
      for (int j = 0; j < d; ++j) {
        const double a = (double)w1[j], c = (double)w1[d + j];
        const float* pr = w->pose_w + (size_t)j * nf;
        const float* tr = w->objtraj_w + (size_t)j * 9;
        for (int k = 0; k < nf; ++k) row[k] += a * (double)pr[k];
        for (int k = 0; k < 9; ++k) row[TRAJ_COL + k] += c * (double)tr[k];
        bacc += a * (double)w->pose_b[j] + c * (double)w->objtraj_b[j];
      }
      for (int k = 0; k < KPAD; ++k) wf[(size_t)n * KPAD + k] = (float)row[k];
      mb[n] = (float)bacc;
    }
    TRY(upload_bf16(h, &h->wfold, wf.data(), d, KPAD, KPAD));
    TRY(upload_f32(h, &h->merge_bias, mb.data(), d));
  }
  TRY(upload_bf16(h, &h->wm2, w->merge2_w, d, d, d));
  TRY(upload_bf16(h, &h->wfin, w->final_w, nf, d, d));
  TRY(make_tmap_2d_bf16(&h->tm_wfold, h->wfold, KPAD, d, (uint64_t)KPAD * 2, 64, gemm_b_box_rows(256, 2)));
  TRY(make_tmap_2d_bf16(&h->tm_wm2, h->wm2, d, d, (uint64_t)d * 2, 64, gemm_b_box_rows(256, 2)));
  TRY(make_tmap_2d_bf16(&h->tm_wfin, h->wfin, d, nf, (uint64_t)d * 2, 64, gemm_b_box_rows(128, 2)));
  TRY(h->enc.upload(h->pool, w->layers, d, ff, L, H));
  {
    // schedule: fp32 lookups exactly like _extract_into_tensor(...).float() (gaussian_diffusion.py:1275);
    // sigma[t] = 1[t != 0] * exp(0.5 * logvar[t]) evaluated in fp32 (:459)
    const int n = cfg->num_steps;
    std::vector<float> c1(n), c2(n), sg(n);
    for (int t = 0; t < n; ++t) {
      c1[t] = (float)w->posterior_mean_coef1[t];
      c2[t] = (float)w->posterior_mean_coef2[t];
      sg[t] = (t == 0) ? 0.f : expf(0.5f * (float)w->posterior_log_variance_clipped[t]);
    }
    TRY(upload_f32(h, &h->c1, c1.data(), n));
    TRY(upload_f32(h, &h->c2, c2.data(), n));
    TRY(upload_f32(h, &h->sigma, sg.data(), n));
    h->k1 = h->c1, h->k2 = h->c2, h->ks = h->sigma, h->K = n;
    TRY(dev_alloc(h, (void**)&h->alt1, (size_t)n * sizeof(float)));
    TRY(dev_alloc(h, (void**)&h->alt2, (size_t)n * sizeof(float)));
    TRY(dev_alloc(h, (void**)&h->alts, (size_t)n * sizeof(float)));
    TRY(dev_alloc(h, (void**)&h->alt_map, (size_t)n * sizeof(int)));
  }
  {
    // timestep-token table ttab[t] = time_embed(pe[t]) (TimestepEmbedder, interaction_segment_mdm.py:208-215)
    float *t0w, *t0b, *t2w, *t2b, *tmp;
    TRY(upload_f32(h, &t0w, w->time0_w, (size_t)d * d));
    TRY(upload_f32(h, &t0b, w->time0_b, d));
    TRY(upload_f32(h, &t2w, w->time2_w, (size_t)d * d));
    TRY(upload_f32(h, &t2b, w->time2_b, d));
    const int n = cfg->num_steps;
    TRY(dev_alloc(h, (void**)&tmp, (size_t)n * d * sizeof(float)));
    TRY(dev_alloc(h, (void**)&h->ttab, (size_t)n * d * sizeof(float)));
    TRY(dev_alloc(h, (void**)&h->alt_ttab, (size_t)n * d * sizeof(float)));
    h->k_ttab = h->ttab;
    TRY(linear_f32(h->pe, d, t0w, d, t0b, tmp, d, n, d, d, 1, nullptr, 0, nullptr));
    TRY(linear_f32(tmp, d, t2w, d, t2b, h->ttab, d, n, d, d, 0, nullptr, 0, nullptr));
    if (cudaDeviceSynchronize() != cudaSuccess) {
      set_error("tamf_denoiser_create: timestep table kernels failed");
      tamf_denoiser_destroy(h);
      return TAMF_E_CUDA;
    }
  }
  TRY((configure_gemm<256, EPI_BIAS_SILU_BF16, 2>()));
  TRY((configure_gemm<256, EPI_TOKEN_OUT, 2>()));
  TRY((configure_gemm<128, EPI_POSTERIOR, 2>()));
  TRY(configure_encoder_kernels());
#undef TRY
  *out = h;
  return TAMF_OK;
}

namespace tamf {
struct WsLayout {
  size_t off[32];
  size_t total;
};
static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }
static WsLayout ws_layout(const tamf_denoiser* h, int B, int T) {
  const size_t d = h->d, ff = h->ff, S = T + 5, M = (size_t)B * S, Mf = (size_t)B * T;
  const size_t sz[] = {
      M * d * 2,                                 // 0 Xlo
      M * d * 2,                                 // 1 Xb
      M * 3 * d * 2,                             // 2 QKV
      M * d * 2,                                 // 3 ATT
      M * ff * 2,                                // 4 H
      Mf * KPAD * 2,                             // 5 A0
      Mf * d * 2,                                // 6 H0
      encoder_aux_bytes((int)M, (int)d, (int)ff),  // 7 encoder aux: chain-kernel sync words, row statistics, schedules
      (size_t)B * 4 * d * 4,                     // 8 prefix
      256,                                       // 9 seed_dev
      Mf * 9 * 4,                                // 10 trajmean
      (size_t)B * 16 * 4,                        // 11 shapemean
      (size_t)B * h->cfg.obj_embed_dim * 4,      // 12 embmean
      (size_t)B * h->nfeat * T * 4,              // 13 xbuf
      (size_t)B * h->cfg.clip_dim * 4,           // 14 st_text
      (size_t)B * T * h->cfg.hand_shape_dim * 4, // 15 st_shape
      (size_t)B * MAX_NOBJ * T * 9 * 4,          // 16 st_traj
      (size_t)B * MAX_NOBJ * h->cfg.obj_embed_dim * 4,  // 17 st_emb
      (size_t)B * 4,                             // 18 st_side
      (size_t)B * 4 * 2,                         // 19 t_dev (chain counter) | t_cur (timestep of the running step)
  };
  WsLayout L{};
  size_t o = 0;
  for (int i = 0; i < 20; ++i) {
    L.off[i] = o;
    o += al(sz[i]);
  }
  L.total = o;
  return L;
}
}  // namespace tamf

extern "C" size_t tamf_denoiser_workspace_bytes(const tamf_denoiser* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return ws_layout(h, B, T).total;
}

extern "C" int tamf_denoiser_bind(tamf_denoiser* h, int B, int T, void* ws, size_t ws_bytes) {
  TAMF_REQUIRE(h && ws, TAMF_E_BADARG, "tamf_denoiser_bind: null argument");
  TAMF_REQUIRE(B > 0 && T > 0, TAMF_E_BADARG, "tamf_denoiser_bind: B and T must be positive");
  TAMF_REQUIRE(T + 5 <= ATC_KP, TAMF_E_BADARG, "tamf_denoiser_bind: T + 5 tokens must be <= 176");
  TAMF_REQUIRE(T + 5 <= h->pe_rows, TAMF_E_BADARG, "tamf_denoiser_bind: pe table too short");
  TAMF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, TAMF_E_ALIGN, "workspace must be 256-byte aligned");
  WsLayout L = ws_layout(h, B, T);
  TAMF_REQUIRE(ws_bytes >= L.total, TAMF_E_BADARG, "tamf_denoiser_bind: workspace too small");
  drop_graph(h);
  uint8_t* p = static_cast<uint8_t*>(ws);
  h->B = B, h->T = T, h->S = T + 5, h->M = B * (T + 5), h->Mf = B * T;
  h->buf.B = B, h->buf.S = T + 5, h->buf.M = h->M;
  h->buf.Xlo = (__nv_bfloat16*)(p + L.off[0]);
  h->buf.Xb = (__nv_bfloat16*)(p + L.off[1]);
  h->buf.QKV = (__nv_bfloat16*)(p + L.off[2]);
  h->buf.ATT = (__nv_bfloat16*)(p + L.off[3]);
  h->buf.Hb = (__nv_bfloat16*)(p + L.off[4]);
  h->A0 = (__nv_bfloat16*)(p + L.off[5]);
  h->H0 = (__nv_bfloat16*)(p + L.off[6]);
  h->buf.aux = p + L.off[7];
  h->prefix = (float*)(p + L.off[8]);
  h->seed_dev = (unsigned long long*)(p + L.off[9]);
  h->trajmean = (float*)(p + L.off[10]);
  h->shapemean = (float*)(p + L.off[11]);
  h->embmean = (float*)(p + L.off[12]);
  h->xbuf = (float*)(p + L.off[13]);
  h->st_text = (float*)(p + L.off[14]);
  h->st_shape = (float*)(p + L.off[15]);
  h->st_traj = (float*)(p + L.off[16]);
  h->st_emb = (float*)(p + L.off[17]);
  h->st_side = (int*)(p + L.off[18]);
  h->t_dev = (int*)(p + L.off[19]);
  h->t_cur = h->t_dev + B;
  const int d = h->d, ff = h->ff;
  int rc;
  if ((rc = h->buf.make_maps(d, ff, h->L, h->H))) return rc;
  h->tm_Xb_fin = h->buf.tm_Xb;
  if ((rc = make_tmap_2d_bf16(&h->tm_A0, h->A0, KPAD, h->Mf, (uint64_t)KPAD * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&h->tm_H0, h->H0, d, h->Mf, (uint64_t)d * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&h->tm_H0_st, h->H0, d, h->Mf, (uint64_t)d * 2, 64, 32))) return rc;
  if ((rc = make_token_out_maps(&h->tm_tok_hi, &h->tm_tok_lo, h->buf.Xb, h->buf.Xlo, h->B, h->T, h->S, 5, h->d))) return rc;
  h->bound = true;
  h->cond_set = false;
  return TAMF_OK;
}

extern "C" int tamf_denoiser_set_cond(tamf_denoiser* h, const float* text_feat, const int32_t* hand_side,
                                      const float* shape, const float* obj_traj, const float* obj_emb, int nobj_max,
                                      void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  TAMF_REQUIRE(h && h->bound, TAMF_E_STATE, "tamf_denoiser_set_cond: bind a workspace first");
  TAMF_REQUIRE(text_feat && hand_side && shape && obj_traj && obj_emb, TAMF_E_BADARG, "set_cond: null pointer");
  TAMF_REQUIRE(nobj_max >= 1, TAMF_E_BADARG, "set_cond: nobj_max must be >= 1");
  const int d = h->d, B = h->B, T = h->T, Mf = h->Mf;
  const tamf_cfg& c = h->cfg;
  int rc;
  // means over frames / (padded) objects: interaction_segment_mdm.py:300, :260, :245
  {
    TAMF_REQUIRE(c.obj_input_dim == 9, TAMF_E_BADARG, "obj_input_dim must be 9");
    if ((rc = mean_axis(shape, h->shapemean, B, T, c.hand_shape_dim, s))) return rc;
    if ((rc = mean_axis(obj_emb, h->embmean, B, nobj_max, c.obj_embed_dim, s))) return rc;
    if ((rc = traj_mean(obj_traj, h->trajmean, B, nobj_max, T, s))) return rc;
  }
  // prefix tokens 1..4 -> prefix[b, 0..3, :]
  if ((rc = linear_f32(text_feat, c.clip_dim, h->text_w, c.clip_dim, h->text_b, h->prefix + 0 * d, 4 * d, B, d,
                       c.clip_dim, 0, nullptr, 0, s)))
    return rc;
  if ((rc = linear_f32(h->shapemean, c.hand_shape_dim, h->shape_w, c.hand_shape_dim, h->shape_b, h->prefix + 2 * d,
                       4 * d, B, d, c.hand_shape_dim, 0, nullptr, 0, s)))
    return rc;
  if ((rc = linear_f32(h->embmean, c.obj_embed_dim, h->objemb_w, c.obj_embed_dim, h->objemb_b, h->prefix + 3 * d, 4 * d,
                       B, d, c.obj_embed_dim, 0, nullptr, 0, s)))
    return rc;
  {
    const size_t n = (size_t)B * 4 * d;
    prefix_finish_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->prefix, hand_side, h->pe, B, d);
    TAMF_LAUNCH_CHECK();
  }
  // trajectory columns of the embed-a operand (the object half of input_merge.0 is folded into Wfold)
  {
    const size_t n = (size_t)Mf * (KPAD - TRAJ_COL);
    a0_cond_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->trajmean, h->A0, Mf);
    TAMF_LAUNCH_CHECK();
  }
  h->cond_set = true;
  return TAMF_OK;
}

extern "C" int tamf_denoiser_forward(tamf_denoiser* h, const float* x_t, const int32_t* t, float* x0_out, void* stream) {
  TAMF_REQUIRE(h && h->bound && h->cond_set, TAMF_E_STATE, "tamf_denoiser_forward: bind + set_cond first");
  TAMF_REQUIRE(x_t && t && x0_out, TAMF_E_BADARG, "tamf_denoiser_forward: null pointer");
  // (t is only written in chain mode, `advance`)
  return enqueue_step(h, x_t, const_cast<int32_t*>(t), nullptr, x0_out, nullptr, 0, (cudaStream_t)stream, nullptr,
                      /*model_level=*/true);
}

extern "C" int tamf_denoiser_set_sampler(tamf_denoiser* h, int K, const float* c1, const float* c2, const float* sigma,
                                         const int32_t* timestep_map) {
  TAMF_REQUIRE(h, TAMF_E_BADARG, "tamf_denoiser_set_sampler: null handle");
  const int n = h->cfg.num_steps;
  drop_graph(h);  // the captured step has the table pointers baked in
  if (K == 0) {
    h->k1 = h->c1, h->k2 = h->c2, h->ks = h->sigma, h->k_ttab = h->ttab, h->K = n;
    return TAMF_OK;
  }
  TAMF_REQUIRE(K >= 1 && K <= n, TAMF_E_BADARG, "tamf_denoiser_set_sampler: need 1 <= K <= num_steps");
  TAMF_REQUIRE(c1 && c2 && sigma && timestep_map, TAMF_E_BADARG, "tamf_denoiser_set_sampler: null table");
  for (int i = 0; i < K; ++i)
    TAMF_REQUIRE(timestep_map[i] >= 0 && timestep_map[i] < n, TAMF_E_BADARG,
                 "tamf_denoiser_set_sampler: timestep_map entry outside [0, num_steps)");
  TAMF_CUDA_CHECK(cudaDeviceSynchronize());  // nothing in flight may still read the tables being replaced
  TAMF_CUDA_CHECK(cudaMemcpy(h->alt1, c1, (size_t)K * 4, cudaMemcpyHostToDevice));
  TAMF_CUDA_CHECK(cudaMemcpy(h->alt2, c2, (size_t)K * 4, cudaMemcpyHostToDevice));
  TAMF_CUDA_CHECK(cudaMemcpy(h->alts, sigma, (size_t)K * 4, cudaMemcpyHostToDevice));
  TAMF_CUDA_CHECK(cudaMemcpy(h->alt_map, timestep_map, (size_t)K * 4, cudaMemcpyHostToDevice));
  gather_rows_kernel<<<K, 128>>>(h->ttab, h->alt_map, h->alt_ttab, h->d);
  TAMF_LAUNCH_CHECK();
  TAMF_CUDA_CHECK(cudaDeviceSynchronize());
  h->k1 = h->alt1, h->k2 = h->alt2, h->ks = h->alts, h->k_ttab = h->alt_ttab, h->K = K;
  return TAMF_OK;
}

extern "C" int tamf_denoiser_sampler_steps(const tamf_denoiser* h) { return h ? h->K : 0; }

extern "C" int tamf_p_sample_step(tamf_denoiser* h, float* x_io, int t, const float* noise, uint64_t seed,
                                  float* x0_out, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  TAMF_REQUIRE(h && h->bound && h->cond_set, TAMF_E_STATE, "tamf_p_sample_step: bind + set_cond first");
  TAMF_REQUIRE(x_io, TAMF_E_BADARG, "tamf_p_sample_step: null pointer");
  TAMF_REQUIRE(t >= 0 && t < h->K, TAMF_E_BADARG, "tamf_p_sample_step: t out of range");
  {
    int rc0 = fill_int(h->t_dev, h->B, t, s);
    if (rc0) return rc0;
  }
  return enqueue_step(h, x_io, h->t_dev, x_io, x0_out, noise, seed, s);
}

extern "C" int tamf_denoiser_profile_step(tamf_denoiser* h, float* x_io, int t, uint64_t seed, float* ms_out, int cap,
                                          int* n_out, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  TAMF_REQUIRE(h && h->bound && h->cond_set, TAMF_E_STATE, "tamf_denoiser_profile_step: bind + set_cond first");
  TAMF_REQUIRE(x_io && ms_out && n_out, TAMF_E_BADARG, "tamf_denoiser_profile_step: null pointer");
  TAMF_REQUIRE(t >= 0 && t < h->K, TAMF_E_BADARG, "tamf_denoiser_profile_step: t out of range");
  {
    int rc0 = fill_int(h->t_dev, h->B, t, s);
    if (rc0) return rc0;
  }
  std::vector<cudaEvent_t> marks;
  int rc = enqueue_step(h, x_io, h->t_dev, x_io, nullptr, nullptr, seed, s, &marks);
  cudaError_t e = cudaStreamSynchronize(s);
  const int n = (int)marks.size() - 1;
  if (rc == TAMF_OK && e == cudaSuccess) {
    for (int i = 0; i < n && i < cap; ++i) cudaEventElapsedTime(&ms_out[i], marks[i], marks[i + 1]);
    *n_out = n;
  }
  for (cudaEvent_t ev : marks) cudaEventDestroy(ev);
  if (rc) return rc;
  TAMF_CUDA_CHECK(e);
  return TAMF_OK;
}

namespace tamf {
__global__ void set_u64_kernel(unsigned long long* p, unsigned long long v) { *p = v; }
__global__ void ktime_init_kernel(long long* kt, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    kt[4 * i + 0] = 0x7fffffffffffffffLL, kt[4 * i + 1] = 0x7fffffffffffffffLL;
    kt[4 * i + 2] = 0, kt[4 * i + 3] = 0;
  }
}
}  // namespace tamf

// In-graph per-kernel breakdown: the step is captured in a CUDA graph exactly like tamf_p_sample_chain captures it (same
// kernels, same programmatic dependent launches), but every kernel also records, on the globaltimer every SM shares, the
// earliest CTA entry, the earliest end of a dependency wait and the latest CTA exit.  n_steps replays run back to back;
// the figures of the last replays are averaged.  Per kernel k (launch order): entry_us[k], ready_us[k], exit_us[k] relative
// to the first kernel's entry of the same step; step_us = mean step period.
extern "C" int tamf_denoiser_profile_graph(tamf_denoiser* h, float* x_io, int t_start, int n_steps, uint64_t seed,
                                           double* entry_us, double* ready_us, double* exit_us, int cap, int* n_out,
                                           double* step_us, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  TAMF_REQUIRE(h && h->bound && h->cond_set, TAMF_E_STATE, "tamf_denoiser_profile_graph: bind + set_cond first");
  TAMF_REQUIRE(x_io && entry_us && ready_us && exit_us && n_out && step_us, TAMF_E_BADARG,
               "tamf_denoiser_profile_graph: null pointer");
  TAMF_REQUIRE(n_steps >= 4 && t_start < h->K && t_start - n_steps + 1 >= 0, TAMF_E_BADARG,
               "tamf_denoiser_profile_graph: need n_steps >= 4 steps inside the schedule");
  const int MAXK = 128, AVG = 3;  // the last AVG steps are averaged
  long long* kt = nullptr;
  TAMF_CUDA_CHECK(cudaMalloc(&kt, (size_t)AVG * MAXK * 4 * sizeof(long long)));
  cudaStream_t cap_s = nullptr;
  TAMF_CUDA_CHECK(cudaStreamCreateWithFlags(&cap_s, cudaStreamNonBlocking));
  cudaGraphExec_t exec[AVG] = {nullptr, nullptr, nullptr};
  int nk = 0, rc = TAMF_OK;
  const uint64_t before = g_launches.load();
  for (int a = 0; a < AVG && rc == TAMF_OK; ++a) {  // one graph per averaged step: each writes its own timing slots
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(cap_s, cudaStreamCaptureModeThreadLocal);
    if (e == cudaSuccess) {
      long long* slot = kt + (size_t)a * MAXK * 4;
      ktime_init_kernel<<<1, MAXK, 0, cap_s>>>(slot, MAXK);
      const uint64_t b0 = g_launches.load();
      rc = enqueue_step(h, x_io, h->t_dev, x_io, nullptr, nullptr, seed, cap_s, nullptr, false, /*advance=*/true, slot);
      nk = (int)(g_launches.load() - b0);
      e = cudaStreamEndCapture(cap_s, &g);
    }
    if (rc == TAMF_OK && e == cudaSuccess) e = cudaGraphInstantiate(&exec[a], g, 0);
    if (g) cudaGraphDestroy(g);
    if (rc == TAMF_OK && e != cudaSuccess) {
      set_error(std::string("tamf_denoiser_profile_graph: ") + cudaGetErrorString(e));
      rc = TAMF_E_CUDA;
    }
  }
  g_launches.store(before);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  float ms = 0.f;
  if (rc == TAMF_OK && nk <= MAXK && nk <= cap) {
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    rc = fill_int(h->t_dev, h->B, t_start, s);
    set_u64_kernel<<<1, 1, 0, s>>>(h->seed_dev, seed);
    // warm-up replays with the product graph shape (slot 0), then the timed tail: ... AVG graphs last
    for (int i = 0; i < n_steps && rc == TAMF_OK; ++i) {
      const int a = (i >= n_steps - AVG) ? i - (n_steps - AVG) : 0;
      if (i == n_steps - AVG) cudaEventRecord(e0, s);
      if (cudaGraphLaunch(exec[a], s) != cudaSuccess) rc = TAMF_E_CUDA;
      count_launch(nk + 1);
    }
    cudaEventRecord(e1, s);
    if (cudaStreamSynchronize(s) != cudaSuccess) rc = TAMF_E_CUDA;
    if (rc == TAMF_OK) cudaEventElapsedTime(&ms, e0, e1);
  } else if (rc == TAMF_OK) {
    set_error("tamf_denoiser_profile_graph: output capacity too small");
    rc = TAMF_E_BADARG;
  }
  if (rc == TAMF_OK) {
    std::vector<long long> hst((size_t)AVG * MAXK * 4);
    if (cudaMemcpy(hst.data(), kt, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost) != cudaSuccess) rc = TAMF_E_CUDA;
    for (int k = 0; k < nk && rc == TAMF_OK; ++k) {
      double en = 0, rd = 0, ex = 0;
      for (int a = 0; a < AVG; ++a) {
        const long long* z = hst.data() + (size_t)a * MAXK * 4;
        const long long t0 = z[0];  // entry of the step's first kernel
        en += (double)(z[4 * k + 0] - t0), rd += (double)(z[4 * k + 1] - t0), ex += (double)(z[4 * k + 2] - t0);
      }
      entry_us[k] = en / AVG * 1e-3, ready_us[k] = rd / AVG * 1e-3, exit_us[k] = ex / AVG * 1e-3;
    }
    *n_out = nk;
    *step_us = (double)ms * 1e3 / AVG;
  }
  for (int a = 0; a < AVG; ++a)
    if (exec[a]) cudaGraphExecDestroy(exec[a]);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaStreamDestroy(cap_s);
  cudaFree(kt);
  return rc;
}

extern "C" int tamf_p_sample_chain(tamf_denoiser* h, float* x_io, int t_start, int t_end, uint64_t seed,
                                   void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  TAMF_REQUIRE(h && h->bound && h->cond_set, TAMF_E_STATE, "tamf_p_sample_chain: bind + set_cond first");
  TAMF_REQUIRE(x_io, TAMF_E_BADARG, "tamf_p_sample_chain: null pointer");
  TAMF_REQUIRE(t_start < h->K && t_end >= 0 && t_end <= t_start, TAMF_E_BADARG,
               "tamf_p_sample_chain: need sampler steps > t_start >= t_end >= 0");
  // the captured step depends on the x buffer and the stream only: the timestep counter and the seed are read from device
  // memory, so consecutive chains on the same buffer (sample_dataset, bench.py --sequences) replay ONE graph
  if (!h->graph_exec || h->graph_x != x_io || h->graph_stream != s) {
    drop_graph(h);
    cudaStream_t cap = s;
    cudaStream_t own = nullptr;
    if (cap == nullptr) {  // the legacy default stream cannot be captured
      TAMF_CUDA_CHECK(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
      cap = own;
    }
    cudaGraph_t g = nullptr;
    TAMF_CUDA_CHECK(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    const uint64_t before = g_launches.load();
    int rc = enqueue_step(h, x_io, h->t_dev, x_io, nullptr, nullptr, seed, cap, nullptr, false, /*advance=*/true);
    h->graph_kernels = (int)(g_launches.load() - before);
    g_launches.store(before);  // capture records, it does not launch; replays are counted below
    cudaError_t e = cudaStreamEndCapture(cap, &g);
    if (own) cudaStreamDestroy(own);
    if (rc != TAMF_OK) {
      if (g) cudaGraphDestroy(g);
      return rc;
    }
    TAMF_CUDA_CHECK(e);
    e = cudaGraphInstantiate(&h->graph_exec, g, 0);
    cudaGraphDestroy(g);
    TAMF_CUDA_CHECK(e);
    h->graph_x = x_io, h->graph_stream = s;
  }
  {
    int rc0 = fill_int(h->t_dev, h->B, t_start, s);
    if (rc0) return rc0;
    set_u64_kernel<<<1, 1, 0, s>>>(h->seed_dev, seed);
    TAMF_LAUNCH_CHECK();
  }
  const int per_step = h->graph_kernels;  // kernels of one captured step
  for (int t = t_start; t >= t_end; --t) {
    TAMF_CUDA_CHECK(cudaGraphLaunch(h->graph_exec, s));
    count_launch(per_step);
  }
  return TAMF_OK;
}

extern "C" int tamf_p_sample_loop_host(tamf_denoiser* h, const float* text_feat, const int32_t* hand_side,
                                       const float* shape, const float* obj_traj, const float* obj_emb, int nobj_max,
                                       const float* x_T, uint64_t seed, float* sample_out, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  TAMF_REQUIRE(h && h->bound, TAMF_E_STATE, "tamf_p_sample_loop_host: bind a workspace first");
  TAMF_REQUIRE(text_feat && hand_side && shape && obj_traj && obj_emb && sample_out, TAMF_E_BADARG,
               "tamf_p_sample_loop_host: null pointer");
  TAMF_REQUIRE(nobj_max >= 1 && nobj_max <= MAX_NOBJ, TAMF_E_BADARG, "tamf_p_sample_loop_host: 1 <= nobj_max <= 8");
  const int B = h->B, T = h->T;
  const tamf_cfg& c = h->cfg;
  const size_t nx = (size_t)B * h->nfeat * T;
  TAMF_CUDA_CHECK(cudaMemcpyAsync(h->st_text, text_feat, (size_t)B * c.clip_dim * 4, cudaMemcpyHostToDevice, s));
  TAMF_CUDA_CHECK(cudaMemcpyAsync(h->st_side, hand_side, (size_t)B * 4, cudaMemcpyHostToDevice, s));
  TAMF_CUDA_CHECK(cudaMemcpyAsync(h->st_shape, shape, (size_t)B * T * c.hand_shape_dim * 4, cudaMemcpyHostToDevice, s));
  TAMF_CUDA_CHECK(cudaMemcpyAsync(h->st_traj, obj_traj, (size_t)B * nobj_max * T * 9 * 4, cudaMemcpyHostToDevice, s));
  TAMF_CUDA_CHECK(
      cudaMemcpyAsync(h->st_emb, obj_emb, (size_t)B * nobj_max * c.obj_embed_dim * 4, cudaMemcpyHostToDevice, s));
  int rc = tamf_denoiser_set_cond(h, h->st_text, h->st_side, h->st_shape, h->st_traj, h->st_emb, nobj_max, s);
  if (rc) return rc;
  if (x_T) {
    TAMF_CUDA_CHECK(cudaMemcpyAsync(h->xbuf, x_T, nx * 4, cudaMemcpyHostToDevice, s));
  } else {
    if ((rc = philox_fill(h->xbuf, nx, seed, (uint32_t)h->K, s))) return rc;  // th.randn(*shape), :604
  }
  if ((rc = tamf_p_sample_chain(h, h->xbuf, h->K - 1, 0, seed, s))) return rc;
  TAMF_CUDA_CHECK(cudaMemcpyAsync(sample_out, h->xbuf, nx * 4, cudaMemcpyDeviceToHost, s));
  TAMF_CUDA_CHECK(cudaStreamSynchronize(s));
  return TAMF_OK;
}
