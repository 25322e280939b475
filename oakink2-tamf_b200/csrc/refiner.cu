// refiner.cu -- MF-MDM R (SegmentRefineModel) transformer pass: residual pose refinement from the sampled pose,
// the object trajectory and the per-vertex hand->object distances.
//
// Replaces the tensor part of SegmentRefineModel.forward (src/oakink2_tamf/model/segment_refine_model.py:170-217):
//   prefix tokens (hand side, shape, object embedding)   :177-186
//   input_process / obj_input_process / h2o_dist_input_process + input_merge (3d -> d, SiLU, d -> d)  :189-208
//   + positional encoding, 8-layer encoder, [3:], output_process, x_in + output, nan_to_num            :211-217
// FK and the hand->object distances (:193-201, :220-232) are tamf_mano_fk(_select) / tamf_h2o_dist; the Python
// drop-in (tamf_b200/refine.py) sequences the calls.
//
// Frame-token embedding, folded like G's (exact in real arithmetic; double accumulation, rounded once):
//   h = W1a (Wp x + bp) + W1b mean_o(Wo traj_o + bo) + W1c (Wd dist + bd) + b1
//     = [W1a Wp | W1b Wo | W1c Wd] . [x ; mean_o(traj_o) ; dist]  +  (b1 + W1a bp + W1b bo + W1c bd)
//   A0 [B*T, 960] bf16 = [x (99) | 0 | traj mean (9) | 0.. (-> 128) | dist (778 -> 832)]
//                                                                 one tcgen05 GEMM with K = 960, SiLU epilogue
//   tok[3+tau] = nan_to_num(silu(h) . W2^T + b2) + pe[3+tau]     second GEMM, token epilogue
#include "encoder.cuh"
#include "gemm.cuh"

namespace tamf {

constexpr int R_NV = 778;         // hand vertices = h2o_dist features
constexpr int R_KX = 128;         // 99 pose features, zero padded
constexpr int R_KD = 832;         // 778 distances, zero padded to a multiple of 64
constexpr int R_K = R_KX + R_KD;  // 960
constexpr int R_PREFIX = 3;
constexpr int R_TRAJ_COL = 100;   // first column of mean_obj(obj_traj) inside the pose block

// A0[r, :] = [x_in[r, 0:99], 0, trajmean[r, 0:9], 0.., dist[r, 0:778], 0..] as bf16; one thread per pair of columns.
__global__ void r_prep_kernel(const float* __restrict__ x_in, const float* __restrict__ trajmean,
                              const float* __restrict__ dist, __nv_bfloat16* __restrict__ A0, int rows, int nfeat) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * (R_K / 2)) return;
  const int r = (int)(i / (R_K / 2)), c = (int)(i % (R_K / 2)) * 2;
  float v[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int cc = c + e;
    if (cc < R_KX)
      v[e] = cc < nfeat ? x_in[(size_t)r * nfeat + cc]
                        : ((cc >= R_TRAJ_COL && cc < R_TRAJ_COL + 9) ? trajmean[(size_t)r * 9 + (cc - R_TRAJ_COL)] : 0.f);
    else
      v[e] = (cc - R_KX) < R_NV ? dist[(size_t)r * R_NV + (cc - R_KX)] : 0.f;
  }
  *reinterpret_cast<uint32_t*>(A0 + (size_t)r * R_K + c) = pack_bf16x2(v[0], v[1]);
}

// prefix[b, s, :] (s = 0 hand side: rh -> 0, lh -> e0; s = 1 shape token; s = 2 object token, both precomputed)
// -> nan_to_num(.) + pe[s] written to token rows b*S + s of the residual planes Xb / Xlo.
__global__ void r_prefix_kernel(const float* __restrict__ prefix, const int* __restrict__ hand_side,
                                const float* __restrict__ pe, __nv_bfloat16* __restrict__ Xlo,
                                __nv_bfloat16* __restrict__ Xb, int B, int S, int d) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * R_PREFIX * d) return;
  const int c = (int)(i % d), s = (int)((i / d) % R_PREFIX), b = (int)(i / (R_PREFIX * (size_t)d));
  float v = prefix[i];
  if (s == 0) v = (hand_side[b] == 1 && c == 0) ? 1.f : 0.f;
  v = nan_to_num(v) + pe[(size_t)s * d + c];
  store_hilo(Xb, Xlo, ((size_t)b * S + s) * d + c, v);
}

}  // namespace tamf

using namespace tamf;

struct tamf_refiner {
  tamf_cfg cfg{};
  int d = 0, ff = 0, nfeat = 0;
  DevPool pool;
  EncoderStack enc;
  EncoderBuffers buf;
  float *shape_w, *shape_b, *objemb_w, *objemb_b, *merge_bias, *pe, *b_m2, *b_fin;
  int pe_rows = 0;
  __nv_bfloat16 *wfold /*[d,960]*/, *wm2, *wfin;
  CUtensorMap tm_wfold, tm_wm2, tm_wfin, tm_A0, tm_H0, tm_H0_st;
  CUtensorMap tm_tok_hi, tm_tok_lo;  // EPI_TOKEN_OUT stores (make_token_out_maps)
  int B = 0, T = 0, S = 0, M = 0, Mf = 0;
  bool bound = false;
  float *prefix, *trajmean, *shapemean, *embmean;
  __nv_bfloat16 *A0, *H0;
};

extern "C" int tamf_refiner_destroy(tamf_refiner* h) {
  if (!h) return TAMF_OK;
  h->buf.release();
  h->pool.free_all();
  delete h;
  return TAMF_OK;
}

extern "C" int tamf_refiner_create(const tamf_cfg* cfg, const tamf_r_weights* w, tamf_refiner** out) {
  TAMF_REQUIRE(cfg && w && out, TAMF_E_BADARG, "tamf_refiner_create: null argument");
  int rc = check_device();
  if (rc) return rc;
  const int d = cfg->latent_dim, ff = cfg->ff_size, nf = cfg->input_dim;
  TAMF_REQUIRE(d == 256 || d == 512, TAMF_E_BADARG, "latent_dim must be 256 or 512");
  TAMF_REQUIRE(nf > 0 && nf < R_TRAJ_COL, TAMF_E_BADARG, "input_dim must be < 100");
  TAMF_REQUIRE(cfg->obj_input_dim == 9, TAMF_E_BADARG, "obj_input_dim must be 9");
  TAMF_REQUIRE(w->pe && w->pe_rows >= R_PREFIX + 1, TAMF_E_BADARG, "pe table missing");
  TAMF_REQUIRE(w->pose_w && w->pose_b && w->dist_w && w->dist_b && w->merge0_w && w->merge0_b && w->objtraj_w &&
                   w->objtraj_b,
               TAMF_E_BADARG, "null weight pointer");
  tamf_refiner* h = new tamf_refiner();
  h->cfg = *cfg, h->d = d, h->ff = ff, h->nfeat = nf;
#define TRY(x)                 \
  if ((rc = (x)) != TAMF_OK) { \
    tamf_refiner_destroy(h);   \
    return rc;                 \
  }
  DevPool& P = h->pool;
  TRY(P.upload_f32(&h->shape_w, w->shape_w, (size_t)d * cfg->hand_shape_dim));
  TRY(P.upload_f32(&h->shape_b, w->shape_b, d));
  TRY(P.upload_f32(&h->objemb_w, w->objemb_w, (size_t)d * cfg->obj_embed_dim));
  TRY(P.upload_f32(&h->objemb_b, w->objemb_b, d));
  h->pe_rows = w->pe_rows;
  TRY(P.upload_f32(&h->pe, w->pe, (size_t)w->pe_rows * d));
  TRY(P.upload_f32(&h->b_m2, w->merge2_b, d));
  TRY(P.upload_f32(&h->b_fin, w->final_b, nf));
  {
    // fold: Wfold[n, 0:99] = sum_j W1[n, j] Wp[j, :], Wfold[n, 100:109] = sum_j W1[n, d + j] Wo[j, :],
    //       Wfold[n, 128:906] = sum_j W1[n, 2d + j] Wd[j, :]
    //       bias[n] = b1[n] + sum_j (W1[n, j] bp[j] + W1[n, d + j] bo[j] + W1[n, 2d + j] bd[j])
    std::vector<float> wf((size_t)d * R_K, 0.f), mb(d);
    std::vector<double> row(R_K);
    for (int n = 0; n < d; ++n) {
      std::fill(row.begin(), row.end(), 0.0);
      double bacc = (double)w->merge0_b[n];
      const float* w1 = w->merge0_w + (size_t)n * 3 * d;
      for (int j = 0; j < d; ++j) {
        const double a = (double)w1[j], o = (double)w1[d + j], c = (double)w1[2 * d + j];
        const float* pr = w->pose_w + (size_t)j * nf;
        const float* tr = w->objtraj_w + (size_t)j * 9;
        const float* dr = w->dist_w + (size_t)j * R_NV;
        for (int k = 0; k < nf; ++k) row[k] += a * (double)pr[k];
        for (int k = 0; k < 9; ++k) row[R_TRAJ_COL + k] += o * (double)tr[k];
        for (int k = 0; k < R_NV; ++k) row[R_KX + k] += c * (double)dr[k];
        bacc += a * (double)w->pose_b[j] + o * (double)w->objtraj_b[j] + c * (double)w->dist_b[j];
      }
      for (int k = 0; k < R_K; ++k) wf[(size_t)n * R_K + k] = (float)row[k];
      mb[n] = (float)bacc;
    }
    TRY(P.upload_bf16(&h->wfold, wf.data(), d, R_K, R_K));
    TRY(P.upload_f32(&h->merge_bias, mb.data(), d));
  }
  TRY(P.upload_bf16(&h->wm2, w->merge2_w, d, d, d));
  TRY(P.upload_bf16(&h->wfin, w->final_w, nf, d, d));
  TRY(make_tmap_2d_bf16(&h->tm_wfold, h->wfold, R_K, d, (uint64_t)R_K * 2, 64, gemm_b_box_rows(256, 2)));
  TRY(make_tmap_2d_bf16(&h->tm_wm2, h->wm2, d, d, (uint64_t)d * 2, 64, gemm_b_box_rows(256, 2)));
  TRY(make_tmap_2d_bf16(&h->tm_wfin, h->wfin, d, nf, (uint64_t)d * 2, 64, gemm_b_box_rows(128, 2)));
  TRY(h->enc.upload(P, w->layers, d, ff, cfg->num_layers, cfg->num_heads));
  TRY((configure_gemm<256, EPI_BIAS_SILU_BF16, 2>()));
  TRY((configure_gemm<256, EPI_TOKEN_OUT, 2>()));
  TRY((configure_gemm<128, EPI_RESIDUAL_OUT, 2>()));
  TRY(configure_encoder_kernels());
#undef TRY
  *out = h;
  return TAMF_OK;
}

namespace tamf {
struct RWs {
  size_t off[16];
  size_t total;
};
static RWs r_layout(const tamf_refiner* h, int B, int T) {
  const size_t d = h->d, ff = h->ff, S = T + R_PREFIX, M = (size_t)B * S, Mf = (size_t)B * T;
  const size_t sz[] = {
      M * d * 2,                             // 0 Xlo
      M * d * 2,                             // 1 Xb
      M * 3 * d * 2,                         // 2 QKV
      M * d * 2,                             // 3 ATT
      M * ff * 2,                            // 4 H
      Mf * R_K * 2,                          // 5 A0
      Mf * d * 2,                            // 6 H0
      encoder_aux_bytes((int)M, (int)d, (int)ff),  // 7 encoder aux: chain-kernel sync words, row statistics, schedules
      (size_t)B * R_PREFIX * d * 4,          // 8 prefix
      256,                                   // 9 (unused)
      Mf * 9 * 4,                            // 10 trajmean
      (size_t)B * 16 * 4,                    // 11 shapemean
      (size_t)B * h->cfg.obj_embed_dim * 4,  // 12 embmean
  };
  RWs L{};
  size_t o = 0;
  for (int i = 0; i < 13; ++i) {
    L.off[i] = o;
    o += (sz[i] + 255) & ~(size_t)255;
  }
  L.total = o;
  return L;
}
}  // namespace tamf

extern "C" size_t tamf_refiner_workspace_bytes(const tamf_refiner* h, int B, int T) {
  if (!h || B <= 0 || T <= 0) return 0;
  return r_layout(h, B, T).total;
}

extern "C" int tamf_refiner_bind(tamf_refiner* h, int B, int T, void* ws, size_t ws_bytes) {
  TAMF_REQUIRE(h && ws, TAMF_E_BADARG, "tamf_refiner_bind: null argument");
  TAMF_REQUIRE(B > 0 && T > 0, TAMF_E_BADARG, "tamf_refiner_bind: B and T must be positive");
  TAMF_REQUIRE(T + R_PREFIX <= 176, TAMF_E_BADARG, "tamf_refiner_bind: T + 3 tokens must be <= 176");
  TAMF_REQUIRE(T + R_PREFIX <= h->pe_rows, TAMF_E_BADARG, "tamf_refiner_bind: pe table too short");
  TAMF_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, TAMF_E_ALIGN, "workspace must be 256-byte aligned");
  RWs L = r_layout(h, B, T);
  TAMF_REQUIRE(ws_bytes >= L.total, TAMF_E_BADARG, "tamf_refiner_bind: workspace too small");
  uint8_t* p = static_cast<uint8_t*>(ws);
  h->B = B, h->T = T, h->S = T + R_PREFIX, h->M = B * h->S, h->Mf = B * T;
  h->buf.B = B, h->buf.S = h->S, h->buf.M = h->M;
  h->buf.Xlo = (__nv_bfloat16*)(p + L.off[0]);
  h->buf.Xb = (__nv_bfloat16*)(p + L.off[1]);
  h->buf.QKV = (__nv_bfloat16*)(p + L.off[2]);
  h->buf.ATT = (__nv_bfloat16*)(p + L.off[3]);
  h->buf.Hb = (__nv_bfloat16*)(p + L.off[4]);
  h->A0 = (__nv_bfloat16*)(p + L.off[5]);
  h->H0 = (__nv_bfloat16*)(p + L.off[6]);
  h->buf.aux = p + L.off[7];
  h->prefix = (float*)(p + L.off[8]);
  h->trajmean = (float*)(p + L.off[10]);
  h->shapemean = (float*)(p + L.off[11]);
  h->embmean = (float*)(p + L.off[12]);
  int rc;
  if ((rc = h->buf.make_maps(h->d, h->ff, h->enc.L, h->enc.H))) return rc;
  if ((rc = make_tmap_2d_bf16(&h->tm_A0, h->A0, R_K, h->Mf, (uint64_t)R_K * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&h->tm_H0, h->H0, h->d, h->Mf, (uint64_t)h->d * 2, 64, 128))) return rc;
  if ((rc = make_tmap_2d_bf16(&h->tm_H0_st, h->H0, h->d, h->Mf, (uint64_t)h->d * 2, 64, 32))) return rc;
  if ((rc = make_token_out_maps(&h->tm_tok_hi, &h->tm_tok_lo, h->buf.Xb, h->buf.Xlo, h->B, h->T, h->S, R_PREFIX, h->d))) return rc;
  h->bound = true;
  return TAMF_OK;
}

extern "C" int tamf_refiner_forward(tamf_refiner* h, const float* sample_pose_repr, const float* h2o_dist,
                                    const int32_t* hand_side, const float* shape, const float* obj_traj,
                                    const float* obj_emb, int nobj_max, float* refine_out, void* stream_) {
  cudaStream_t s = (cudaStream_t)stream_;
  TAMF_REQUIRE(h && h->bound, TAMF_E_STATE, "tamf_refiner_forward: bind a workspace first");
  TAMF_REQUIRE(sample_pose_repr && h2o_dist && hand_side && shape && obj_traj && obj_emb && refine_out, TAMF_E_BADARG,
               "tamf_refiner_forward: null pointer");
  TAMF_REQUIRE(nobj_max >= 1, TAMF_E_BADARG, "tamf_refiner_forward: nobj_max must be >= 1");
  const int d = h->d, B = h->B, T = h->T, S = h->S, M = h->M, Mf = h->Mf;
  const tamf_cfg& c = h->cfg;
  int rc;
  if ((rc = encoder_begin_evaluation(h->buf, s))) return rc;
  // ---- conditioning (hand_shape_process :300-301, obj_embed_process :260-261, obj_input_process :243-246) ----
  if ((rc = mean_axis(shape, h->shapemean, B, T, c.hand_shape_dim, s))) return rc;
  if ((rc = mean_axis(obj_emb, h->embmean, B, nobj_max, c.obj_embed_dim, s))) return rc;
  if ((rc = traj_mean(obj_traj, h->trajmean, B, nobj_max, T, s))) return rc;
  if ((rc = linear_f32(h->shapemean, c.hand_shape_dim, h->shape_w, c.hand_shape_dim, h->shape_b, h->prefix + 1 * d,
                       R_PREFIX * d, B, d, c.hand_shape_dim, 0, nullptr, 0, s)))
    return rc;
  if ((rc = linear_f32(h->embmean, c.obj_embed_dim, h->objemb_w, c.obj_embed_dim, h->objemb_b, h->prefix + 2 * d,
                       R_PREFIX * d, B, d, c.obj_embed_dim, 0, nullptr, 0, s)))
    return rc;
  {
    const size_t n = (size_t)B * R_PREFIX * d;
    r_prefix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(h->prefix, hand_side, h->pe, h->buf.Xlo, h->buf.Xb, B, S, d);
    TAMF_LAUNCH_CHECK();
  }
  // ---- frame tokens ----
  {
    const size_t n = (size_t)Mf * (R_K / 2);
    r_prep_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(sample_pose_repr, h->trajmean, h2o_dist, h->A0, Mf,
                                                              h->nfeat);
    TAMF_LAUNCH_CHECK();
  }
  {
    GemmParams p{};
    p.M = Mf, p.N = d, p.K = R_K, p.bias = h->merge_bias, p.out_bf16 = h->H0, p.ld_bf16 = d, p.tmC = &h->tm_H0_st;
    if ((rc = launch_gemm<256, EPI_BIAS_SILU_BF16, 2>(h->tm_A0, h->tm_wfold, p, s))) return rc;
  }
  {
    GemmParams p{};
    p.M = Mf, p.N = d, p.K = d, p.bias = h->b_m2, p.pe = h->pe, p.T = T, p.S = S, p.P0 = R_PREFIX, p.Xlo = h->buf.Xlo,
    p.Xb = h->buf.Xb;
    p.tmC = &h->tm_tok_hi, p.tmX = &h->tm_tok_lo;
    if ((rc = launch_gemm<256, EPI_TOKEN_OUT, 2>(h->tm_H0, h->tm_wm2, p, s))) return rc;
  }
  if ((rc = enqueue_encoder(h->enc, h->buf, s, nullptr))) return rc;
  {
    PdlBlock no_early_launch(h->buf.stack);  // (see denoiser.cu: the kernel after the encoder's stack form)
    GemmParams p{};
    p.M = M, p.N = h->nfeat, p.K = d, p.bias = h->b_fin, p.T = T, p.S = S, p.P0 = R_PREFIX, p.nfeat = h->nfeat;
    p.x_t = sample_pose_repr, p.x_out = refine_out;
    if ((rc = launch_gemm<128, EPI_RESIDUAL_OUT, 2>(h->buf.tm_Xb, h->tm_wfin, p, s))) return rc;
  }
  return TAMF_OK;
}
