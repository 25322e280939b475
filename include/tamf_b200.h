/* tamf_b200.h -- C ABI of libtamf_b200.so: the B200 (sm_100a) sampling hot path of OakInk2-TaMF.
 *
 * The reference has no FFI for this path (it is pure Python over torch / pytorch3d); each entry point
 * below names the reference interface it replaces (paths relative to the reference tree).  The Python
 * drop-in classes in oakink2-tamf_b200/tamf_b200/ bind these with ctypes (INTEGRATION.md shows the
 * stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types.  Unless a name ends in `_host`, pointers are
 *     DEVICE pointers, contiguous, 16-byte aligned, owned by the caller.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous.
 *   - every entry returns 0 on success or a negative TAMF_E_* code; tamf_last_error() gives the
 *     thread-local message.  Nothing throws across the ABI.
 *   - handles are not thread-safe; one handle per device per process (the reference is one process
 *     per GPU, launch/sample.py:272-289).
 *   - the library refuses to run on anything but compute capability 10.x (TAMF_E_ARCH): there is no
 *     CPU fallback.
 */
#ifndef TAMF_B200_H_
#define TAMF_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAMF_OK 0
#define TAMF_E_BADARG (-1) /* bad shape / null pointer / unsupported size */
#define TAMF_E_CUDA (-2)   /* a CUDA runtime or driver call failed */
#define TAMF_E_ARCH (-3)   /* device is not sm_100 */
#define TAMF_E_ALIGN (-4)  /* pointer not 16-byte aligned */
#define TAMF_E_STATE (-5)  /* call order violated (e.g. forward before bind / set_cond) */

int tamf_version(void);
const char* tamf_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Chamfer nearest neighbour (K=1).
 * Replaces pytorch3d.ops.knn_points(x, y, K=1) as called by
 *   thirdparty/chamfer_distance/chamfer_distance/chamfer_distance.py:147-148 (ChamferDistance.forward)
 * and consumed by src/oakink2_tamf/model/loss/chamfer_distance.py:36-62 (point2point_signed).
 * x [N,P1,3], y [N,P2,3] fp32 -> d2 [N,P1] fp32 (squared L2), idx [N,P1] int64.
 * Arithmetic: (dx*dx + dy*dy) + dz*dz, each op rounded to fp32 (no FMA), lowest index on ties.
 * `idx` doubles as the packed (d2,idx) scratch during the call.                                    */
int tamf_nn_query(const float* x, const float* y, int N, int P1, int P2, float* d2, int64_t* idx, void* stream);

/* Fused hand->object distance.
 * Replaces SegmentRefineModel.multi_object_h2o_dist (src/oakink2_tamf/model/segment_refine_model.py:142-168):
 * per sequence b and frame t the canonical clouds of its objects are moved by tslrot6d_to_transf /
 * transf_point_array (src/dev_fn/transform/transform.py:148-154,36-53), concatenated, and every hand
 * vertex gets the unsigned distance to its nearest object point.
 *   verts     [B,T,V,3]            hand vertices (world)
 *   obj_traj  [B,nobj_max,T,9]     tsl(3)+rot6d(6), zero padded over objects
 *   obj_points[sum_b nobj_b, P, 3] canonical clouds, packed in batch order
 *   obj_first [B+1] int32          prefix sums of nobj_b (host pointer)
 *   dist      [B,T,V] fp32         |v - nearest|  (sqrt of the squared distance)
 *   idx       [B,T,V] int64        index into the concatenated cloud of that sequence (scratch + output)
 * Device scratch: the B+1 prefix sums (and, for tamf_nn_query with more than 4096 queries per cloud, the packed
 * per-query minima) live in a grow-only buffer the library keeps per (device, stream): the FIRST call on a stream, or
 * a call that needs more than any earlier one, allocates (cudaMalloc); steady-state calls do not. */
int tamf_h2o_dist(const float* verts, const float* obj_traj, const float* obj_points, const int32_t* obj_first_host,
                  int B, int T, int V, int nobj_max, int P, float* dist, int64_t* idx, void* stream);
/* Two-step form for callers that query the same objects several times (SegmentRefineModel.forward runs the query for
 * the sampled, refined and target hands): build the block index of the canonical clouds once into caller-owned device
 * memory, then query it.  tamf_h2o_index_bytes returns 0 when P > 8192 (no index: use tamf_h2o_dist). */
size_t tamf_h2o_index_bytes(int total_obj, int P);
int tamf_h2o_index_build(const float* obj_points, int total_obj, int P, void* index, size_t index_bytes, void* stream);
int tamf_h2o_dist_indexed(const float* verts, const float* obj_traj, const void* index, const int32_t* obj_first_host,
                          int B, int T, int V, int nobj_max, int P, float* dist, int64_t* idx, void* stream);

/* Same contract, exhaustive scan of every (vertex, point) pair.  tamf_h2o_dist rejects whole blocks of 64 points by
 * their bounding boxes in the object frame when P <= 8192 (exact: bit-identical dist / idx, csrc/nn.cu); this entry
 * is the cross-check the parity tests compare it with and the path taken for larger P. */
int tamf_h2o_dist_exhaustive(const float* verts, const float* obj_traj, const float* obj_points, const int32_t* obj_first_host,
                  int B, int T, int V, int nobj_max, int P, float* dist, int64_t* idx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MANO forward kinematics.
 * Replaces manotorch ManoLayer(rot_mode="quat", center_idx=0, use_pca=False, flat_hand_mean=True)
 *   thirdparty/manotorch/manotorch/manolayer.py:39-98 (buffers), :128-266 (skinning_layer), :268-285 (forward)
 * and the pose_repr front end of SegmentRefineModel.batch_recover_mano_from_pose_repr
 *   src/oakink2_tamf/model/segment_refine_model.py:117-131.
 * Asset pointers are HOST fp32 arrays with the ManoLayer buffer shapes:
 *   shapedirs [778,3,10], posedirs [778,3,135], v_template [778,3], J_regressor [16,778], weights [778,16]. */
typedef struct tamf_mano tamf_mano;
int tamf_mano_create(const float* shapedirs_host, const float* posedirs_host, const float* v_template_host,
                     const float* j_regressor_host, const float* weights_host, int is_right, tamf_mano** out);
int tamf_mano_destroy(tamf_mano* h);
#define TAMF_POSE_QUAT 0      /* pose [N,16,4] quaternions (w,x,y,z): ManoLayer.forward(pose_coeffs, betas)   */
#define TAMF_POSE_REPR 1      /* pose [N,99] = tsl(3) + 16 x rot6d: batch_recover_mano_from_pose_repr (+tsl)  */
/* betas [N,10]; verts [N,778,3]; joints [N,21,3] (root-centred; +tsl in TAMF_POSE_REPR mode). */
int tamf_mano_fk(const tamf_mano* h, int pose_mode, const float* pose, const float* betas, int N, float* verts,
                 float* joints, void* stream);
/* Same kernel with the remaining MANOOutput fields (manolayer.py:242-265): center_joint [N,3] = the root joint before
 * the centre shift, transforms_abs [N,16,4,4] = global joint transforms G_k with the translation centre-shifted (no tsl).
 * Either may be null. */
int tamf_mano_fk_full(const tamf_mano* h, int pose_mode, const float* pose, const float* betas, int N, float* verts,
                      float* joints, float* center_joint, float* transforms_abs, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MF-MDM G denoiser + ancestral DDPM sampler.
 * Replaces InterationSegmentMDM.forward(x, timesteps, batch)
 *   src/oakink2_tamf/model/interaction_segment_mdm.py:134-174 (+ sub-modules :181-318)
 * and GaussianDiffusion.p_sample / p_sample_loop
 *   src/oakink2_tamf/model/diffusion/gaussian_diffusion.py:412-460, 506-640 (START_X, FIXED_SMALL).  */
typedef struct tamf_denoiser tamf_denoiser;

typedef struct tamf_cfg {
  int32_t input_dim;      /* 99  */
  int32_t obj_input_dim;  /* 9   */
  int32_t hand_shape_dim; /* 10  */
  int32_t obj_embed_dim;  /* 768 */
  int32_t latent_dim;     /* 256 | 512 */
  int32_t ff_size;        /* 1024 | 2048 */
  int32_t num_layers;     /* 8 */
  int32_t num_heads;      /* 4 */
  int32_t clip_dim;       /* 512 */
  int32_t num_steps;      /* 1000 diffusion steps (cosine schedule is built by the caller) */
} tamf_cfg;

/* HOST fp32 weight pointers, reference state_dict names in comments. */
typedef struct tamf_layer_weights {
  const float *in_proj_w, *in_proj_b;   /* seqTransEncoder.layers.L.self_attn.in_proj_{weight,bias}  [3d,d],[3d] */
  const float *out_proj_w, *out_proj_b; /* ...self_attn.out_proj.{weight,bias}                         [d,d],[d]  */
  const float *lin1_w, *lin1_b;         /* ...linear1.{weight,bias}                                     [ff,d],[ff]*/
  const float *lin2_w, *lin2_b;         /* ...linear2.{weight,bias}                                     [d,ff],[d] */
  const float *norm1_w, *norm1_b;       /* ...norm1.{weight,bias}                                       [d]        */
  const float *norm2_w, *norm2_b;       /* ...norm2.{weight,bias}                                       [d]        */
} tamf_layer_weights;

typedef struct tamf_g_weights {
  const float *shape_w, *shape_b;         /* hand_shape_process.shape_embed      [d,10]   */
  const float *objemb_w, *objemb_b;       /* obj_embed_process.embedding         [d,768]  */
  const float *pose_w, *pose_b;           /* input_process.poseEmbedding         [d,99]   */
  const float *objtraj_w, *objtraj_b;     /* obj_input_process.poseEmbedding     [d,9]    */
  const float *merge0_w, *merge0_b;       /* input_merge.0                       [d,2d]   */
  const float *merge2_w, *merge2_b;       /* input_merge.2                       [d,d]    */
  const float *time0_w, *time0_b;         /* embed_timestep.time_embed.0         [d,d]    */
  const float *time2_w, *time2_b;         /* embed_timestep.time_embed.2         [d,d]    */
  const float *text_w, *text_b;           /* embed_text                          [d,clip] */
  const float *final_w, *final_b;         /* output_process.poseFinal            [99,d]   */
  const float *pe;                        /* sequence_pos_encoder.pe             [>=max(num_steps,5+T), d] */
  int32_t pe_rows;
  const tamf_layer_weights* layers;       /* [num_layers] */
  /* float64 schedule tables, [num_steps] each (GaussianDiffusion.__init__, gaussian_diffusion.py:149-157) */
  const double *posterior_mean_coef1, *posterior_mean_coef2, *posterior_log_variance_clipped;
} tamf_g_weights;

int tamf_denoiser_create(const tamf_cfg* cfg, const tamf_g_weights* w, tamf_denoiser** out);
int tamf_denoiser_destroy(tamf_denoiser* h);

/* Workspace protocol: the library never allocates in the hot path. */
size_t tamf_denoiser_workspace_bytes(const tamf_denoiser* h, int B, int T);
int tamf_denoiser_bind(tamf_denoiser* h, int B, int T, void* workspace, size_t workspace_bytes);

/* Per-sample conditioning, computed ONCE per batch (constant over the reverse chain):
 *   text_feat [B,clip_dim]        = clip_model.encode_text(tokens).float()   (interaction_segment_mdm.py:132)
 *   hand_side [B] int32           0 = "rh", 1 = "lh"                          (:270-274)
 *   shape     [B,T,10], obj_traj [B,nobj_max,T,9], obj_emb [B,nobj_max,768]  (zero padded over objects) */
int tamf_denoiser_set_cond(tamf_denoiser* h, const float* text_feat, const int32_t* hand_side, const float* shape,
                           const float* obj_traj, const float* obj_emb, int nobj_max, void* stream);

/* x0 = model(x_t, t): x_t, x0_out [B,99,1,T] fp32; t [B] int32 (device). */
int tamf_denoiser_forward(tamf_denoiser* h, const float* x_t, const int32_t* t, float* x0_out, void* stream);

/* Installs the K-step update rule the sampler entries below run, in place of the ancestral rule given at create:
 *   x_{i-1} = c1[i] x0 + c2[i] x_i + sigma[i] eps,   x0 = model(x_i, timestep_map[i]),   i = K-1 .. 0.
 * Covers SpacedDiffusion's strided schedules (respace.py:8-57 space_timesteps, :69-83 re-derived betas, :114-119
 * timestep_map) with either the ancestral posterior (gaussian_diffusion.py:209-229) or DDIM with any eta
 * (ddim_sample, gaussian_diffusion.py:642-690: eps re-derived from x0 folds into c1/c2).  HOST fp32 / int32 arrays of
 * length K <= num_steps; K = 0 restores the ancestral rule.  Synchronises the device; not a hot-path call.
 * tamf_denoiser_forward is unaffected (it takes ORIGINAL timesteps). */
int tamf_denoiser_set_sampler(tamf_denoiser* h, int K, const float* c1, const float* c2, const float* sigma,
                              const int32_t* timestep_map);
int tamf_denoiser_sampler_steps(const tamf_denoiser* h);

/* One ancestral step p_sample (all rows at the same t):
 *   x_{t-1} = c1[t] x0 + c2[t] x_t + 1[t!=0] exp(0.5 logvar[t]) eps,   x0 = model(x_t, t)
 * x_io is updated in place.  eps = `noise` [B,99,1,T] if non-NULL, else Philox4x32-10(seed, t, element) normals.
 * x0_out may be NULL. */
int tamf_p_sample_step(tamf_denoiser* h, float* x_io, int t, const float* noise, uint64_t seed, float* x0_out,
                       void* stream);

/* Steps t_start, t_start-1, ..., t_end (inclusive) with in-kernel Philox noise, replayed from one captured
 * CUDA graph per step (p_sample_loop_progressive, gaussian_diffusion.py:621-640).  x_io [B,99,1,T] in place. */
int tamf_p_sample_chain(tamf_denoiser* h, float* x_io, int t_start, int t_end, uint64_t seed, void* stream);

/* Whole p_sample_loop with HOST buffers (what the e2e figure times): copies the conditioning in, draws
 * x_T ~ N(0,I) (Philox, counter t = num_steps) unless x_T_host is given, runs the chain, copies the sample out.
 * Synchronous. */
int tamf_p_sample_loop_host(tamf_denoiser* h, const float* text_feat_host, const int32_t* hand_side_host,
                            const float* shape_host, const float* obj_traj_host, const float* obj_emb_host,
                            int nobj_max, const float* x_T_host, uint64_t seed, float* sample_out_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * MF-MDM R (refine) transformer pass.
 * Replaces the tensor part of SegmentRefineModel.forward
 *   src/oakink2_tamf/model/segment_refine_model.py:170-217
 * (prefix tokens :177-186, input / object / h2o-distance embeddings + input_merge :189-208, positional encoding,
 * 8-layer encoder, output_process, x_in + output, nan_to_num :211-217).  FK and hand->object distances of the same
 * forward (:193-201, :220-232) are tamf_mano_fk(_select) and tamf_h2o_dist; tamf_b200/refine.py sequences them.   */
typedef struct tamf_refiner tamf_refiner;

typedef struct tamf_r_weights {
  const float *shape_w, *shape_b;     /* hand_shape_process.shape_embed          [d,10]   */
  const float *objemb_w, *objemb_b;   /* obj_embed_process.embedding             [d,768]  */
  const float *pose_w, *pose_b;       /* input_process.poseEmbedding             [d,99]   */
  const float *objtraj_w, *objtraj_b; /* obj_input_process.poseEmbedding         [d,9]    */
  const float *dist_w, *dist_b;       /* h2o_dist_input_process.poseEmbedding    [d,778]  */
  const float *merge0_w, *merge0_b;   /* input_merge.0                           [d,3d]   */
  const float *merge2_w, *merge2_b;   /* input_merge.2                           [d,d]    */
  const float *final_w, *final_b;     /* output_process.poseFinal                [99,d]   */
  const float *pe;                    /* sequence_pos_encoder.pe                 [>= 3+T, d] */
  int32_t pe_rows;
  const tamf_layer_weights* layers;   /* [num_layers] */
} tamf_r_weights;

/* cfg: clip_dim / num_steps are ignored. */
int tamf_refiner_create(const tamf_cfg* cfg, const tamf_r_weights* w, tamf_refiner** out);
int tamf_refiner_destroy(tamf_refiner* h);
size_t tamf_refiner_workspace_bytes(const tamf_refiner* h, int B, int T);
int tamf_refiner_bind(tamf_refiner* h, int B, int T, void* workspace, size_t workspace_bytes);
/* sample_pose_repr [B,T,99]; h2o_dist [B,T,778] (of the sampled pose); hand_side [B] int32 (0 rh, 1 lh);
 * shape [B,T,10]; obj_traj [B,nobj_max,T,9]; obj_emb [B,nobj_max,768] -> refine_out [B,T,99] = x_in + delta. */
int tamf_refiner_forward(tamf_refiner* h, const float* sample_pose_repr, const float* h2o_dist,
                         const int32_t* hand_side, const float* shape, const float* obj_traj, const float* obj_emb,
                         int nobj_max, float* refine_out, void* stream);

/* FK of the frames listed in frame_ids [n] (device int32) only, results written at those frame indices of the full
 * verts [N,778,3] / joints [N,21,3] arrays: lets one batch mix right and left hands without gather/scatter copies
 * (the reference loops over batch items, segment_refine_model.py:113-129).  pose/betas are the FULL [N,..] arrays. */
int tamf_mano_fk_select(const tamf_mano* h, int pose_mode, const float* pose, const float* betas,
                        const int32_t* frame_ids, int n, float* verts, float* joints, void* stream);

/* Area-weighted vertex normals (pytorch3d Meshes.verts_normals_packed as used at segment_refine_model.py:132-133;
 * semantics restated in oracle/tamf_oracle.py:vertex_normals -- parity unpinned, pytorch3d is absent).
 * verts [N,V,3] fp32, faces [F,3] int32 -> normals [N,V,3]. */
int tamf_vertex_normals(const float* verts, const int32_t* faces, int N, int V, int F, float* normals, void* stream);

/* Measurement aid (bench.py roofline): runs ONE p_sample step at timestep t eagerly (no graph) with a CUDA event
 * between consecutive kernels and returns the per-kernel device times in launch order:
 *   prep, embed-a, embed-b, L x {in_proj, attention, out_proj+LN1, linear1+GELU, linear2+LN2}, final+posterior.
 * ms_out [cap] HOST floats, *n_out = number of kernels (3 + 5 L + 1).  x_io is advanced one step.  Synchronous. */
int tamf_denoiser_profile_step(tamf_denoiser* h, float* x_io, int t, uint64_t seed, float* ms_out_host, int cap,
                               int* n_out_host, void* stream);
/* In-graph per-kernel breakdown (bench.py `roofline.kernels_in_graph`): the step is captured in a CUDA graph exactly as
 * tamf_p_sample_chain captures it (same kernels, same programmatic dependent launches); every kernel additionally
 * records, on the globaltimer all SMs share, its earliest CTA entry, the earliest end of a dependency wait and its latest
 * CTA exit.  n_steps (>= 4) replays run back to back from t_start downwards; the last three are averaged.  Outputs per
 * kernel in launch order (microseconds relative to the entry of the step's first kernel), n_out kernels, and the mean
 * step period step_us.  x_io is advanced n_steps sampler steps. */
int tamf_denoiser_profile_graph(tamf_denoiser* h, float* x_io, int t_start, int n_steps, uint64_t seed, double* entry_us,
                                double* ready_us, double* exit_us, int cap, int* n_out, double* step_us, void* stream);

/* Number of kernels this library launched since load (for bench.py's gpu_launches claim). */
uint64_t tamf_kernel_launch_count(void);

/* The raw Philox normal generator used by the sampler, exposed for parity tests: out [n] fp32. */
int tamf_philox_normal(float* out, size_t n, uint64_t seed, uint32_t t, void* stream);

/* Self-test of the tcgen05 GEMM against caller-provided data: C[M,N] fp32 = A[M,K] bf16 . W[N,K]^T bf16 + bias.
 * a, w are device bf16 (uint16 bit patterns), bias device fp32 [N] or NULL, c device fp32.
 * tile_n in {128,256,512}; cta_group 1 (one CTA per 128-row tile) or 2 (CTA pair, tcgen05.mma.cta_group::2). */
int tamf_gemm_selftest(const uint16_t* a, const uint16_t* w, const float* bias, float* c, int M, int N, int K,
                       int tile_n, int cta_group, void* stream);

/* Self-test of the fused attention kernel: qkv device bf16 [B*S, 3d] (rows = tokens, columns q | k | v, head h owns
 * columns h*hd .. of each, the layout F.multi_head_attention_forward derives from in_proj; torch
 * nn/functional.py), out device bf16 [B*S, d] = softmax(q k^T / sqrt(hd)) v per (sequence, head), no mask
 * (interaction_segment_mdm.py:63-70,171).  S <= 176, hd = d/H in {64, 128}. */
int tamf_attn_selftest(const uint16_t* qkv, uint16_t* out, int B, int S, int H, int d, void* stream);

/* Debug aid (tools/attn_trace.py): tamf_attn_selftest with per-CTA clock64 stamps of the kernel's phases,
 * trace int64 [H*B][16] (device): 0 start, 1 setup done, 2 Q+K landed, 3 V landed, 4/5 P of tile 0/1 ready (MMA thread),
 * 6/7 scores of tile 0/1 ready, 8/9 O ready, 10/11 tile stored, 12 end, 13 SM id, 14/15 globaltimer start/end. */
int tamf_attn_trace(const uint16_t* qkv, uint16_t* out, int B, int S, int H, int d, long long* trace, void* stream);

/* Debug aid (tools/gemm_trace.py): one launch of a hot-path GEMM shape with per-CTA clock64 event timestamps.
 * which: 0 = in_proj-like (bias -> bf16), 1 = linear1-like (bias + GELU -> bf16), 2 = LayerNorm GEMM (N = 512).
 * a [M,K], w [N,K] bf16; bias [N]; out bf16 [M,N]; X fp32 [M,N] (which 2); trace int64 [148][64] (device). */
int tamf_gemm_trace(int which, const uint16_t* a, const uint16_t* w, const float* bias, void* out, float* X, int M,
                    int N, int K, long long* trace, void* stream);

/* Debug / self-test aid (tools/layer_trace.py, tests/test_gemm_gpu.py): ONE launch of the layer kernel
 * (csrc/layer_chain.cuh: everything between two attention kernels of the encoder stack in one persistent kernel) on
 * caller data:  X = LN1(X + att . w_out^T + b_out);  H = gelu(Xh . w1^T + b1);  X = LN2(X + H . w2^T + b2);
 * qkv = Xh . w_in^T + b_in (n_inp = 3 d; n_inp = 0 skips it, as after the last layer).
 * att [M,d], w_out [d,d], w1 [ff,d], w2 [d,ff], w_in [3d,d] bf16; ln_params = [b_out | g1 | be1 | b2 | g2 | be2] 6*d fp32;
 * X = Xh + Xl, two bf16 planes [M,d], updated in place; Hbuf bf16 [M,ff]; qkv bf16 [M,3d]; aux: tamf_layer_aux_bytes(M, d,
 * ff) bytes of device scratch; trace: null or int64 [148][64] per-CTA clock64 stamps. */
size_t tamf_layer_aux_bytes(int M, int d, int ff);
int tamf_layer_run(const uint16_t* att, const uint16_t* w_out, const uint16_t* w1, const uint16_t* w2, const uint16_t* w_in,
                   const float* ln_params, const float* b1, const float* b_in, uint16_t* Xh, uint16_t* Xl, uint16_t* Hbuf,
                   uint16_t* qkv, int M, int d, int ff, int n_inp, void* aux, size_t aux_bytes, long long* trace,
                   void* stream);
/* Debug aid (tools/chain_trace_model.py): the layer kernel of encoder layer `layer` writes per-CTA clock64 stamps (int64
 * [148][64], device) on every later launch of any handle in this process; a null pointer switches it off. */
int tamf_debug_chain_trace(long long* trace, int layer);
/* Host-only (no GPU needed): the static schedule the layer kernel runs for an [M, d] problem on `slots` CTA pairs.
 * off_out [slots + 1], units_out [cap] unit codes kind << 28 | row tile << 8 | column tile (kind 0 LN1, 1 L1, 2 LN2,
 * 3 INP); makespan_out: the cost model's estimate in cycles.  Returns the pair count, or a negative TAMF_E_* code. */
int tamf_layer_schedule(int M, int d, int ff, int n_inp, int slots, int* off_out, int* units_out, int cap,
                        double* makespan_out);
/* Same for the stack form (TAMF_CHAIN=2): all `layers` layers in ONE launch on `slots` pairs next to a persistent attention
 * kernel of `att_ctas` CTAs; unit codes kind << 28 | layer << 24 | row tile << 8 | column tile. */
int tamf_stack_schedule(int M, int d, int ff, int layers, int slots, int att_ctas, int S, int heads, double att_unit,
                        int* off_out, int* units_out, int cap, double* makespan_out);

#ifdef __cplusplus
}
#endif
#endif /* TAMF_B200_H_ */
