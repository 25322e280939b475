// mano.cu -- MANO forward kinematics / linear blend skinning, one CTA per tile of FK_FT frames.
//
// Replaces manotorch ManoLayer.forward in quat mode, center_idx=0
//   thirdparty/manotorch/manotorch/manolayer.py:119-126 (rotation_by_quaternion), :128-266 (skinning_layer)
// and the pose_repr front end of SegmentRefineModel.batch_recover_mano_from_pose_repr
//   src/oakink2_tamf/model/segment_refine_model.py:117-131 (rot6d -> rotmat -> quat -> ManoLayer, + tsl).
//
// Data layout in HBM (built once by tamf_mano_create from the ManoLayer buffers):
//   blendT [145][778*3]  rows 0..134 = posedirs^T, rows 135..144 = shapedirs^T  (coalesced over the vertex axis)
//   vt     [778*3]       v_template
//   J0     [16*3]        J_regressor . v_template          (joint regression folded through the shape basis:
//   JS     [16*3][10]    J_regressor . shapedirs            J = J0 + JS . betas, manolayer.py:139-142)
//   wT     [16][778]     skinning weights^T
// A CTA stages the per-frame features (135 pose + 10 shape coefficients) of its FK_FT frames in shared memory,
// then each thread owns vertices v, v+256, ... and keeps 3 x FK_FT accumulators in registers, so blendT is
// streamed from L2 once per FK_FT frames.  The kinematic chain (16 joints, 3 levels) runs on 16 threads/frame.
#include "common.cuh"

namespace tamf {

constexpr int FK_FT = 8;       // frames per CTA
constexpr int FK_THREADS = 256;
constexpr int NV = 778;
constexpr int NFEAT = 145;     // 135 pose-blend + 10 shape coefficients

struct tamf_mano_impl {
  float *blendT, *vt, *J0, *JS, *wT;
  int is_right;
};

// rotation.py:446-467 -> rotation.py:167-213 (rotmat_to_quat, standardised w>=0) -> geometry.py:225-253
__device__ __forceinline__ void rot6d_to_R_via_quat(const float* d6, float (&R)[9]) {
  float a1x = d6[0], a1y = d6[1], a1z = d6[2], a2x = d6[3], a2y = d6[4], a2z = d6[5];
  float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);
  float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  float dt = b1x * a2x + b1y * a2y + b1z * a2z;
  float b2x = a2x - dt * b1x, b2y = a2y - dt * b1y, b2z = a2z - dt * b1z;
  float n2 = fmaxf(sqrtf(b2x * b2x + b2y * b2y + b2z * b2z), 1e-12f);
  b2x /= n2, b2y /= n2, b2z /= n2;
  float m00 = b1x, m01 = b1y, m02 = b1z, m10 = b2x, m11 = b2y, m12 = b2z;
  float m20 = b1y * b2z - b1z * b2y, m21 = b1z * b2x - b1x * b2z, m22 = b1x * b2y - b1y * b2x;
  // rotmat_to_quat
  float qa[4] = {1.0f + m00 + m11 + m22, 1.0f + m00 - m11 - m22, 1.0f - m00 + m11 - m22, 1.0f - m00 - m11 + m22};
#pragma unroll
  for (int i = 0; i < 4; ++i) qa[i] = qa[i] > 0.f ? sqrtf(qa[i]) : 0.f;
  int best = 0;  // torch.argmax: first maximal index
#pragma unroll
  for (int i = 1; i < 4; ++i)
    if (qa[i] > qa[best]) best = i;
  float c0, c1, c2, c3;
  if (best == 0) {
    c0 = qa[0] * qa[0], c1 = m21 - m12, c2 = m02 - m20, c3 = m10 - m01;
  } else if (best == 1) {
    c0 = m21 - m12, c1 = qa[1] * qa[1], c2 = m10 + m01, c3 = m02 + m20;
  } else if (best == 2) {
    c0 = m02 - m20, c1 = m10 + m01, c2 = qa[2] * qa[2], c3 = m12 + m21;
  } else {
    c0 = m10 - m01, c1 = m20 + m02, c2 = m21 + m12, c3 = qa[3] * qa[3];
  }
  float den = 2.0f * fmaxf(qa[best], 0.1f);
  float r = c0 / den, i = c1 / den, j = c2 / den, k = c3 / den;
  if (r < 0.f) r = -r, i = -i, j = -j, k = -k;
  // quaternion_to_matrix (un-normalised)
  float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  R[0] = 1 - two_s * (j * j + k * k), R[1] = two_s * (i * j - k * r), R[2] = two_s * (i * k + j * r);
  R[3] = two_s * (i * j + k * r), R[4] = 1 - two_s * (i * i + k * k), R[5] = two_s * (j * k - i * r);
  R[6] = two_s * (i * k - j * r), R[7] = two_s * (j * k + i * r), R[8] = 1 - two_s * (i * i + j * j);
}

__device__ __forceinline__ void quat_to_R(const float* q, float (&R)[9]) {
  float r = q[0], i = q[1], j = q[2], k = q[3];
  float two_s = 2.0f / (r * r + i * i + j * j + k * k);
  R[0] = 1 - two_s * (j * j + k * k), R[1] = two_s * (i * j - k * r), R[2] = two_s * (i * k + j * r);
  R[3] = two_s * (i * j + k * r), R[4] = 1 - two_s * (i * i + k * k), R[5] = two_s * (j * k - i * r);
  R[6] = two_s * (i * k - j * r), R[7] = two_s * (j * k + i * r), R[8] = 1 - two_s * (i * i + j * j);
}

// parent of joint k in MANO order (root -> 5 fingers x 3 levels; manolayer.py:164-193)
__device__ __forceinline__ int fk_parent(int k) { return (k == 0) ? -1 : ((k - 1) % 3 == 0 ? 0 : k - 1); }

__global__ void __launch_bounds__(FK_THREADS)
    mano_fk_kernel(const float* __restrict__ blendT, const float* __restrict__ vt, const float* __restrict__ J0,
                   const float* __restrict__ JS, const float* __restrict__ wT, int is_right, int pose_mode,
                   const float* __restrict__ pose, const float* __restrict__ betas, const int* __restrict__ frame_ids,
                   int N, float* __restrict__ verts, float* __restrict__ joints, float* __restrict__ center_out,
                   float* __restrict__ transf_out) {
  __shared__ __align__(16) float sFeat[NFEAT][FK_FT];      // [feature][frame]: one LDS.128 x2 feeds 8 frames
  __shared__ float sR[FK_FT][16][9];         // local rotations
  __shared__ float sJ[FK_FT][16][3];         // rest joints J
  __shared__ float sG[FK_FT][16][12];        // global transforms (3x4), then G' = G - [0 | G.J]
  __shared__ float sTsl[FK_FT][3];
  __shared__ float sCenter[FK_FT][3];
  // N frames are processed; slot i of this launch is frame frame_ids[i] of the full arrays (identity when null)
  __shared__ int sFrame[FK_FT];
  const int f0 = blockIdx.x * FK_FT;
  const int tid = threadIdx.x;
  if (tid < FK_FT) {
    const int i = min(f0 + tid, N - 1);
    sFrame[tid] = frame_ids ? frame_ids[i] : i;
  }
  __syncthreads();

  // ---- stage 1: rotations (FK_FT x 16 threads), shape coefficients, translations ----
  if (tid < FK_FT * 16) {
    const int fl = tid / 16, k = tid % 16, f = sFrame[fl];
    float R[9];
    if (pose_mode == TAMF_POSE_REPR)
      rot6d_to_R_via_quat(pose + (size_t)f * 99 + 3 + 6 * k, R);
    else
      quat_to_R(pose + (size_t)f * 64 + 4 * k, R);
#pragma unroll
    for (int e = 0; e < 9; ++e) sR[fl][k][e] = R[e];
    if (k >= 1) {
#pragma unroll
      for (int e = 0; e < 9; ++e) sFeat[(k - 1) * 9 + e][fl] = R[e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f);
    }
    if (k < 10) sFeat[135 + k][fl] = betas[(size_t)f * 10 + k];
    if (k < 3) sTsl[fl][k] = (pose_mode == TAMF_POSE_REPR) ? pose[(size_t)f * 99 + k] : 0.f;
  }
  __syncthreads();
  // ---- stage 2: rest joints J = J0 + JS . betas  (FK_FT x 48 values) ----
  for (int i = tid; i < FK_FT * 48; i += FK_THREADS) {
    const int fl = i / 48, c = i % 48;
    float acc = J0[c];
#pragma unroll
    for (int s = 0; s < 10; ++s) acc = fmaf(JS[c * 10 + s], sFeat[135 + s][fl], acc);
    sJ[fl][c / 3][c % 3] = acc;
  }
  __syncthreads();
  // ---- stage 3: kinematic chain, one thread per (frame, finger) walks its 3 levels; root first ----
  if (tid < FK_FT) {
    const int fl = tid;
#pragma unroll
    for (int e = 0; e < 9; ++e) sG[fl][0][(e / 3) * 4 + e % 3] = sR[fl][0][e];
#pragma unroll
    for (int r = 0; r < 3; ++r) sG[fl][0][r * 4 + 3] = sJ[fl][0][r];
  }
  __syncthreads();
  if (tid < FK_FT * 5) {
    const int fl = tid / 5, finger = tid % 5;
    for (int lev = 0; lev < 3; ++lev) {
      const int k = 1 + finger * 3 + lev, p = fk_parent(k);
      float rel[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) rel[r] = sJ[fl][k][r] - sJ[fl][p][r];
      float G[12];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float g0 = sG[fl][p][r * 4 + 0], g1 = sG[fl][p][r * 4 + 1], g2 = sG[fl][p][r * 4 + 2],
                    g3 = sG[fl][p][r * 4 + 3];
#pragma unroll
        for (int c = 0; c < 3; ++c) G[r * 4 + c] = g0 * sR[fl][k][c] + g1 * sR[fl][k][3 + c] + g2 * sR[fl][k][6 + c];
        G[r * 4 + 3] = g0 * rel[0] + g1 * rel[1] + g2 * rel[2] + g3;
      }
#pragma unroll
      for (int e = 0; e < 12; ++e) sG[fl][k][e] = G[e];
    }
  }
  __syncthreads();
  // ---- stage 4: joints out (16 chain joints; tips are written by the vertex owners), centre, G' ----
  if (tid < FK_FT * 16) {
    const int fl = tid / 16, k = tid % 16, f = sFrame[fl];
    const bool f_ok = f0 + fl < N;
    // reorder map of manolayer.py:240 : output slot of MANO joint k
    const int slot_of[16] = {0, 5, 6, 7, 9, 10, 11, 17, 18, 19, 13, 14, 15, 1, 2, 3};
    float g3[3], gj[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      g3[r] = sG[fl][k][r * 4 + 3];
      gj[r] = sG[fl][k][r * 4 + 0] * sJ[fl][k][0] + sG[fl][k][r * 4 + 1] * sJ[fl][k][1] +
              sG[fl][k][r * 4 + 2] * sJ[fl][k][2];
    }
    if (k == 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r) sCenter[fl][r] = g3[r];  // joints[:, center_idx=0]
    }
    __syncwarp();
    if (f_ok) {
      // centre = root joint translation = sG[fl][0][.][3] (read directly: sCenter may not be visible yet)
#pragma unroll
      for (int r = 0; r < 3; ++r)
        joints[((size_t)f * 21 + slot_of[k]) * 3 + r] = g3[r] - sG[fl][0][r * 4 + 3] + sTsl[fl][r];
      // MANOOutput.center_joint: the root joint before the centre shift (manolayer.py:242-245)
      if (center_out && k == 0) {
#pragma unroll
        for (int r = 0; r < 3; ++r) center_out[(size_t)f * 3 + r] = g3[r];
      }
      // MANOOutput.transforms_abs [16,4,4]: global joint transforms G_k, translation centre-shifted (:251-258)
      if (transf_out) {
        float* o = transf_out + ((size_t)f * 16 + k) * 16;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int c = 0; c < 3; ++c) o[r * 4 + c] = sG[fl][k][r * 4 + c];
          o[r * 4 + 3] = g3[r] - sG[fl][0][r * 4 + 3];
        }
        o[12] = 0.f, o[13] = 0.f, o[14] = 0.f, o[15] = 1.f;
      }
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 3; ++r) sG[fl][k][r * 4 + 3] = g3[r] - gj[r];  // manolayer.py:202-204
  }
  __syncthreads();
  // ---- stage 5: blend shapes + LBS, thread per vertex ----
  const int tipR[5] = {745, 317, 444, 556, 673};
  const int tip_slot[5] = {4, 8, 12, 16, 20};  // thumb, index, middle, ring, pinky tips after the :240 reorder
  for (int v = tid; v < NV; v += FK_THREADS) {
    float acc[3][FK_FT];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float t0 = vt[v * 3 + c];
#pragma unroll
      for (int fl = 0; fl < FK_FT; ++fl) acc[c][fl] = t0;
    }
    // T_P = v_template + B_S + B_P   (manolayer.py:139,154,157)
    for (int k = 0; k < NFEAT; ++k) {
      const float p0 = blendT[(size_t)k * (NV * 3) + v * 3 + 0];
      const float p1 = blendT[(size_t)k * (NV * 3) + v * 3 + 1];
      const float p2 = blendT[(size_t)k * (NV * 3) + v * 3 + 2];
      const float4 fa = *reinterpret_cast<const float4*>(&sFeat[k][0]);
      const float4 fb = *reinterpret_cast<const float4*>(&sFeat[k][4]);
      const float fv[8] = {fa.x, fa.y, fa.z, fa.w, fb.x, fb.y, fb.z, fb.w};
#pragma unroll
      for (int fl = 0; fl < FK_FT; ++fl) {
        acc[0][fl] = fmaf(p0, fv[fl], acc[0][fl]);
        acc[1][fl] = fmaf(p1, fv[fl], acc[1][fl]);
        acc[2][fl] = fmaf(p2, fv[fl], acc[2][fl]);
      }
    }
    float w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = wT[k * NV + v];
    int tip = -1;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      int tv = tipR[i];
      if (i == 2 && !is_right) tv = 445;  // manolayer.py:224-227
      if (v == tv) tip = i;
    }
#pragma unroll
    for (int fl = 0; fl < FK_FT; ++fl) {
      if (f0 + fl >= N) break;
      const int f = sFrame[fl];
      float Tm[12];
#pragma unroll
      for (int e = 0; e < 12; ++e) Tm[e] = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
#pragma unroll
        for (int e = 0; e < 12; ++e) Tm[e] = fmaf(w[k], sG[fl][k][e], Tm[e]);  // manolayer.py:208
      }
      float out[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        out[r] = Tm[r * 4 + 0] * acc[0][fl] + Tm[r * 4 + 1] * acc[1][fl] + Tm[r * 4 + 2] * acc[2][fl] + Tm[r * 4 + 3];
        out[r] = out[r] - sCenter[fl][r] + sTsl[fl][r];  // :242-249 centre shift, then + tsl (refine model :130)
      }
      verts[((size_t)f * NV + v) * 3 + 0] = out[0];
      verts[((size_t)f * NV + v) * 3 + 1] = out[1];
      verts[((size_t)f * NV + v) * 3 + 2] = out[2];
      if (tip >= 0) {
#pragma unroll
        for (int r = 0; r < 3; ++r) joints[((size_t)f * 21 + tip_slot[tip]) * 3 + r] = out[r];
      }
    }
  }
}

}  // namespace tamf

using namespace tamf;

struct tamf_mano {
  tamf_mano_impl d;
};

extern "C" int tamf_mano_create(const float* shapedirs, const float* posedirs, const float* v_template,
                                const float* j_regressor, const float* weights, int is_right, tamf_mano** out) {
  TAMF_REQUIRE(shapedirs && posedirs && v_template && j_regressor && weights && out, TAMF_E_BADARG,
               "tamf_mano_create: null pointer");
  int rc = check_device();
  if (rc) return rc;
  const int C = NV * 3;
  std::string err;
  float* h_blend = new float[(size_t)NFEAT * C];
  for (int c = 0; c < C; ++c) {
    for (int k = 0; k < 135; ++k) h_blend[(size_t)k * C + c] = posedirs[(size_t)c * 135 + k];
    for (int s = 0; s < 10; ++s) h_blend[(size_t)(135 + s) * C + c] = shapedirs[(size_t)c * 10 + s];
  }
  float h_J0[48], h_JS[480], *h_wT = new float[16 * NV];
  for (int j = 0; j < 16; ++j)
    for (int r = 0; r < 3; ++r) {
      double a = 0;  // fold in double, round once
      for (int v = 0; v < NV; ++v) a += (double)j_regressor[j * NV + v] * (double)v_template[v * 3 + r];
      h_J0[j * 3 + r] = (float)a;
      for (int s = 0; s < 10; ++s) {
        double b = 0;
        for (int v = 0; v < NV; ++v) b += (double)j_regressor[j * NV + v] * (double)shapedirs[(v * 3 + r) * 10 + s];
        h_JS[(j * 3 + r) * 10 + s] = (float)b;
      }
    }
  for (int v = 0; v < NV; ++v)
    for (int k = 0; k < 16; ++k) h_wT[k * NV + v] = weights[v * 16 + k];
  tamf_mano* h = new tamf_mano();
  h->d.is_right = is_right;
  auto up = [&](float** dptr, const float* src, size_t n) -> cudaError_t {
    cudaError_t e = cudaMalloc(dptr, n * sizeof(float));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dptr, src, n * sizeof(float), cudaMemcpyHostToDevice);
  };
  cudaError_t e = up(&h->d.blendT, h_blend, (size_t)NFEAT * C);
  if (e == cudaSuccess) e = up(&h->d.vt, v_template, C);
  if (e == cudaSuccess) e = up(&h->d.J0, h_J0, 48);
  if (e == cudaSuccess) e = up(&h->d.JS, h_JS, 480);
  if (e == cudaSuccess) e = up(&h->d.wT, h_wT, 16 * NV);
  delete[] h_blend;
  delete[] h_wT;
  if (e != cudaSuccess) {
    set_error(std::string("tamf_mano_create: ") + cudaGetErrorString(e));
    delete h;
    return TAMF_E_CUDA;
  }
  *out = h;
  return TAMF_OK;
}

extern "C" int tamf_mano_destroy(tamf_mano* h) {
  if (!h) return TAMF_OK;
  cudaFree(h->d.blendT);
  cudaFree(h->d.vt);
  cudaFree(h->d.J0);
  cudaFree(h->d.JS);
  cudaFree(h->d.wT);
  delete h;
  return TAMF_OK;
}

static int launch_fk(const tamf_mano* h, int pose_mode, const float* pose, const float* betas, const int32_t* frame_ids,
                     int n, float* verts, float* joints, cudaStream_t stream, const char* who,
                     float* center_out = nullptr, float* transf_out = nullptr) {
  TAMF_REQUIRE(h, TAMF_E_BADARG, std::string(who) + ": null handle");
  TAMF_REQUIRE(pose_mode == TAMF_POSE_QUAT || pose_mode == TAMF_POSE_REPR, TAMF_E_BADARG, std::string(who) + ": bad pose_mode");
  TAMF_REQUIRE(n >= 0, TAMF_E_BADARG, std::string(who) + ": negative frame count");
  if (n == 0) return TAMF_OK;
  TAMF_REQUIRE(pose && betas && verts && joints, TAMF_E_BADARG, std::string(who) + ": null pointer");
  mano_fk_kernel<<<(n + FK_FT - 1) / FK_FT, FK_THREADS, 0, stream>>>(h->d.blendT, h->d.vt, h->d.J0, h->d.JS, h->d.wT,
                                                                    h->d.is_right, pose_mode, pose, betas, frame_ids, n,
                                                                    verts, joints, center_out, transf_out);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

extern "C" int tamf_mano_fk(const tamf_mano* h, int pose_mode, const float* pose, const float* betas, int N,
                            float* verts, float* joints, void* stream) {
  return launch_fk(h, pose_mode, pose, betas, nullptr, N, verts, joints, (cudaStream_t)stream, "tamf_mano_fk");
}

extern "C" int tamf_mano_fk_full(const tamf_mano* h, int pose_mode, const float* pose, const float* betas, int N,
                                 float* verts, float* joints, float* center_joint, float* transforms_abs, void* stream) {
  return launch_fk(h, pose_mode, pose, betas, nullptr, N, verts, joints, (cudaStream_t)stream, "tamf_mano_fk_full",
                   center_joint, transforms_abs);
}

extern "C" int tamf_mano_fk_select(const tamf_mano* h, int pose_mode, const float* pose, const float* betas,
                                   const int32_t* frame_ids, int n, float* verts, float* joints, void* stream) {
  TAMF_REQUIRE(frame_ids || n == 0, TAMF_E_BADARG, "tamf_mano_fk_select: null frame_ids");
  return launch_fk(h, pose_mode, pose, betas, frame_ids, n, verts, joints, (cudaStream_t)stream, "tamf_mano_fk_select");
}

// ------------------------------------------------------------------------------------------------
// vertex normals: one CTA per mesh; face normals accumulated into shared memory, then normalised.
// cross(v2-v1, v0-v1) -> v1, cross(v0-v2, v1-v2) -> v2, cross(v1-v0, v2-v0) -> v0 (all three equal the face's
// area-weighted normal), normalize(eps 1e-6)  -- oracle/tamf_oracle.py:vertex_normals.
// ------------------------------------------------------------------------------------------------
namespace tamf {
// The per-vertex sums are accumulated in 2^-44 fixed point with 64-bit integer atomics: every fp32 contribution converts
// exactly (|x| < 2^19), integer addition is associative, so the result does not depend on the order in which the faces
// arrive -- bit-identical from run to run (fp32 atomics were not) and rounded to fp32 once.  A non-finite contribution
// marks its vertex, whose normal becomes NaN as it would in floating point.
constexpr float VN_SCALE = 17592186044416.0f;  // 2^44
__global__ void __launch_bounds__(256)
    vertex_normals_kernel(const float* __restrict__ verts, const int* __restrict__ faces, int V, int F,
                          float* __restrict__ normals) {
  extern __shared__ unsigned long long sAcc[];  // [V*3] fixed-point accumulators, then float [V*3] vertices, int [V] flags
  float* sV = reinterpret_cast<float*>(sAcc + V * 3);
  int* sBad = reinterpret_cast<int*>(sV + V * 3);
  const float* v = verts + (size_t)blockIdx.x * V * 3;
  for (int i = threadIdx.x; i < V * 3; i += blockDim.x) {
    sAcc[i] = 0ull;
    sV[i] = v[i];
  }
  for (int i = threadIdx.x; i < V; i += blockDim.x) sBad[i] = 0;
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    const int i0 = faces[3 * f], i1 = faces[3 * f + 1], i2 = faces[3 * f + 2];
    const float a[3] = {sV[3 * i0], sV[3 * i0 + 1], sV[3 * i0 + 2]};
    const float b[3] = {sV[3 * i1], sV[3 * i1 + 1], sV[3 * i1 + 2]};
    const float c[3] = {sV[3 * i2], sV[3 * i2 + 1], sV[3 * i2 + 2]};
    auto cross_add = [&](const float* p, const float* q, const float* o, int dst) {  // cross(p-o, q-o) -> dst
      const float ux = p[0] - o[0], uy = p[1] - o[1], uz = p[2] - o[2];
      const float wx = q[0] - o[0], wy = q[1] - o[1], wz = q[2] - o[2];
      const float n[3] = {uy * wz - uz * wy, uz * wx - ux * wz, ux * wy - uy * wx};
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (fabsf(n[k]) < 524288.f)  // false for NaN / inf as well
          atomicAdd(&sAcc[3 * dst + k], (unsigned long long)__float2ll_rn(n[k] * VN_SCALE));
        else
          atomicOr(&sBad[dst], 1);
      }
    };
    cross_add(c, a, b, i1);
    cross_add(a, b, c, i2);
    cross_add(b, c, a, i0);
  }
  __syncthreads();
  float* out = normals + (size_t)blockIdx.x * V * 3;
  for (int i = threadIdx.x; i < V; i += blockDim.x) {
    const float inv_s = 1.0f / VN_SCALE;
    float x = (float)(long long)sAcc[3 * i] * inv_s, y = (float)(long long)sAcc[3 * i + 1] * inv_s,
          z = (float)(long long)sAcc[3 * i + 2] * inv_s;
    if (sBad[i]) x = y = z = __int_as_float(0x7fc00000);
    const float inv = 1.0f / fmaxf(sqrtf(x * x + y * y + z * z), 1e-6f);
    out[3 * i] = x * inv, out[3 * i + 1] = y * inv, out[3 * i + 2] = z * inv;
  }
}
}  // namespace tamf

extern "C" int tamf_vertex_normals(const float* verts, const int32_t* faces, int N, int V, int F, float* normals,
                                   void* stream) {
  TAMF_REQUIRE(N >= 0 && V > 0 && F >= 0, TAMF_E_BADARG, "tamf_vertex_normals: bad size");
  if (N == 0) return TAMF_OK;
  TAMF_REQUIRE(verts && faces && normals, TAMF_E_BADARG, "tamf_vertex_normals: null pointer");
  TAMF_REQUIRE((size_t)V * 40 <= 48 * 1024, TAMF_E_BADARG, "tamf_vertex_normals: V too large for one CTA");
  int rc = check_device();
  if (rc) return rc;
  vertex_normals_kernel<<<N, 256, (size_t)V * 40, (cudaStream_t)stream>>>(verts, faces, V, F, normals);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}
