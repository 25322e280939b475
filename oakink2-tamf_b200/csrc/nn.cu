// nn.cu -- exact brute-force chamfer nearest neighbour (K=1) and the fused hand->object distance.
//
// Replaces pytorch3d.ops.knn_points(K=1) behind ChamferDistance.forward
// (thirdparty/chamfer_distance/chamfer_distance/chamfer_distance.py:147-162) and
// SegmentRefineModel.multi_object_h2o_dist (src/oakink2_tamf/model/segment_refine_model.py:142-168).
//
// Layout: one CTA scans one (cloud n, candidate split s) pair for ALL queries of that cloud.  Candidates are
// staged through shared memory (coalesced global reads, broadcast LDS.128 + LDS.64 in the inner loop); each thread
// keeps QPT = 4 queries in registers as two packed fp32 pairs, so one candidate load feeds 4 distance evaluations.  Splits are merged with a
// 64-bit atomicMin on (float_bits(d2) << 32 | idx): d2 >= 0 so the bit pattern orders like the value, and the
// index in the low word makes the LOWEST index win exact ties -- the oracle's rule.  The int64 `idx` output
// buffer itself is the packed scratch (memset to all-ones, finalised in place), so no workspace is needed.
//
// Arithmetic (bit-exact contract, oracle/nn_oracle.c): d = (dx*dx + dy*dy) + dz*dz, every operation rounded to fp32
// and none contracted into an FMA.
//
// Inner loop (the kernel is fp32-issue bound: 51 MFLOP per frame and object against 19 KB of traffic): TWO queries per
// instruction on packed fp32 pairs -- 3 FADD2 (differences) + 3 FMUL2 (squares) per candidate and query pair -- then
// 2 scalar FADD per query (ptxas contracts mul.f32x2 + add.f32x2 into FFMA2 even when both carry .rn, so the sums
// stay scalar) and ONE FMNMX per query for a running minimum: 6 instructions per (query, candidate) pair instead of
// 11 with a compare + two selects.  The argmin is recovered afterwards: per group of 16 candidates the thread notes
// whether its running minimum dropped (strict <, so the earliest group wins ties); after each staged chunk the noted
// group is re-evaluated with the same arithmetic and the first candidate equal to the minimum is the index.
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace tamf {

constexpr int NN_QPT = 4;       // queries per thread
constexpr int NN_CHUNK = 1024;  // candidates staged per shared-memory round (16 KB as float4)

struct NNObj {  // one rigid object of a sequence (fused h2o mode)
  int first;    // index of the sequence's first object in the packed cloud array
  int count;    // number of objects of the sequence
};

constexpr int NN_GROUP = 16;    // candidates per argmin-tracking group

typedef unsigned long long nn_f32x2;
__device__ __forceinline__ nn_f32x2 nn_pk2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ nn_f32x2 nn_sub2(nn_f32x2 a, nn_f32x2 b) {
  nn_f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ nn_f32x2 nn_sq2(nn_f32x2 a) {
  nn_f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(r) : "l"(a));
  return r;
}
__device__ __forceinline__ float nn_lo(nn_f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float nn_hi(nn_f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }

// Staged candidates: sxy[j] = (x, x, y, y), sz[j] = (z, z) -- every component duplicated so that one 64-bit operand
// broadcasts it to both queries of a pair.  Slots n_c .. round_up(n_c, NN_GROUP) hold +inf (distance +inf: never a
// minimum).
__device__ __forceinline__ void nn_stage(float4* sxy, float2* sz, int j, float x, float y, float z) {
  sxy[j] = make_float4(x, x, y, y);
  sz[j] = make_float2(z, z);
}
__device__ __forceinline__ void nn_stage_pad(float4* sxy, float2* sz, int n_c) {
  const int n_pad = (n_c + NN_GROUP - 1) / NN_GROUP * NN_GROUP;
  const float inf = __int_as_float(0x7f800000);
  for (int j = n_c + (int)threadIdx.x; j < n_pad; j += blockDim.x) nn_stage(sxy, sz, j, inf, inf, inf);
}

__device__ __forceinline__ void nn_scan_chunk(const float4* __restrict__ sxy, const float2* __restrict__ sz, int n_c,
                                              int base_idx, const float (&qx)[NN_QPT], const float (&qy)[NN_QPT],
                                              const float (&qz)[NN_QPT], float (&best)[NN_QPT], int (&bidx)[NN_QPT]) {
  static_assert(NN_QPT == 4, "two packed query pairs per thread");
  const nn_f32x2 qx2[2] = {nn_pk2(qx[0], qx[1]), nn_pk2(qx[2], qx[3])};
  const nn_f32x2 qy2[2] = {nn_pk2(qy[0], qy[1]), nn_pk2(qy[2], qy[3])};
  const nn_f32x2 qz2[2] = {nn_pk2(qz[0], qz[1]), nn_pk2(qz[2], qz[3])};
  float m[NN_QPT], mprev[NN_QPT];
  int grp[NN_QPT];
#pragma unroll
  for (int q = 0; q < NN_QPT; ++q) m[q] = mprev[q] = best[q], grp[q] = 0;
  const int n_grp = (n_c + NN_GROUP - 1) / NN_GROUP;
  for (int g = 0; g < n_grp; ++g) {
#pragma unroll
    for (int jj = 0; jj < NN_GROUP; ++jj) {
      const int j = g * NN_GROUP + jj;
      const float4 cxy = sxy[j];
      const float2 czz = sz[j];
      const nn_f32x2 cx = nn_pk2(cxy.x, cxy.y), cy = nn_pk2(cxy.z, cxy.w), cz = nn_pk2(czz.x, czz.y);
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        const nn_f32x2 px = nn_sq2(nn_sub2(qx2[pr], cx)), py = nn_sq2(nn_sub2(qy2[pr], cy)), pz = nn_sq2(nn_sub2(qz2[pr], cz));
        const float d0 = __fadd_rn(__fadd_rn(nn_lo(px), nn_lo(py)), nn_lo(pz));
        const float d1 = __fadd_rn(__fadd_rn(nn_hi(px), nn_hi(py)), nn_hi(pz));
        m[2 * pr] = fminf(m[2 * pr], d0);  // a NaN distance never becomes the minimum (like `d < best`)
        m[2 * pr + 1] = fminf(m[2 * pr + 1], d1);
      }
    }
#pragma unroll
    for (int q = 0; q < NN_QPT; ++q) {
      if (m[q] < mprev[q]) {  // strict: the earliest group holding the minimum is remembered
        mprev[q] = m[q];
        grp[q] = g;
      }
    }
  }
  // argmin: first candidate of the remembered group whose distance equals the new minimum (same operations, same
  // rounding as above -> bit-identical distances); ascending scan keeps the lowest index on ties
#pragma unroll
  for (int q = 0; q < NN_QPT; ++q) {
    if (m[q] < best[q]) {
      best[q] = m[q];
      int found = -1;
      for (int jj = NN_GROUP - 1; jj >= 0; --jj) {
        const int j = grp[q] * NN_GROUP + jj;
        const float4 cxy = sxy[j];
        const float2 czz = sz[j];
        const float dx = __fsub_rn(qx[q], cxy.x), dy = __fsub_rn(qy[q], cxy.z), dz = __fsub_rn(qz[q], czz.x);
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d == m[q]) found = j;
      }
      bidx[q] = base_idx + found;
    }
  }
}

__device__ __forceinline__ void nn_publish(unsigned long long* packed, float best, int bidx) {
  if (bidx >= 0) {
    unsigned long long v = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned int)bidx;
    atomicMin(packed, v);
  }
}

// grid (N, splits, query blocks of <= 4096); block = ceil(min(P1, 4096) / QPT) rounded up to a warp
__global__ void __launch_bounds__(1024) nn_scan_kernel(const float* __restrict__ x, const float* __restrict__ y, int P1,
                                                       int P2, int per_split, unsigned long long* __restrict__ packed) {
  __shared__ float4 sxy[NN_CHUNK];
  __shared__ float2 sz[NN_CHUNK];
  const int n = blockIdx.x;
  const int c_begin = blockIdx.y * per_split;
  const int c_end = min(P2, c_begin + per_split);
  const float* xn = x + (size_t)n * P1 * 3;
  const float* yn = y + (size_t)n * P2 * 3;
  const int q0 = blockIdx.z * (blockDim.x * NN_QPT);  // query block (P1 > 4096: the object cloud as the query set)

  float qx[NN_QPT], qy[NN_QPT], qz[NN_QPT], best[NN_QPT];
  int bidx[NN_QPT];
#pragma unroll
  for (int q = 0; q < NN_QPT; ++q) {
    const int i = q0 + threadIdx.x + q * blockDim.x;
    const bool ok = i < P1;
    qx[q] = ok ? xn[3 * i + 0] : 0.f;
    qy[q] = ok ? xn[3 * i + 1] : 0.f;
    qz[q] = ok ? xn[3 * i + 2] : 0.f;
    best[q] = __int_as_float(0x7f800000);  // +inf
    bidx[q] = -1;
  }
  for (int c0 = c_begin; c0 < c_end; c0 += NN_CHUNK) {
    const int n_c = min(NN_CHUNK, c_end - c0);
    __syncthreads();
    for (int j = threadIdx.x; j < n_c; j += blockDim.x) {
      const float* p = yn + (size_t)(c0 + j) * 3;
      nn_stage(sxy, sz, j, p[0], p[1], p[2]);
    }
    nn_stage_pad(sxy, sz, n_c);
    __syncthreads();
    nn_scan_chunk(sxy, sz, n_c, c0, qx, qy, qz, best, bidx);
  }
#pragma unroll
  for (int q = 0; q < NN_QPT; ++q) {
    const int i = q0 + threadIdx.x + q * blockDim.x;
    if (i < P1) nn_publish(packed + (size_t)n * P1 + i, best[q], bidx[q]);
  }
}

// rot6d -> rotation matrix rows (src/dev_fn/transform/rotation.py:446-467; F.normalize eps 1e-12)
__device__ __forceinline__ void rot6d_rows(const float* d6, float (&R)[9]) {
  float a1x = d6[0], a1y = d6[1], a1z = d6[2], a2x = d6[3], a2y = d6[4], a2z = d6[5];
  float n1 = fmaxf(sqrtf(a1x * a1x + a1y * a1y + a1z * a1z), 1e-12f);
  float b1x = a1x / n1, b1y = a1y / n1, b1z = a1z / n1;
  float dt = b1x * a2x + b1y * a2y + b1z * a2z;
  float b2x = a2x - dt * b1x, b2y = a2y - dt * b1y, b2z = a2z - dt * b1z;
  float n2 = fmaxf(sqrtf(b2x * b2x + b2y * b2y + b2z * b2z), 1e-12f);
  b2x /= n2, b2y /= n2, b2z /= n2;
  R[0] = b1x, R[1] = b1y, R[2] = b1z;
  R[3] = b2x, R[4] = b2y, R[5] = b2z;
  R[6] = b1y * b2z - b1z * b2y;
  R[7] = b1z * b2x - b1x * b2z;
  R[8] = b1x * b2y - b1y * b2x;
}

// Fused variant: grid (B*T, splits).  Candidate j of the sequence's concatenated cloud is object o = j / P,
// point p = j % P, moved to the world by frame t's transform of object o while it is staged.
__global__ void __launch_bounds__(1024)
    h2o_scan_kernel(const float* __restrict__ verts, const float* __restrict__ obj_traj,
                    const float* __restrict__ obj_points, const int* __restrict__ obj_first, int T, int V, int nobj_max,
                    int P, int per_split, unsigned long long* __restrict__ packed) {
  __shared__ float4 sxy[NN_CHUNK];
  __shared__ float2 sz[NN_CHUNK];
  __shared__ float sR[12];
  const int f = blockIdx.x, b = f / T, t = f % T;
  const int first = obj_first[b], nobj = obj_first[b + 1] - first;
  const int P2 = nobj * P;
  const int c_begin = blockIdx.y * per_split;
  const int c_end = min(P2, c_begin + per_split);
  const float* xn = verts + (size_t)f * V * 3;

  float qx[NN_QPT], qy[NN_QPT], qz[NN_QPT], best[NN_QPT];
  int bidx[NN_QPT];
#pragma unroll
  for (int q = 0; q < NN_QPT; ++q) {
    const int i = threadIdx.x + q * blockDim.x;
    const bool ok = i < V;
    qx[q] = ok ? xn[3 * i + 0] : 0.f;
    qy[q] = ok ? xn[3 * i + 1] : 0.f;
    qz[q] = ok ? xn[3 * i + 2] : 0.f;
    best[q] = __int_as_float(0x7f800000);
    bidx[q] = -1;
  }
  int cur_obj = -1;
  for (int c0 = c_begin; c0 < c_end;) {
    const int o = c0 / P;
    const int n_c = min(min(NN_CHUNK, c_end - c0), (o + 1) * P - c0);  // a chunk never straddles objects
    __syncthreads();
    if (o != cur_obj) {
      if (threadIdx.x == 0) {
        const float* tr = obj_traj + (((size_t)b * nobj_max + o) * T + t) * 9;
        float R[9];
        rot6d_rows(tr + 3, R);
#pragma unroll
        for (int k = 0; k < 9; ++k) sR[k] = R[k];
        sR[9] = tr[0], sR[10] = tr[1], sR[11] = tr[2];
      }
      cur_obj = o;
      __syncthreads();
    }
    const float* pts = obj_points + ((size_t)(first + o) * P + (c0 - o * P)) * 3;
    for (int j = threadIdx.x; j < n_c; j += blockDim.x) {
      const float px = pts[3 * j], py = pts[3 * j + 1], pz = pts[3 * j + 2];
      // transf_point_array: R p + t   (src/dev_fn/transform/transform.py:36-53)
      nn_stage(sxy, sz, j, fmaf(sR[2], pz, fmaf(sR[1], py, sR[0] * px)) + sR[9],
               fmaf(sR[5], pz, fmaf(sR[4], py, sR[3] * px)) + sR[10],
               fmaf(sR[8], pz, fmaf(sR[7], py, sR[6] * px)) + sR[11]);
    }
    nn_stage_pad(sxy, sz, n_c);
    __syncthreads();
    nn_scan_chunk(sxy, sz, n_c, c0, qx, qy, qz, best, bidx);
    c0 += n_c;
  }
#pragma unroll
  for (int q = 0; q < NN_QPT; ++q) {
    const int i = threadIdx.x + q * blockDim.x;
    if (i < V) nn_publish(packed + (size_t)f * V + i, best[q], bidx[q]);
  }
}


// ------------------------------------------------------------------------------------------------
// Exact search with block pruning (fused h2o mode, P <= 8192 points per object).
//
// The brute-force scan above spends 6 instructions on each of the 778 x 8192 (query, candidate) pairs of a frame and
// object.  A rigid transform preserves distances, so blocks of candidates can be rejected in the OBJECT frame:
//   build (once per batch and object):  k-d order the canonical cloud (recursive median splits of the widest axis)
//                                      into blocks of 64 consecutive points, keep each block's bounding box;
//   query (CTA per frame and object):  stage the sorted cloud moved to the world exactly like the scan kernel does;
//     one WARP per hand vertex: q' = R^T (v - t); lower bound LB of |q' - p| to every block box (lanes over blocks);
//     evaluate the block with the smallest LB -> first upper bound `best`; then only blocks with
//         LB <= best (1 + 1e-3) + margin
//     are evaluated (lanes over candidates).  Evaluated candidates use the WORLD points and the scan kernel's
//     arithmetic, so the distances -- and, with the (d2, lowest index) rule, the indices -- are bit-identical to the
//     exhaustive scan.  A rejected block holds only points strictly farther than the final minimum: the margin
//     (1e-3 relative + 2e-5 of the coordinate magnitudes) is >= 10x the fp32 error of the transform, of q' and of a
//     rot6d rotation's deviation from orthonormality; if R R^T differs from I by more than 1e-4 (degenerate rot6d)
//     nothing is rejected.
// ------------------------------------------------------------------------------------------------
constexpr int NNP_BS = 64;          // candidates per block
constexpr int NNP_MAXP = 8192;      // points per object handled by the pruned path
constexpr int NNP_QWARPS = 32;      // query warps per CTA
constexpr unsigned NNP_PAD = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned nnp_ord(float f) {  // order-preserving float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float nnp_unord(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// Index build, one CTA per object: k-d ordering of the canonical cloud.  The points are split recursively at the
// median of the widest axis of their bounding box (log2(np2 / 64) levels) until runs of 64 consecutive points remain;
// every level is one bitonic sort of the composite keys (segment | coordinate | index) restricted to the segment
// length.  Balanced runs with tight boxes: what the query kernel's block bounds prune on.
// grid = objects; block 1024.  sorted [obj][Ppad] float4 (x, y, z, original index bits; pads = +inf / NNP_PAD),
// boxes [obj][2 * (nblk + 1)] float4: lo/hi of each block, then lo/hi of the whole cloud.
__global__ void __launch_bounds__(1024) nnp_build_kernel(const float* __restrict__ pts, int P, int Ppad, int np2,
                                                         float4* __restrict__ sorted, float4* __restrict__ boxes) {
  extern __shared__ unsigned long long keys[];  // np2
  __shared__ unsigned segbox[6][NNP_MAXP / NNP_BS];  // per segment: ord(lo xyz), ord(hi xyz)
  const int o = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp = (int)(blockDim.x >> 5);
  const float* po = pts + (size_t)o * P * 3;
  const float inf = __int_as_float(0x7f800000);
  for (int i = tid; i < np2; i += blockDim.x) keys[i] = i < P ? (unsigned long long)i : ~0ull;
  __syncthreads();
  for (int seglen = np2; seglen > NNP_BS; seglen >>= 1) {
    const int nseg = np2 / seglen;
    for (int sgi = tid; sgi < nseg; sgi += blockDim.x)
#pragma unroll
      for (int a = 0; a < 3; ++a) segbox[a][sgi] = 0xFFFFFFFFu, segbox[3 + a][sgi] = 0u;
    __syncthreads();
    // segment boxes: a warp's 32 consecutive positions share a segment (seglen >= 128)
    for (int base = warp * 32; base < P; base += nwarp * 32) {
      const int i = base + lane;
      unsigned lo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, hi[3] = {0u, 0u, 0u};
      if (i < P) {
        const unsigned id = (unsigned)(keys[i] & 0x1FFFu);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float c = po[3 * id + a];
          if (c == c) lo[a] = hi[a] = nnp_ord(c);  // NaN coordinates do not shape the boxes
        }
      }
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        lo[a] = __reduce_min_sync(0xffffffffu, lo[a]);
        hi[a] = __reduce_max_sync(0xffffffffu, hi[a]);
      }
      if (lane == 0) {
        const int sgi = base / seglen;
#pragma unroll
        for (int a = 0; a < 3; ++a) atomicMin(&segbox[a][sgi], lo[a]), atomicMax(&segbox[3 + a][sgi], hi[a]);
      }
    }
    __syncthreads();
    for (int i = tid; i < P; i += blockDim.x) {
      const int sgi = i / seglen;
      const unsigned id = (unsigned)(keys[i] & 0x1FFFu);
      float ext[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) ext[a] = nnp_unord(segbox[3 + a][sgi]) - nnp_unord(segbox[a][sgi]);
      const int ax = (ext[0] >= ext[1] && ext[0] >= ext[2]) ? 0 : (ext[1] >= ext[2] ? 1 : 2);
      keys[i] = ((unsigned long long)sgi << 45) | ((unsigned long long)nnp_ord(po[3 * id + ax]) << 13) | id;
    }
    __syncthreads();
    for (int k = 2; k <= seglen; k <<= 1)  // keys of different segments never meet: the network stops at seglen
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < np2; i += blockDim.x) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = keys[i], b = keys[ixj];
            const bool up = (i & k) == 0 || k == seglen;  // the last merge of a segment sorts it ascending
            if ((a > b) == up) keys[i] = b, keys[ixj] = a;
          }
        }
        __syncthreads();
      }
  }
  float4* so = sorted + (size_t)o * Ppad;
  for (int j = tid; j < Ppad; j += blockDim.x) {
    float4 v = make_float4(inf, inf, inf, __uint_as_float(NNP_PAD));
    if (j < P) {
      const unsigned i = (unsigned)(keys[j] & 0x1FFFu);
      v = make_float4(po[3 * i], po[3 * i + 1], po[3 * i + 2], __uint_as_float(i));
    }
    so[j] = v;
  }
  const int nblk = Ppad / NNP_BS;
  float4* bo = boxes + (size_t)o * 2 * (nblk + 1);
  float gl[3] = {inf, inf, inf}, gh[3] = {-inf, -inf, -inf};
  for (int b = warp; b < nblk; b += nwarp) {
    float l[3] = {inf, inf, inf}, h[3] = {-inf, -inf, -inf};
#pragma unroll
    for (int e = 0; e < NNP_BS / 32; ++e) {
      const int j = b * NNP_BS + e * 32 + lane;
      if (j < P) {
        const unsigned i = (unsigned)(keys[j] & 0x1FFFu);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          const float c = po[3 * i + a];
          l[a] = fminf(l[a], c), h[a] = fmaxf(h[a], c);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int sft = 16; sft; sft >>= 1) {
        l[a] = fminf(l[a], __shfl_xor_sync(0xffffffffu, l[a], sft));
        h[a] = fmaxf(h[a], __shfl_xor_sync(0xffffffffu, h[a], sft));
      }
      gl[a] = fminf(gl[a], l[a]), gh[a] = fmaxf(gh[a], h[a]);
    }
    if (lane == 0) bo[2 * b] = make_float4(l[0], l[1], l[2], 0.f), bo[2 * b + 1] = make_float4(h[0], h[1], h[2], 0.f);
  }
  // whole-cloud box (only its magnitude is used, for the rejection margin)
  __syncthreads();
  float* red = reinterpret_cast<float*>(&segbox[0][0]);  // 32 warps x 6 floats; the segment boxes are dead
  if (lane == 0)
#pragma unroll
    for (int a = 0; a < 3; ++a) red[warp * 6 + a] = gl[a], red[warp * 6 + 3 + a] = gh[a];
  __syncthreads();
  if (tid == 0) {
    float l[3] = {inf, inf, inf}, h[3] = {-inf, -inf, -inf};
    for (int w = 0; w < nwarp; ++w)
#pragma unroll
      for (int a = 0; a < 3; ++a) l[a] = fminf(l[a], red[w * 6 + a]), h[a] = fmaxf(h[a], red[w * 6 + 3 + a]);
    bo[2 * nblk] = make_float4(l[0], l[1], l[2], 0.f);
    bo[2 * nblk + 1] = make_float4(h[0], h[1], h[2], 0.f);
  }
}

__device__ unsigned long long nnp_stats[2];  // debug (TAMF_NN_STATS=1): blocks evaluated, queries

__device__ __forceinline__ float nnp_d2(float qx, float qy, float qz, const float4& c) {  // the scan kernel's arithmetic
  const float dx = __fsub_rn(qx, c.x), dy = __fsub_rn(qy, c.y), dz = __fsub_rn(qz, c.z);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// grid (B*T, max objects per sequence); block (NNP_QWARPS * 32).  Dynamic shared memory: world points float4[Ppad],
// block boxes float4[2 * nblk].
__global__ void __launch_bounds__(NNP_QWARPS * 32, 1)
    h2o_pruned_kernel(const float* __restrict__ verts, const float* __restrict__ obj_traj,
                      const float4* __restrict__ sorted, const float4* __restrict__ boxes,
                      const int* __restrict__ obj_first, int T, int V, int nobj_max, int P, int Ppad,
                      unsigned long long* __restrict__ packed, int stats) {
  extern __shared__ float4 nnp_smem[];
  __shared__ float sR[12];
  __shared__ float4 ssup[2 * (NNP_MAXP / NNP_BS / 4)];  // super-block boxes (lo, hi)
  __shared__ float s_boxmax;
  __shared__ int s_prune;
  const int f = blockIdx.x, b = f / T, t = f % T, o = blockIdx.y;
  const int first = obj_first[b], nobj = obj_first[b + 1] - first;
  if (o >= nobj) return;
  const int nblk = Ppad / NNP_BS;
  float4* wp = nnp_smem;
  float4* sbox = nnp_smem + Ppad;
  const float4* so = sorted + (size_t)(first + o) * Ppad;
  const float4* bo = boxes + (size_t)(first + o) * 2 * (nblk + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float inf = __int_as_float(0x7f800000);
  if (tid == 0) {
    const float* tr = obj_traj + (((size_t)b * nobj_max + o) * T + t) * 9;
    float R[9];
    rot6d_rows(tr + 3, R);
#pragma unroll
    for (int k = 0; k < 9; ++k) sR[k] = R[k];
    sR[9] = tr[0], sR[10] = tr[1], sR[11] = tr[2];
    float dev = 0.f;  // || R R^T - I ||_max
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = i; j < 3; ++j) {
        const float g = R[3 * i] * R[3 * j] + R[3 * i + 1] * R[3 * j + 1] + R[3 * i + 2] * R[3 * j + 2];
        dev = fmaxf(dev, fabsf(g - (i == j ? 1.f : 0.f)));
      }
    s_prune = (dev <= 1e-4f) ? 1 : 0;  // NaN -> 0
    const float4 l = bo[2 * nblk], h = bo[2 * nblk + 1];
    s_boxmax = fmaxf(fmaxf(fmaxf(fabsf(l.x), fabsf(l.y)), fabsf(l.z)), fmaxf(fmaxf(fabsf(h.x), fabsf(h.y)), fabsf(h.z)));
  }
  __syncthreads();
  for (int j = tid; j < Ppad; j += blockDim.x) {
    float4 c = so[j];
    if (__float_as_uint(c.w) != NNP_PAD) {
      // transf_point_array: R p + t  (src/dev_fn/transform/transform.py:36-53) -- same operations as h2o_scan_kernel
      const float wx = fmaf(sR[2], c.z, fmaf(sR[1], c.y, sR[0] * c.x)) + sR[9];
      const float wy = fmaf(sR[5], c.z, fmaf(sR[4], c.y, sR[3] * c.x)) + sR[10];
      const float wz = fmaf(sR[8], c.z, fmaf(sR[7], c.y, sR[6] * c.x)) + sR[11];
      c.x = wx, c.y = wy, c.z = wz;
    }
    wp[j] = c;
  }
  for (int j = tid; j < 2 * nblk; j += blockDim.x) sbox[j] = bo[j];
  __syncthreads();
  if (tid < (nblk + 3) / 4) {
    // super boxes: union of 4 consecutive block boxes.  fminf / fmaxf drop a NaN operand, and a block whose box has a NaN
    // (or infinite) corner must never be rejected: `poison` turns the whole super box into NaN then (NaN bound: kept).
    float4 l = sbox[8 * tid], h = sbox[8 * tid + 1];
    float poison = ((l.x - l.x) + (l.y - l.y) + (l.z - l.z)) + ((h.x - h.x) + (h.y - h.y) + (h.z - h.z));
    for (int j = 1; j < 4 && 4 * tid + j < nblk; ++j) {
      const float4 l2 = sbox[2 * (4 * tid + j)], h2 = sbox[2 * (4 * tid + j) + 1];
      poison += ((l2.x - l2.x) + (l2.y - l2.y) + (l2.z - l2.z)) + ((h2.x - h2.x) + (h2.y - h2.y) + (h2.z - h2.z));
      l.x = fminf(l.x, l2.x), l.y = fminf(l.y, l2.y), l.z = fminf(l.z, l2.z);
      h.x = fmaxf(h.x, h2.x), h.y = fmaxf(h.y, h2.y), h.z = fmaxf(h.z, h2.z);
    }
    l.x += poison;
    ssup[2 * tid] = l, ssup[2 * tid + 1] = h;
  }
  __syncthreads();

  const bool prune = s_prune != 0;
  const float tmax = fmaxf(fmaxf(fabsf(sR[9]), fabsf(sR[10])), fabsf(sR[11]));
  const float* xn = verts + (size_t)f * V * 3;
  // Two levels of boxes: a SUPER block = 4 consecutive blocks of the Morton order (256 points), its box the union of
  // theirs -- at most 32 of them, one per lane.  A query computes the 32 super bounds once and the 4 block bounds only of
  // the supers it cannot reject (the 128 block bounds of the one-level form were half of the per-vertex work).  A bound
  // of a union is <= the bounds of its parts, so rejecting a super under the same rule rejects only blocks the rule
  // would have rejected: the result stays bit-identical to the exhaustive scan.
  const int nsup = (nblk + 3) >> 2;
  auto box_lb = [&](const float4& l, const float4& h, float px, float py, float pz) {
    const float dx = fmaxf(fmaxf(l.x - px, px - h.x), 0.f);
    const float dy = fmaxf(fmaxf(l.y - py, py - h.y), 0.f);
    const float dz = fmaxf(fmaxf(l.z - pz, pz - h.z), 0.f);
    return dx * dx + dy * dy + dz * dz;
  };
  for (int q = warp; q < V; q += NNP_QWARPS) {
    const float vx = xn[3 * q], vy = xn[3 * q + 1], vz = xn[3 * q + 2];
    // the query in the object frame
    const float ux = vx - sR[9], uy = vy - sR[10], uz = vz - sR[11];
    const float px = sR[0] * ux + sR[3] * uy + sR[6] * uz;
    const float py = sR[1] * ux + sR[4] * uy + sR[7] * uz;
    const float pz = sR[2] * ux + sR[5] * uy + sR[8] * uz;
    const float mag = fmaxf(fmaxf(fmaxf(fabsf(vx), fabsf(vy)), fabsf(vz)),
                            fmaxf(fmaxf(fmaxf(fabsf(px), fabsf(py)), fabsf(pz)), fmaxf(tmax, s_boxmax)));
    const float absm = 2e-5f * mag + 1e-30f;
    // bounds are sums of squares (>= +0 or NaN): their bit patterns order like the values (NaN patterns above +inf), so
    // one REDUX finds a minimum; any block is a valid start
    const float slb = lane < nsup ? box_lb(ssup[2 * lane], ssup[2 * lane + 1], px, py, pz) : inf;
    int s0;
    {
      const unsigned mine = lane < nsup ? __float_as_uint(slb) : 0xFFFFFFFFu;
      const unsigned wmin = __reduce_min_sync(0xffffffffu, mine);
      s0 = __ffs(__ballot_sync(0xffffffffu, mine == wmin)) - 1;
    }
    // the four block bounds of a super: lane j & 3 takes block 4 s + (j & 3) (all lanes compute, lanes 0..3 count)
    auto block_lbs = [&](int s) {
      const int blk = 4 * s + (lane & 3);
      return blk < nblk ? box_lb(sbox[2 * blk], sbox[2 * blk + 1], px, py, pz) : inf;
    };
    float bd = inf;
    unsigned bi = NNP_PAD;
    auto eval_block = [&](int blk) {
#pragma unroll
      for (int e = 0; e < NNP_BS / 32; ++e) {
        const float4 c = wp[blk * NNP_BS + e * 32 + lane];
        const float d = nnp_d2(vx, vy, vz, c);
        const unsigned ci = __float_as_uint(c.w);
        if (d < bd || (d == bd && bi != NNP_PAD && ci < bi)) bd = d, bi = ci;
      }
    };
    auto warp_min = [&]() {  // bd >= +0 or +inf, never NaN: unsigned order of the bits == order of the values
      return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(bd)));
    };
    auto bound_now = [&]() {
      if (!prune) return inf;
      const float r = sqrtf(warp_min()) * 1.001f + absm;
      return r * r;
    };
    // first block: the best block of the best super
    int fb;
    {
      const float b0 = block_lbs(s0);
      const bool have = (lane & 3) == lane && 4 * s0 + lane < nblk;  // lanes 0..3 with an existing block
      const unsigned mine = have ? __float_as_uint(b0) : 0xFFFFFFFFu;
      const unsigned wmin = __reduce_min_sync(0xffffffffu, mine);
      fb = 4 * s0 + (__ffs(__ballot_sync(0xffffffffu, have && mine == wmin)) - 1);
    }
    eval_block(fb);
    int n_eval = 1;
    float bound2 = bound_now();
    // every super that cannot be rejected, the first one included (its other blocks), in index order
    unsigned sm = __ballot_sync(0xffffffffu, lane < nsup && !(slb > bound2));
    while (sm) {
      const int s = __ffs(sm) - 1;
      sm &= sm - 1;
      if (__shfl_sync(0xffffffffu, slb, s) > bound2) continue;  // (the bound has tightened since the ballot)
      const float bl = block_lbs(s);
      const int blk_l = 4 * s + (lane & 3);
      unsigned bm = __ballot_sync(0xffffffffu, lane < 4 && blk_l < nblk && blk_l != fb && !(bl > bound2));
      const bool any = bm != 0;
      while (bm) {
        const int j = __ffs(bm) - 1;
        bm &= bm - 1;
        eval_block(4 * s + j);
        ++n_eval;
      }
      if (any && sm) bound2 = bound_now();  // tighten for the remaining supers
    }
    // lexicographic (d2, index) minimum over the lanes: smallest distance first, lowest index among its holders
    {
      const unsigned dmin = __reduce_min_sync(0xffffffffu, bi != NNP_PAD ? __float_as_uint(bd) : 0xFFFFFFFFu);
      const unsigned imin = __reduce_min_sync(
          0xffffffffu, (bi != NNP_PAD && __float_as_uint(bd) == dmin) ? bi : NNP_PAD);
      bd = __uint_as_float(dmin), bi = imin;
    }
    if (lane == 0 && bi != NNP_PAD) nn_publish(packed + (size_t)f * V + q, bd, o * P + (int)bi);
    if (stats && lane == 0) {
      atomicAdd(&nnp_stats[0], (unsigned long long)n_eval);
      atomicAdd(&nnp_stats[1], 1ull);
    }
  }
}

// packed -> (d2 | dist, idx) in place
__global__ void nn_finalize_kernel(unsigned long long* __restrict__ packed, float* __restrict__ d_out, size_t n,
                                   int take_sqrt) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long v = packed[i];
  float d = __uint_as_float((unsigned int)(v >> 32));
  d_out[i] = take_sqrt ? sqrtf(d) : d;
  reinterpret_cast<long long*>(packed)[i] = (long long)(unsigned int)(v & 0xffffffffull);
}

// Grow-only device scratch keyed by (device, stream, slot): the one-shot entry points stage small host tables / the
// search index here.  Growth frees the old block (cudaFree waits for the device), steady state allocates nothing.
static int nn_scratch(cudaStream_t stream, int slot, size_t bytes, void** out) {
  struct Buf {
    void* p = nullptr;
    size_t cap = 0;
  };
  static std::mutex mu;
  static std::map<std::tuple<int, cudaStream_t, int>, Buf> bufs;
  int dev = 0;
  TAMF_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  Buf& b = bufs[std::make_tuple(dev, stream, slot)];
  if (b.cap < bytes) {
    if (b.p) {
      TAMF_CUDA_CHECK(cudaDeviceSynchronize());
      cudaFree(b.p);
      b.p = nullptr, b.cap = 0;
    }
    TAMF_CUDA_CHECK(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
  }
  *out = b.p;
  return TAMF_OK;
}

static int pick_splits(int n_clouds, int P2) {
  // enough CTAs for >= 2 waves of 148 SMs, but never split below one shared-memory chunk
  int want = (2 * 148 + n_clouds - 1) / n_clouds;
  int max_splits = (P2 + NN_CHUNK - 1) / NN_CHUNK;
  int s = want < 1 ? 1 : want;
  if (s > max_splits) s = max_splits;
  if (s < 1) s = 1;
  return s;
}

}  // namespace tamf

using namespace tamf;

extern "C" int tamf_nn_query(const float* x, const float* y, int N, int P1, int P2, float* d2, int64_t* idx,
                             void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TAMF_REQUIRE(N >= 0 && P1 >= 0, TAMF_E_BADARG, "tamf_nn_query: negative size");
  if (N == 0 || P1 == 0) return TAMF_OK;  // empty query set: nothing to write
  TAMF_REQUIRE(P2 > 0, TAMF_E_BADARG, "tamf_nn_query: empty candidate cloud (P2 == 0) has no nearest neighbour");
  TAMF_REQUIRE(x && y && d2 && idx, TAMF_E_BADARG, "tamf_nn_query: null pointer");
  TAMF_REQUIRE((long long)P2 < 2147483647LL && (long long)P1 <= 65535LL * 1024 * NN_QPT, TAMF_E_BADARG,
               "tamf_nn_query: size overflow");
  TAMF_REQUIRE(aligned16(idx), TAMF_E_ALIGN, "tamf_nn_query: idx must be 16-byte aligned");
  int rc = check_device();
  if (rc) return rc;
  const size_t total = (size_t)N * P1;
  TAMF_CUDA_CHECK(cudaMemsetAsync(idx, 0xFF, total * sizeof(int64_t), stream));
  // queries are tiled in blocks of <= 4096 over gridDim.z: the reverse pass of ChamferDistance (chamfer_distance.py:148)
  // queries with the nobj * 8192 object points
  const int qblocks = (P1 + 1024 * NN_QPT - 1) / (1024 * NN_QPT);
  const int q_per_block = (P1 + qblocks - 1) / qblocks;
  int threads = ((q_per_block + NN_QPT - 1) / NN_QPT + 31) / 32 * 32;
  int splits = pick_splits(N * qblocks, P2);
  int per_split = ((P2 + splits - 1) / splits + 3) / 4 * 4;
  splits = (P2 + per_split - 1) / per_split;
  // gridDim.x limit is 2^31-1; N beyond 65535 splits is fine on x
  nn_scan_kernel<<<dim3(N, splits, qblocks), threads, 0, stream>>>(x, y, P1, P2, per_split, (unsigned long long*)idx);
  TAMF_LAUNCH_CHECK();
  nn_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((unsigned long long*)idx, d2, total, 0);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

static int nnp_configure() {
  static bool configured = false;
  if (!configured) {
    TAMF_CUDA_CHECK(cudaFuncSetAttribute(nnp_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NNP_MAXP * 8));
    TAMF_CUDA_CHECK(cudaFuncSetAttribute(h2o_pruned_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (NNP_MAXP + 2 * (NNP_MAXP / NNP_BS)) * (int)sizeof(float4)));
    configured = true;
  }
  return TAMF_OK;
}
static size_t nnp_index_bytes(int total_obj, int P) {
  if (P > NNP_MAXP || P <= 0 || total_obj <= 0) return 0;
  const int Ppad = (P + NNP_BS - 1) / NNP_BS * NNP_BS, nblk = Ppad / NNP_BS;
  return ((size_t)total_obj * Ppad + (size_t)total_obj * 2 * (nblk + 1)) * sizeof(float4);
}
static int nnp_build(const float* obj_points, int total_obj, int P, void* index, cudaStream_t stream) {
  int rc = nnp_configure();
  if (rc) return rc;
  const int Ppad = (P + NNP_BS - 1) / NNP_BS * NNP_BS;
  int np2 = 64;
  while (np2 < P) np2 <<= 1;
  float4* d_sorted = (float4*)index;
  float4* d_boxes = d_sorted + (size_t)total_obj * Ppad;
  nnp_build_kernel<<<total_obj, 1024, (size_t)np2 * 8, stream>>>(obj_points, P, Ppad, np2, d_sorted, d_boxes);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

// `index`: block-pruned search over a built index (obj_points unused); else the exhaustive scan of obj_points.
static int h2o_dist_impl(const float* verts, const float* obj_traj, const float* obj_points, const void* index,
                         const int32_t* obj_first_host, int B, int T, int V, int nobj_max, int P, float* dist,
                         int64_t* idx, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  TAMF_REQUIRE(B > 0 && T > 0 && V > 0 && P > 0 && nobj_max > 0, TAMF_E_BADARG, "tamf_h2o_dist: bad size");
  TAMF_REQUIRE(verts && obj_traj && (obj_points || index) && obj_first_host && dist && idx, TAMF_E_BADARG,
               "tamf_h2o_dist: null pointer");
  TAMF_REQUIRE(V <= 1024 * NN_QPT, TAMF_E_BADARG, "tamf_h2o_dist: V > 4096 unsupported");
  TAMF_REQUIRE(B <= 4096, TAMF_E_BADARG, "tamf_h2o_dist: B > 4096 unsupported");
  int rc = check_device();
  if (rc) return rc;
  int max_nobj = 0;
  for (int b = 0; b < B; ++b) {
    int n = obj_first_host[b + 1] - obj_first_host[b];
    TAMF_REQUIRE(n >= 1 && n <= nobj_max, TAMF_E_BADARG,
                 "tamf_h2o_dist: every sequence needs 1..nobj_max objects (empty cloud has no nearest neighbour)");
    if (n > max_nobj) max_nobj = n;
  }
  // obj_first goes to the device through a small async copy into a grow-only buffer owned by this (device, stream):
  // calls on one stream are ordered, calls on different streams or devices never share the buffer
  int* d_first = nullptr;
  if ((rc = nn_scratch(stream, 0, sizeof(int) * (size_t)(B + 1), (void**)&d_first))) return rc;
  TAMF_CUDA_CHECK(cudaMemcpyAsync(d_first, obj_first_host, sizeof(int) * (size_t)(B + 1), cudaMemcpyHostToDevice, stream));
  const size_t total = (size_t)B * T * V;
  TAMF_CUDA_CHECK(cudaMemsetAsync(idx, 0xFF, total * sizeof(int64_t), stream));
  static const bool env_stats = getenv("TAMF_NN_STATS") && getenv("TAMF_NN_STATS")[0] == '1';
  if (index) {
    TAMF_REQUIRE(P <= NNP_MAXP, TAMF_E_BADARG, "tamf_h2o_dist_indexed: an index exists only for P <= 8192");
    if ((rc = nnp_configure())) return rc;
    const int total_obj = obj_first_host[B];
    const int Ppad = (P + NNP_BS - 1) / NNP_BS * NNP_BS, nblk = Ppad / NNP_BS;
    const float4* d_sorted = (const float4*)index;
    const float4* d_boxes = d_sorted + (size_t)total_obj * Ppad;
    h2o_pruned_kernel<<<dim3(B * T, max_nobj), NNP_QWARPS * 32, (size_t)(Ppad + 2 * nblk) * sizeof(float4), stream>>>(
        verts, obj_traj, d_sorted, d_boxes, d_first, T, V, nobj_max, P, Ppad, (unsigned long long*)idx, env_stats ? 1 : 0);
    TAMF_LAUNCH_CHECK();
    if (env_stats) {
      unsigned long long h[2] = {0, 0};
      cudaStreamSynchronize(stream);
      cudaMemcpyFromSymbol(h, nnp_stats, sizeof(h));
      fprintf(stderr, "[tamf nn stats] blocks evaluated per query: %.2f of %d (cumulative over %llu queries)\n",
              h[1] ? (double)h[0] / (double)h[1] : 0.0, nblk, h[1]);
    }
  } else {
    int threads = ((V + NN_QPT - 1) / NN_QPT + 31) / 32 * 32;
    int P2 = max_nobj * P;
    int splits = pick_splits(B * T, P2);
    int per_split = ((P2 + splits - 1) / splits + 3) / 4 * 4;
    splits = (P2 + per_split - 1) / per_split;
    h2o_scan_kernel<<<dim3(B * T, splits), threads, 0, stream>>>(verts, obj_traj, obj_points, d_first, T, V, nobj_max, P,
                                                                 per_split, (unsigned long long*)idx);
    TAMF_LAUNCH_CHECK();
  }
  nn_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>((unsigned long long*)idx, dist, total, 1);
  TAMF_LAUNCH_CHECK();
  return TAMF_OK;
}

extern "C" size_t tamf_h2o_index_bytes(int total_obj, int P) { return nnp_index_bytes(total_obj, P); }

extern "C" int tamf_h2o_index_build(const float* obj_points, int total_obj, int P, void* index, size_t index_bytes,
                                    void* stream) {
  TAMF_REQUIRE(obj_points && index && total_obj > 0 && P > 0, TAMF_E_BADARG, "tamf_h2o_index_build: bad argument");
  TAMF_REQUIRE(P <= NNP_MAXP, TAMF_E_BADARG, "tamf_h2o_index_build: P > 8192 has no index (use tamf_h2o_dist)");
  TAMF_REQUIRE(index_bytes >= nnp_index_bytes(total_obj, P), TAMF_E_BADARG, "tamf_h2o_index_build: index buffer too small");
  TAMF_REQUIRE(aligned16(index), TAMF_E_ALIGN, "tamf_h2o_index_build: index must be 16-byte aligned");
  int rc = check_device();
  if (rc) return rc;
  return nnp_build(obj_points, total_obj, P, index, (cudaStream_t)stream);
}

extern "C" int tamf_h2o_dist_indexed(const float* verts, const float* obj_traj, const void* index,
                                     const int32_t* obj_first_host, int B, int T, int V, int nobj_max, int P, float* dist,
                                     int64_t* idx, void* stream) {
  TAMF_REQUIRE(index, TAMF_E_BADARG, "tamf_h2o_dist_indexed: null index");
  return h2o_dist_impl(verts, obj_traj, nullptr, index, obj_first_host, B, T, V, nobj_max, P, dist, idx, stream);
}

extern "C" int tamf_h2o_dist(const float* verts, const float* obj_traj, const float* obj_points,
                             const int32_t* obj_first_host, int B, int T, int V, int nobj_max, int P, float* dist,
                             int64_t* idx, void* stream) {
  static const bool env_exhaustive = getenv("TAMF_NN_EXHAUSTIVE") && getenv("TAMF_NN_EXHAUSTIVE")[0] == '1';
  TAMF_REQUIRE(obj_points && obj_first_host && B > 0, TAMF_E_BADARG, "tamf_h2o_dist: null pointer");
  if (env_exhaustive || P > NNP_MAXP || P <= 0)
    return h2o_dist_impl(verts, obj_traj, obj_points, nullptr, obj_first_host, B, T, V, nobj_max, P, dist, idx, stream);
  // one-shot form: the index lives in a grow-only device scratch of this (device, stream) and is rebuilt on every call
  // (the allocation happens on first use / growth only; hot loops use tamf_h2o_index_build + tamf_h2o_dist_indexed)
  const int total_obj = obj_first_host[B];
  TAMF_REQUIRE(total_obj > 0, TAMF_E_BADARG, "tamf_h2o_dist: no objects");
  const size_t need = nnp_index_bytes(total_obj, P);
  int rc = check_device();
  if (rc) return rc;
  void* d_scratch = nullptr;
  if ((rc = nn_scratch((cudaStream_t)stream, 1, need, &d_scratch))) return rc;
  if ((rc = nnp_build(obj_points, total_obj, P, d_scratch, (cudaStream_t)stream))) return rc;
  return h2o_dist_impl(verts, obj_traj, nullptr, d_scratch, obj_first_host, B, T, V, nobj_max, P, dist, idx, stream);
}

extern "C" int tamf_h2o_dist_exhaustive(const float* verts, const float* obj_traj, const float* obj_points,
                                        const int32_t* obj_first_host, int B, int T, int V, int nobj_max, int P,
                                        float* dist, int64_t* idx, void* stream) {
  return h2o_dist_impl(verts, obj_traj, obj_points, nullptr, obj_first_host, B, T, V, nobj_max, P, dist, idx, stream);
}
