/* TEST INFRASTRUCTURE ONLY -- plain-C oracle for the chamfer nearest-neighbour query (K=1).
 *
 * Restates what thirdparty/chamfer_distance/chamfer_distance/chamfer_distance.py:147-162 asks of
 * pytorch3d.ops.knn_points(x, y, K=1): for every x[n,i] the squared L2 distance to, and the index of,
 * its nearest y[n,j].  pytorch3d==0.7.2 (requirements.dist.txt:331) is not vendored and not installed,
 * so its arithmetic cannot be executed here: PARITY UNPINNED.  The arithmetic is therefore *defined*
 * here: fp32; diff = a - b per axis; d = (dx*dx + dy*dy) + dz*dz with every operation
 * rounded to fp32 (compile with -ffp-contract=off); strict '<' scan in ascending j so the lowest index
 * wins ties.  tests/ check this file bit-for-bit against the numpy statement of the same rule.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC -fopenmp -o libtamf_oracle.so nn_oracle.c
 */
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void oracle_nn_query(const float* x, const float* y, int N, int P1, int P2, float* d2, int64_t* idx, int threads) {
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
  long total = (long)N * P1;
#pragma omp parallel for schedule(static)
  for (long q = 0; q < total; ++q) {
    int n = (int)(q / P1);
    const float* xp = x + 3 * q;
    const float* yb = y + (size_t)3 * P2 * n;
    volatile float vx = xp[0], vy = xp[1], vz = xp[2];
    float px = vx, py = vy, pz = vz;
    float best = 0.f;
    int64_t bi = -1;
    for (int j = 0; j < P2; ++j) {
      float dx = px - yb[3 * j + 0];
      float dy = py - yb[3 * j + 1];
      float dz = pz - yb[3 * j + 2];
      float a = dx * dx;
      float b = dy * dy;
      float c = dz * dz;
      float s = a + b;
      float d = s + c;
      if (bi < 0 || d < best) { best = d; bi = j; }
    }
    d2[q] = best;
    idx[q] = bi;
  }
}

/* Per-frame rigid transform of a canonical cloud, then NN: the fused contract of tamf_h2o_dist.
 * world = R_t p + t_t is computed exactly as torch.matmul on fp32 does for a K=3 dot product on CPU
 * cannot be pinned (BLAS order); tests compare *distances* by value for this entry, indices only
 * through oracle_nn_query on materialised points. */
