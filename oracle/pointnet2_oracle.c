/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.
 *
 * Scalar CPU restatement of the reference's pointnet2 CUDA kernels
 *   core/unopose/model/pointnet2/_ext_src/src/sampling_gpu.cu
 *   core/unopose/model/pointnet2/_ext_src/src/ball_query_gpu.cu
 *   core/unopose/model/pointnet2/_ext_src/src/group_points_gpu.cu
 *   core/unopose/model/pointnet2/_ext_src/src/interpolate_gpu.cu
 * emulating the kernels THREAD BY THREAD (block size, strided ownership and
 * the shared-memory tournament of FPS included) so that index ties resolve
 * exactly as on the GPU.  The reference has no CPU path for these ops
 * ("CPU not supported", sampling.cpp:39,65,87), so this file is a
 * "restatement"; it is pinned on the GPU box against the reference's own
 * extension compiled unmodified (oracle/_ref, tests/test_pointnet2_gpu.py) and
 * against the fixtures under tests/golden/ that were produced by it.
 *
 * Arithmetic: the reference is compiled with nvcc's default -fmad=true, which
 * contracts  a*a + b*b + c*c  into  fma(c,c, fma(a,a, b*b))  (read off the SASS of
 * the reference extension compiled for sm_100a: FMUL on the y term, FFMA x,
 * FFMA z; SURVEY.md Appendix A.1 has x and y swapped).  Build this file with
 * -ffp-contract=off; the fused operations are spelled out with fmaf().
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static inline float sqdist(float dx, float dy, float dz) {
  return fmaf(dz, dz, fmaf(dx, dx, dy * dy));
}

/* cuda_utils.h:20-24  opt_n_threads */
static int opt_n_threads(int work_size) {
  if (work_size < 1) return 1;
  int pow_2 = (int)(log((double)work_size) / log(2.0));
  int v = 1 << pow_2;
  if (v > 512) v = 512;
  if (v < 1) v = 1;
  return v;
}

int oracle_opt_n_threads(int n) { return opt_n_threads(n); }

/* sampling_gpu.cu:74-178 furthest_point_sampling_kernel<block_size>,
 * one simulated block per batch element.  xyz (b,n,3) -> idxs (b,m). */
void oracle_furthest_point_sampling(const float* xyz, int b, int n, int m, int* idxs) {
  if (m <= 0 || n <= 0) return;
  const int bs = opt_n_threads(n); /* sampling_gpu.cu:182 */
  float* temp = (float*)malloc(sizeof(float) * (size_t)n);
  float* dists = (float*)malloc(sizeof(float) * (size_t)bs);
  int* dists_i = (int*)malloc(sizeof(int) * (size_t)bs);
  for (int bi = 0; bi < b; ++bi) {
    const float* dataset = xyz + (size_t)bi * n * 3;
    int* out = idxs + (size_t)bi * m;
    for (int k = 0; k < n; ++k) temp[k] = 1e10f; /* sampling.cpp:78-80 */
    int old = 0;
    out[0] = old; /* :90-91 */
    for (int j = 1; j < m; ++j) {
      float x1 = dataset[old * 3 + 0], y1 = dataset[old * 3 + 1], z1 = dataset[old * 3 + 2];
      for (int tid = 0; tid < bs; ++tid) { /* :95-118 per-thread strided scan */
        int besti = 0;
        float best = -1.f;
        for (int k = tid; k < n; k += bs) {
          float x2 = dataset[k * 3 + 0], y2 = dataset[k * 3 + 1], z2 = dataset[k * 3 + 2];
          float d = sqdist(x2 - x1, y2 - y1, z2 - z1);
          float d2 = fminf(d, temp[k]); /* CUDA min(float,float) == fminf */
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      /* :120-173 tree; __update (:64-70) keeps the lower slot on ties */
      for (int s = bs / 2; s >= 1; s >>= 1) {
        for (int tid = 0; tid < s; ++tid) {
          float v1 = dists[tid], v2 = dists[tid + s];
          int i1 = dists_i[tid], i2 = dists_i[tid + s];
          dists[tid] = v1 > v2 ? v1 : (v2 > v1 ? v2 : (v1 == v1 ? v1 : v2)); /* max() */
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
  }
  free(temp);
  free(dists);
  free(dists_i);
}

/* sampling_gpu.cu:13-25 gather_points_kernel: points (b,c,n), idx (b,m) -> out (b,c,m) */
void oracle_gather_points(const float* points, const int* idx, int b, int c, int n, int m,
                          float* out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j) {
        int a = idx[(size_t)i * m + j];
        out[((size_t)i * c + l) * m + j] = points[((size_t)i * c + l) * n + a];
      }
}

/* sampling_gpu.cu:39-52 gather_points_grad_kernel (scatter-add into zeros) */
void oracle_gather_points_grad(const float* grad_out, const int* idx, int b, int c, int n,
                               int m, float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < m; ++j) {
        int a = idx[(size_t)i * m + j];
        grad_points[((size_t)i * c + l) * n + a] += grad_out[((size_t)i * c + l) * m + j];
      }
}

/* ball_query_gpu.cu:14-49 query_ball_point_kernel; idx pre-zeroed (ball_query.cpp:24-26) */
void oracle_ball_query(const float* new_xyz, const float* xyz, int b, int n, int m,
                       float radius, int nsample, int* idx) {
  memset(idx, 0, sizeof(int) * (size_t)b * m * nsample);
  const float radius2 = radius * radius;
  for (int bi = 0; bi < b; ++bi) {
    const float* p = xyz + (size_t)bi * n * 3;
    const float* q = new_xyz + (size_t)bi * m * 3;
    int* o = idx + (size_t)bi * m * nsample;
    for (int j = 0; j < m; ++j) {
      float nx = q[j * 3 + 0], ny = q[j * 3 + 1], nz = q[j * 3 + 2];
      for (int k = 0, cnt = 0; k < n && cnt < nsample; ++k) {
        float d2 = sqdist(nx - p[k * 3 + 0], ny - p[k * 3 + 1], nz - p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) o[(size_t)j * nsample + l] = k;
          o[(size_t)j * nsample + cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

/* group_points_gpu.cu:13-33: points (b,c,n), idx (b,npoints,nsample) -> out (b,c,npoints,nsample) */
void oracle_group_points(const float* points, const int* idx, int b, int c, int n, int npoints,
                         int nsample, float* out) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) {
          int ii = idx[((size_t)bi * npoints + j) * nsample + k];
          out[(((size_t)bi * c + l) * npoints + j) * nsample + k] =
              points[((size_t)bi * c + l) * n + ii];
        }
}

/* group_points_gpu.cu:48-69 group_points_grad_kernel */
void oracle_group_points_grad(const float* grad_out, const int* idx, int b, int c, int n,
                              int npoints, int nsample, float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * n);
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        for (int k = 0; k < nsample; ++k) {
          int ii = idx[((size_t)bi * npoints + j) * nsample + k];
          grad_points[((size_t)bi * c + l) * n + ii] +=
              grad_out[(((size_t)bi * c + l) * npoints + j) * nsample + k];
        }
}

/* interpolate_gpu.cu:14-64 three_nn_kernel: bests kept in double (:32) */
void oracle_three_nn(const float* unknown, const float* known, int b, int n, int m,
                     float* dist2, int* idx) {
  for (int bi = 0; bi < b; ++bi) {
    const float* u = unknown + (size_t)bi * n * 3;
    const float* kn = known + (size_t)bi * m * 3;
    for (int j = 0; j < n; ++j) {
      float ux = u[j * 3 + 0], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        float d = sqdist(ux - kn[k * 3 + 0], uy - kn[k * 3 + 1], uz - kn[k * 3 + 2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      size_t o = ((size_t)bi * n + j) * 3;
      dist2[o + 0] = (float)best1; dist2[o + 1] = (float)best2; dist2[o + 2] = (float)best3;
      idx[o + 0] = besti1; idx[o + 1] = besti2; idx[o + 2] = besti3;
    }
  }
}

/* interpolate_gpu.cu:77-106 three_interpolate_kernel; p1*w1 + p2*w2 + p3*w3 contracted */
void oracle_three_interpolate(const float* points, const int* idx, const float* weight, int b,
                              int c, int m, int n, float* out) {
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        const int* ix = idx + ((size_t)bi * n + j) * 3;
        const float* w = weight + ((size_t)bi * n + j) * 3;
        const float* p = points + ((size_t)bi * c + l) * m;
        out[((size_t)bi * c + l) * n + j] =
            fmaf(p[ix[2]], w[2], fmaf(p[ix[0]], w[0], p[ix[1]] * w[1]));
      }
}

/* interpolate_gpu.cu:121-148 three_interpolate_grad_kernel */
void oracle_three_interpolate_grad(const float* grad_out, const int* idx, const float* weight,
                                   int b, int c, int n, int m, float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)b * c * m);
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < n; ++j) {
        const int* ix = idx + ((size_t)bi * n + j) * 3;
        const float* w = weight + ((size_t)bi * n + j) * 3;
        float g = grad_out[((size_t)bi * c + l) * n + j];
        float* gp = grad_points + ((size_t)bi * c + l) * m;
        gp[ix[0]] += g * w[0];
        gp[ix[1]] += g * w[1];
        gp[ix[2]] += g * w[2];
      }
}
