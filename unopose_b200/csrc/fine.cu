// Fine pose solve: soft correspondences -> ONE weighted Kabsch per instance -> inlier score.
// Reference: compute_fine_Rt[_overlap], core/unopose/utils/model_utils.py:493-566.
#include <math.h>

#include "common.cuh"
#include "launch_count.h"
#include "pose_internal.h"
#include "solver3.cuh"
#include "../../include/unopose_b200.h"

namespace upk {

constexpr int FK_THREADS = 512;

template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], double* s_buf /* N * 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < N; ++i) s_buf[i * 32 + warp] = v[i];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += s_buf[i * 32 + w];  // fixed order, identical in all threads
    v[i] = t;
  }
}

// weighted_procrustes (model_utils.py:704-730) for one instance per CTA.
//   src = soft correspondences (B,N,3), ref = query points (B,N,3), weights = row sums of the
//   masked assignment, zeroed below `thresh` (0.001 overlap variant :528 / 0.0 plain :502),
//   normalised by (sum + 1e-5).  fp64 block reductions, fp64 3x3 solve, fp32 outputs.
__global__ void __launch_bounds__(FK_THREADS)
k_weighted_kabsch(const float* __restrict__ src, const float* __restrict__ ref,
                  const float* __restrict__ weights, int n, float thresh, float eps,
                  float* __restrict__ R_out, float* __restrict__ t_out, int* __restrict__ zero3 = nullptr) {
  __shared__ double s_buf[9 * 32];
  const int b = blockIdx.x;
  if (zero3 && threadIdx.x < 3) zero3[b * 3 + threadIdx.x] = 0;   // counters of the inlier launch that follows
  src += (size_t)b * n * 3;
  ref += (size_t)b * n * 3;
  const float* wt = weights ? weights + (size_t)b * n : nullptr;
  double acc1[1] = {0.0};
  for (int i = threadIdx.x; i < n; i += FK_THREADS) {
    float w = wt ? wt[i] : 1.f;
    if (w < thresh) w = 0.f;
    acc1[0] += (double)w;
  }
  block_sum<1>(acc1, s_buf);
  const float denom = (float)acc1[0] + eps;
  double c[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += FK_THREADS) {
    float w = wt ? wt[i] : 1.f;
    if (w < thresh) w = 0.f;
    float wn = w / denom;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      c[a] += (double)(src[i * 3 + a] * wn);
      c[3 + a] += (double)(ref[i * 3 + a] * wn);
    }
  }
  block_sum<6>(c, s_buf);
  float cs[3] = {(float)c[0], (float)c[1], (float)c[2]};
  float cr[3] = {(float)c[3], (float)c[4], (float)c[5]};
  double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += FK_THREADS) {
    float w = wt ? wt[i] : 1.f;
    if (w < thresh) w = 0.f;
    float wn = w / denom;
    float s0 = src[i * 3 + 0] - cs[0], s1 = src[i * 3 + 1] - cs[1], s2 = src[i * 3 + 2] - cs[2];
    float r0 = wn * (ref[i * 3 + 0] - cr[0]), r1 = wn * (ref[i * 3 + 1] - cr[1]), r2 = wn * (ref[i * 3 + 2] - cr[2]);
    H[0] += (double)(s0 * r0); H[1] += (double)(s0 * r1); H[2] += (double)(s0 * r2);
    H[3] += (double)(s1 * r0); H[4] += (double)(s1 * r1); H[5] += (double)(s1 * r2);
    H[6] += (double)(s2 * r0); H[7] += (double)(s2 * r1); H[8] += (double)(s2 * r2);
  }
  block_sum<9>(H, s_buf);
  if (threadIdx.x == 0) {
    // the reference hands a float32 H to the SVD
    double Hf[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) Hf[i] = (double)(float)H[i];
    double Rd[9];
    procrustes_rotation(Hf, Rd);
    float R[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      R[i] = (float)Rd[i];
      R_out[(size_t)b * 9 + i] = R[i];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
      t_out[(size_t)b * 3 + a] =
          cr[a] - fmaf(R[a * 3 + 2], cs[2], fmaf(R[a * 3 + 1], cs[1], R[a * 3 + 0] * cs[0]));
  }
}

// Inlier score (model_utils.py:558-564): X = (pts1 - t) @ R; d_i = NN distance to the model cloud
// (expansion form, clamp, sqrt); score = sum_fg[d<thr] / (n_fg + 1e-8) * n_fg / N1.
// A thread scans for a PAIR of query points (the four LDS.128 of a scan step serve both, two dependency chains
// interleave — nn_min_expansion2, like k_score).  The staged model tile is cut into FS_PARTS interleaved slices, ONE PER
// WARP: every lane of a warp then reads the same shared-memory address.  (Slices across the lanes of a warp — 8 distinct
// addresses per quarter warp — cost 4.0 shared-memory wavefronts per LDS.128 in ncu, against 1.85 for a uniform address;
// the kernel sat at 53 % of the shared-memory pipe with short-scoreboard stalls on 58 % of its samples.)  The minima of
// the slices meet in shared memory; the last CTA of an instance (ticket counter) turns the integer counts into the score.
// 8 warps x 32 pairs = 64 query points per CTA: 512 CTAs at B = 16.
constexpr int FS_THREADS = 256;
constexpr int FS_PARTS = FS_THREADS / 32;           // one model slice per warp
constexpr int FS_QPB = 64;                          // query points per CTA (32 pairs)
constexpr int FS_TILE = 2048;

__global__ void __launch_bounds__(FS_THREADS)
k_fine_inliers(const float* __restrict__ pts1, const float* __restrict__ model,
               const float* __restrict__ w1, const float* __restrict__ Rm, const float* __restrict__ tv,
               int n1, int nm, float thr, int* __restrict__ counters /* [b][3]: inliers, fg, ticket */,
               float* __restrict__ nn_out, float* __restrict__ score) {
  __shared__ __align__(16) float sm_model[4 * FS_TILE];
  __shared__ float s_best[FS_PARTS][FS_QPB];
  __shared__ int s_cnt[2];
  const int b = blockIdx.y;
  const int part = threadIdx.x >> 5, pair = threadIdx.x & 31;
  const int ia = blockIdx.x * FS_QPB + 2 * pair, ib = ia + 1;   // this thread's two query points
  const float* R = Rm + (size_t)b * 9;
  const float* t = tv + (size_t)b * 3;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  float xa[4] = {0.f, 0.f, 0.f, 0.f}, xb[4] = {0.f, 0.f, 0.f, 0.f};
  const bool oka = ia < n1, okb = ib < n1;
  auto moved = [&](int i, float (&x)[4]) {
    const float* p = pts1 + ((size_t)b * n1 + i) * 3;
    const float d0 = p[0] - t[0], d1 = p[1] - t[1], d2 = p[2] - t[2];
    x[0] = fmaf(d2, R[6], fmaf(d1, R[3], d0 * R[0]));
    x[1] = fmaf(d2, R[7], fmaf(d1, R[4], d0 * R[1]));
    x[2] = fmaf(d2, R[8], fmaf(d1, R[5], d0 * R[2]));
    x[3] = sumsq3_torch(x[0], x[1], x[2]);
  };
  if (oka) moved(ia, xa);
  if (okb) moved(ib, xb);
  float best_a = INFINITY, best_b = INFINITY;
  const float* mb = model + (size_t)b * nm * 3;
  for (int j0 = 0; j0 < nm; j0 += FS_TILE) {
    const int tn = min(FS_TILE, nm - j0);
    const int tn_pad = (tn + 4 * FS_PARTS - 1) & ~(4 * FS_PARTS - 1);   // every slice scans a multiple of 4 points
    float* mx = sm_model; float* my = mx + FS_TILE; float* mz = my + FS_TILE; float* mn = mz + FS_TILE;
    __syncthreads();
    stage_model_soa(mb + (size_t)j0 * 3, tn, tn_pad, mx, my, mz, mn);
    __syncthreads();
    float ba, bb;
    nn_min_expansion2(mx, my, mz, mn, tn_pad, xa, xb, ba, bb, 4 * part, 4 * FS_PARTS);
    best_a = fminf(best_a, ba);
    best_b = fminf(best_b, bb);
  }
  s_best[part][2 * pair] = best_a;
  s_best[part][2 * pair + 1] = best_b;
  __syncthreads();
  if (threadIdx.x < FS_QPB) {   // warps 0 and 1: one query point per thread
    const int i = blockIdx.x * FS_QPB + threadIdx.x;
    float best = s_best[0][threadIdx.x];
#pragma unroll
    for (int p = 1; p < FS_PARTS; ++p) best = fminf(best, s_best[p][threadIdx.x]);
    int inl = 0, fg = 0;
    if (i < n1) {
      const float dist = sqrtf(fmaxf(best, 0.f));
      if (nn_out) nn_out[(size_t)b * n1 + i] = dist;
      fg = w1[(size_t)b * n1 + i] > 0.f ? 1 : 0;
      inl = (dist < thr && fg) ? 1 : 0;
    }
    inl = __reduce_add_sync(kFull, inl);
    fg = __reduce_add_sync(kFull, fg);
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(&s_cnt[0], inl);
      atomicAdd(&s_cnt[1], fg);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int* c = counters + b * 3;
    atomicAdd(c + 0, s_cnt[0]);
    atomicAdd(c + 1, s_cnt[1]);
    __threadfence();
    if (atomicAdd(c + 2, 1) == (int)gridDim.x - 1) {   // every CTA of this instance has contributed
      const float fi = (float)atomicAdd(c + 0, 0), ff = (float)atomicAdd(c + 1, 0);
      score[b] = (fi / (ff + 1e-8f)) * (ff / (float)n1);
    }
  }
}

// X = (pts - t) @ R per instance: the query cloud moved into the reference frame by a pose
// (fine module :65-72, compute_fine_Rt :558; torch does it with a broadcast subtract + a K=3 cuBLAS GEMM).
__global__ void __launch_bounds__(256)
k_transform_points(const float* __restrict__ pts, const float* __restrict__ Rm, const float* __restrict__ tv, int n,
                   float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float* R = Rm + (size_t)b * 9;
  const float* t = tv + (size_t)b * 3;
  const float* p = pts + ((size_t)b * n + i) * 3;
  const float d0 = p[0] - t[0], d1 = p[1] - t[1], d2 = p[2] - t[2];
  float* o = out + ((size_t)b * n + i) * 3;
  o[0] = fmaf(d2, R[6], fmaf(d1, R[3], d0 * R[0]));
  o[1] = fmaf(d2, R[7], fmaf(d1, R[4], d0 * R[1]));
  o[2] = fmaf(d2, R[8], fmaf(d1, R[5], d0 * R[2]));
}

struct FineWs {
  AssignWs a;
  float* w1; float* w2;
  float4* rowpart4;
  float* soft; float* asum;
  int* counters;
};

static void carve_fine(Carver& cv, int b, int n1, int n2, const AssignGeom& g, FineWs& w) {
  carve_assign(cv, b, g, w.a);
  w.w1 = cv.take<float>((size_t)b * n1);
  w.w2 = cv.take<float>((size_t)b * n2);
  w.rowpart4 = cv.take<float4>((size_t)b * n1 * g.ntc);
  w.soft = cv.take<float>((size_t)b * n1 * 3);
  w.asum = cv.take<float>((size_t)b * n1);
  w.counters = cv.take<int>((size_t)b * 3);
}

}  // namespace upk

using namespace upk;

extern "C" {

size_t upk_fine_pose_workspace_bytes(int b, int n1, int n2) {
  if (b <= 0 || n1 <= 0 || n2 <= 0) return 0;
  Carver cv(nullptr);
  FineWs w;
  carve_fine(cv, b, n1, n2, assign_geom(n1 + 1, n2 + 1), w);
  return cv.bytes();
}

int upk_weighted_procrustes(const float* src, const float* ref, const float* weights, int b, int n,
                            float weight_thresh, float eps, float* R_out, float* t_out, upk_stream_t stream) {
  if (b < 0 || n <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  k_weighted_kabsch<<<b, FK_THREADS, 0, (cudaStream_t)stream>>>(src, ref, weights, n, weight_thresh, eps, R_out, t_out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

static int fine_pose_impl(const float* atten, int atten_ld, const float* stats, size_t stats_bytes, float temp,
                          const float* score1, int score1_ld, const float* score2, int score2_ld,
                          const float* pts1, const float* pts2, const float* model_pts, int n_model, int b, int n1,
                          int n2, float dis_thres, float weight_thresh, void* workspace, size_t workspace_bytes,
                          float* R_out, float* t_out, float* score_out, const upk_fine_debug* dbg,
                          upk_stream_t stream) {
  if (b < 0 || n1 <= 0 || n2 <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  if (!atten || !pts1 || !pts2 || !workspace || !R_out || !t_out || !score_out) return UPK_ERR_INVALID_ARG;
  if ((score1 == nullptr) != (score2 == nullptr)) return UPK_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  AssignGeom g = assign_geom(n1 + 1, n2 + 1);
  Carver cv(workspace);
  FineWs w;
  carve_fine(cv, b, n1, n2, g, w);
  if (cv.bytes() > workspace_bytes) return UPK_ERR_INVALID_ARG;
  if (!model_pts) { model_pts = pts2; n_model = n2; }
  const int ld = atten_ld > 0 ? atten_ld : n2 + 1;
  if (ld < n2 + 1) return UPK_ERR_INVALID_ARG;
  int rc;
  if (stats) {   // pass 1 came out of the similarity GEMM's epilogue (upk_feature_similarity_stats)
    if (g.TR == 32 || stats_bytes < sim_stats_geom(b, n1 + 1, n2 + 1).total_bytes || !(temp > 0.f)) return UPK_ERR_INVALID_ARG;
    rc = run_fine_labels2_fused(atten, ld, stats, temp, score1, score1_ld, score2, score2_ld, b, g, w.a, w.w1, w.w2, st);
  } else {
    rc = run_assignment_labels(atten, score1, score1_ld, score2, score2_ld, b, g, w.a, w.w1, w.w2, st, true, ld);
  }
  if (rc) return rc;
  if ((rc = run_fine_rowsums(atten, score1, score1_ld, score2, score2_ld, b, g, w.a, w.w1, w.w2, pts2,
                             w.rowpart4, w.soft, w.asum, st, ld)))
    return rc;
  k_weighted_kabsch<<<b, FK_THREADS, 0, st>>>(w.soft, pts1, w.asum, n1, weight_thresh, 1e-5f, R_out, t_out, w.counters);
  dim3 grid(ceil_div(n1, FS_QPB), b);
  k_fine_inliers<<<grid, FS_THREADS, 0, st>>>(pts1, model_pts, w.w1, R_out, t_out, n1, n_model, dis_thres,
                                              w.counters, dbg ? dbg->nn : nullptr, score_out);
  count_launch(2);
  if (dbg) {
    if (dbg->w1) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->w1, w.w1, sizeof(float) * (size_t)b * n1, cudaMemcpyDeviceToDevice, st));
    if (dbg->w2) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->w2, w.w2, sizeof(float) * (size_t)b * n2, cudaMemcpyDeviceToDevice, st));
    if (dbg->soft) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->soft, w.soft, sizeof(float) * (size_t)b * n1 * 3, cudaMemcpyDeviceToDevice, st));
    if (dbg->asum) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->asum, w.asum, sizeof(float) * (size_t)b * n1, cudaMemcpyDeviceToDevice, st));
  }
  UPK_RETURN_LAST_ERROR();
}

int upk_fine_pose(const float* atten, const float* score1, int score1_ld, const float* score2, int score2_ld,
                  const float* pts1, const float* pts2, const float* model_pts, int n_model, int b, int n1,
                  int n2, float dis_thres, float weight_thresh, void* workspace, size_t workspace_bytes,
                  float* R_out, float* t_out, float* score_out, const upk_fine_debug* dbg, upk_stream_t stream) {
  return fine_pose_impl(atten, 0, nullptr, 0, 0.f, score1, score1_ld, score2, score2_ld, pts1, pts2, model_pts, n_model,
                        b, n1, n2, dis_thres, weight_thresh, workspace, workspace_bytes, R_out, t_out, score_out, dbg,
                        stream);
}

int upk_fine_pose_stats(const float* atten, const float* stats, size_t stats_bytes, float temp,
                        const float* score1, int score1_ld, const float* score2, int score2_ld,
                        const float* pts1, const float* pts2, const float* model_pts, int n_model, int b, int n1,
                        int n2, float dis_thres, float weight_thresh, void* workspace, size_t workspace_bytes,
                        float* R_out, float* t_out, float* score_out, const upk_fine_debug* dbg,
                        upk_stream_t stream) {
  if (!stats) return UPK_ERR_INVALID_ARG;
  return fine_pose_impl(atten, 0, stats, stats_bytes, temp, score1, score1_ld, score2, score2_ld, pts1, pts2, model_pts,
                        n_model, b, n1, n2, dis_thres, weight_thresh, workspace, workspace_bytes, R_out, t_out,
                        score_out, dbg, stream);
}

int upk_fine_pose_ld(const float* atten, int atten_ld, const float* stats, size_t stats_bytes, float temp,
                     const float* score1, int score1_ld, const float* score2, int score2_ld,
                     const float* pts1, const float* pts2, const float* model_pts, int n_model, int b, int n1,
                     int n2, float dis_thres, float weight_thresh, void* workspace, size_t workspace_bytes,
                     float* R_out, float* t_out, float* score_out, const upk_fine_debug* dbg,
                     upk_stream_t stream) {
  if (atten_ld < n2 + 1) return UPK_ERR_INVALID_ARG;
  return fine_pose_impl(atten, atten_ld, stats, stats_bytes, temp, score1, score1_ld, score2, score2_ld, pts1, pts2,
                        model_pts, n_model, b, n1, n2, dis_thres, weight_thresh, workspace, workspace_bytes, R_out, t_out,
                        score_out, dbg, stream);
}

int upk_transform_points(const float* pts, const float* R, const float* t, int b, int n, float* out,
                         upk_stream_t stream) {
  if (b < 0 || n < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || n == 0) return UPK_OK;
  if (!pts || !R || !t || !out) return UPK_ERR_INVALID_ARG;
  dim3 grid(ceil_div(n, 256), b);
  k_transform_points<<<grid, 256, 0, (cudaStream_t)stream>>>(pts, R, t, n, out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

// Host-side evaluation of the 3x3 solver (same source as the device code) — lets the CPU
// test-suite check kernel family (3)'s maths against LAPACK without a GPU.
int upk_host_procrustes_rotation(const double* H, int n, double* R_out) {
  if (n < 0 || (n > 0 && (!H || !R_out))) return UPK_ERR_INVALID_ARG;
  for (int i = 0; i < n; ++i) procrustes_rotation(H + (size_t)i * 9, R_out + (size_t)i * 9);
  return UPK_OK;
}

}  // extern "C"
