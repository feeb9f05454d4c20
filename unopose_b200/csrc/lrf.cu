// Global (per-cloud) local reference frame: the whole `get_batch_lrf` of the reference models
// (core/unopose/model/oneref_grf_predator_pose_estimation_model.py:78-93 -> LRF.forward,
// core/unopose/utils/model_utils.py:777-823) in ONE kernel, one CTA per cloud:
//   c = mean(p);  r = max |p - c|  (or a given radius);  cov = sum (c - p)(c - p)^T / N;
//   z_raw = eigenvector of the smallest eigenvalue (in-register cyclic Jacobi, fp64);
//   sign by the vote  #(z_raw.(c - p) > 1e-3) - #(z_raw.(c - p) < -1e-3) < 0  ->  z;
//   x = normalise( sum (r - |q|)^2 (z.q)^2 (q - (z.q) z) ),  q = p - c;   y = x cross z;
//   out = [x y z]^T q / r.
// The reference spends ~25 torch launches here (two reductions, a bmm, a cuSOLVER SVD, ~15 elementwise
// passes over (B,3,N)); the sums are accumulated in fp64 here (torch: fp32 trees).  The sign of z_raw is
// immaterial: flipping it flips the vote, and the vote decides the final direction (unless it is exactly 0).
#include <math.h>

#include "common.cuh"
#include "launch_count.h"
#include "solver3.cuh"
#include "../../include/unopose_b200.h"

namespace upk {

constexpr int LRF_THREADS = 512;

template <int N>
__device__ __forceinline__ void lrf_block_sum(double (&v)[N], double* s_buf /* N * 32 */) {
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

// Eigenvector of the smallest eigenvalue of a symmetric 3x3 matrix, WITH THE SIGN cuSOLVER's batched Jacobi SVD gives
// the last column of V (what the reference's `torch.svd(xxt)[2][..., -1]` is on a GPU, pointnet2_utils.py:451-452).
// The sign only matters when the +-1e-3 vote below ties: the reference then keeps the raw sign of its SVD.  Measured on
// B200 (scripts/r2_parity_probe.py, profiles/r2_lrf_sign.json): a two-sided cyclic Jacobi that starts from V = I, takes
// the pivots in the order (0,1), (1,2), (0,2) with the inner rotation (|angle| <= pi/4), and then sorts the eigenvalues
// in descending order reproduces the sign of all three columns of cuSOLVER's V for every full-rank covariance sampled
// (100 % of 4 682 random centres and of the tied centres at r = 0.2 / 256; the pivot order (0,1), (0,2), (1,2) only
// 94 %).  Covariances that are singular to fp32 precision (balls of 1-3 distinct points) have a null direction whose
// sign is rounding noise in every implementation.
__host__ __device__ __forceinline__ void sym3_min_eigenvector(double m00, double m11, double m22, double m01,
                                                              double m02, double m12, double* v) {
  double v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};
  const double scale = fabs(m00) + fabs(m11) + fabs(m22);
#pragma unroll 1
  for (int sweep = 0; sweep < 10; ++sweep) {
    double off = fabs(m01) + fabs(m02) + fabs(m12);
    if (off <= 1e-22 * scale) break;
    jacobi_rotate(m00, m11, m01, m02, m12, v0, v1);   // (p,q) = (0,1)
    jacobi_rotate(m11, m22, m12, m01, m02, v1, v2);   // (1,2)
    jacobi_rotate(m00, m22, m02, m01, m12, v0, v2);   // (0,2)
  }
  // last column after a stable descending sort: the smallest eigenvalue, the highest index among equal ones
  const int i = (m22 <= m00 && m22 <= m11) ? 2 : ((m11 <= m00) ? 1 : 0);
#pragma unroll
  for (int k = 0; k < 3; ++k) v[k] = i == 0 ? v0[k] : (i == 1 ? v1[k] : v2[k]);
}

__global__ void __launch_bounds__(LRF_THREADS)
k_global_lrf(const float* __restrict__ pts, const float* __restrict__ radius, int n, float eps,
             float* __restrict__ out, float* __restrict__ frame_out) {
  __shared__ double s_buf[6 * 32];
  __shared__ float s_max[32];
  const int b = blockIdx.x;
  const float* P = pts + (size_t)b * n * 3;
  // ---- centroid
  double c3[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < n; i += LRF_THREADS) {
    c3[0] += P[i * 3 + 0]; c3[1] += P[i * 3 + 1]; c3[2] += P[i * 3 + 2];
  }
  lrf_block_sum<3>(c3, s_buf);
  const float cx = (float)(c3[0] / n), cy = (float)(c3[1] / n), cz = (float)(c3[2] / n);
  // ---- radius (max norm of the centred cloud) and covariance of (c - p)
  float mx = 0.f;
  double cv[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += LRF_THREADS) {
    const float x = cx - P[i * 3 + 0], y = cy - P[i * 3 + 1], z = cz - P[i * 3 + 2];
    mx = fmaxf(mx, sqrtf(x * x + y * y + z * z));
    cv[0] += (double)(x * x); cv[1] += (double)(y * y); cv[2] += (double)(z * z);
    cv[3] += (double)(x * y); cv[4] += (double)(x * z); cv[5] += (double)(y * z);
  }
  mx = warp_max(mx);
  lrf_block_sum<6>(cv, s_buf);   // (its barriers also order the s_max exchange below)
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = 0.f;
  for (int w = 0; w < LRF_THREADS / 32; ++w) mx = fmaxf(mx, s_max[w]);
  const float r = radius ? radius[b] : mx;
  double zr[3];
  sym3_min_eigenvector(cv[0] / n, cv[1] / n, cv[2] / n, cv[3] / n, cv[4] / n, cv[5] / n, zr);
  float z0 = (float)zr[0], z1 = (float)zr[1], z2 = (float)zr[2];
  // ---- sign vote
  double vote[1] = {0.0};
  for (int i = threadIdx.x; i < n; i += LRF_THREADS) {
    const float h = z0 * (cx - P[i * 3 + 0]) + z1 * (cy - P[i * 3 + 1]) + z2 * (cz - P[i * 3 + 2]);
    vote[0] += (h > 1e-3f ? 1.0 : 0.0) - (h < -1e-3f ? 1.0 : 0.0);
  }
  lrf_block_sum<1>(vote, s_buf);
  if (vote[0] < 0.0) { z0 = -z0; z1 = -z1; z2 = -z2; }
  // ---- x axis: weighted sum of the in-plane components
  double xd[3] = {0, 0, 0};
  for (int i = threadIdx.x; i < n; i += LRF_THREADS) {
    const float qx = P[i * 3 + 0] - cx, qy = P[i * 3 + 1] - cy, qz = P[i * 3 + 2] - cz;
    const float h = z0 * qx + z1 * qy + z2 * qz;
    const float d = sqrtf(qx * qx + qy * qy + qz * qz);
    const float a = (r - d) * (r - d) * (h * h);
    xd[0] += (double)(a * (qx - h * z0)); xd[1] += (double)(a * (qy - h * z1)); xd[2] += (double)(a * (qz - h * z2));
  }
  lrf_block_sum<3>(xd, s_buf);
  const float x0f = (float)xd[0], x1f = (float)xd[1], x2f = (float)xd[2];
  const float xn = sqrtf(x0f * x0f + x1f * x1f + x2f * x2f) + eps;
  const float x0 = x0f / xn, x1 = x1f / xn, x2 = x2f / xn;
  const float y0 = x1 * z2 - x2 * z1, y1 = x2 * z0 - x0 * z2, y2 = x0 * z1 - x1 * z0;   // y = x cross z
  if (frame_out && threadIdx.x == 0) {
    float* f = frame_out + (size_t)b * 13;   // columns x | y | z, centre, radius
    f[0] = x0; f[1] = y0; f[2] = z0; f[3] = x1; f[4] = y1; f[5] = z1; f[6] = x2; f[7] = y2; f[8] = z2;
    f[9] = cx; f[10] = cy; f[11] = cz; f[12] = r;
  }
  // ---- coordinates in the frame
  float* O = out + (size_t)b * n * 3;
  for (int i = threadIdx.x; i < n; i += LRF_THREADS) {
    const float qx = (P[i * 3 + 0] - cx) / r, qy = (P[i * 3 + 1] - cy) / r, qz = (P[i * 3 + 2] - cz) / r;
    O[i * 3 + 0] = x0 * qx + x1 * qy + x2 * qz;
    O[i * 3 + 1] = y0 * qx + y1 * qy + y2 * qz;
    O[i * 3 + 2] = z0 * qx + z1 * qy + z2 * qz;
  }
}

// ---------------------------------------------------------------- per-centre frames of QueryAndLRFGroup
// LRF_batch (core/unopose/model/pointnet2/pointnet2_utils.py:429-481) + the feature assembly of
// QueryAndLRFGroup.forward (:556-571) in one kernel, one warp per centre:
//   grouped [b][3][n][ns] (absolute neighbour coordinates), centres [b][n][3] (the `xyz` the frames are anchored at),
//   new_xyz [b][n][3] (subtracted for the raw offsets)  ->  out [b][3 or 6][n][ns] = [grouped - new_xyz (/r) | frame coordinates]
// The reference spends a cuSOLVER batched SVD (4.1 ms per call at B = 16, n = 2048) and ~15 elementwise passes over
// (B,n,3,ns) here.  Sums in fp64.  z = eigenvector of the smallest covariance eigenvalue, sign by the +-1e-3 vote;
// when the vote is exactly 0 the reference keeps whatever sign its SVD returned (LAPACK and cuSOLVER disagree) —
// sym3_min_eigenvector reproduces the sign cuSOLVER returns, i.e. the reference's GPU path.
constexpr int LG_WARPS = 8;

__global__ void __launch_bounds__(LG_WARPS * 32)
k_lrf_group(const float* __restrict__ centres, const float* __restrict__ new_xyz, const float* __restrict__ grouped,
            int n, int ns, float r, float eps, int use_xyz, int normalize_xyz, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * LG_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const size_t plane = (size_t)n * ns;
  const float* gx = grouped + ((size_t)b * 3 * n + i) * ns;
  const float* gy = gx + plane;
  const float* gz = gy + plane;
  const float* C = centres + ((size_t)b * n + i) * 3;
  const float cx = C[0], cy = C[1], cz = C[2];
  // covariance of (p - p_j)
  double cv[6] = {0, 0, 0, 0, 0, 0};
  for (int j = lane; j < ns; j += 32) {
    const float x = cx - gx[j], y = cy - gy[j], z = cz - gz[j];
    cv[0] += (double)(x * x); cv[1] += (double)(y * y); cv[2] += (double)(z * z);
    cv[3] += (double)(x * y); cv[4] += (double)(x * z); cv[5] += (double)(y * z);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) cv[k] = warp_sum(cv[k]) / ns;
  double zr[3];
  sym3_min_eigenvector(cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], zr);   // identical in every lane
  float z0 = (float)zr[0], z1 = (float)zr[1], z2 = (float)zr[2];
  int vote = 0;
  for (int j = lane; j < ns; j += 32) {
    const float h = z0 * (cx - gx[j]) + z1 * (cy - gy[j]) + z2 * (cz - gz[j]);
    vote += (h > 1e-3f ? 1 : 0) - (h < -1e-3f ? 1 : 0);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vote += __shfl_xor_sync(0xffffffffu, vote, o);
  if (vote < 0) { z0 = -z0; z1 = -z1; z2 = -z2; }
  // x axis: sum (r - |q|)^2 (z.q)^2 (q - (z.q) z),  q = p_j - p
  double xd[3] = {0, 0, 0};
  for (int j = lane; j < ns; j += 32) {
    const float qx = gx[j] - cx, qy = gy[j] - cy, qz = gz[j] - cz;
    const float h = z0 * qx + z1 * qy + z2 * qz;
    const float d = sqrtf(qx * qx + qy * qy + qz * qz);
    const float a = (r - d) * (r - d) * (h * h);
    xd[0] += (double)(a * (qx - h * z0)); xd[1] += (double)(a * (qy - h * z1)); xd[2] += (double)(a * (qz - h * z2));
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) xd[k] = warp_sum(xd[k]);
  const float x0f = (float)xd[0], x1f = (float)xd[1], x2f = (float)xd[2];
  const float xn = sqrtf(x0f * x0f + x1f * x1f + x2f * x2f) + eps;
  const float x0 = x0f / xn, x1 = x1f / xn, x2 = x2f / xn;
  const float y0 = x1 * z2 - x2 * z1, y1 = x2 * z0 - x0 * z2, y2 = x0 * z1 - x1 * z0;   // y = x cross z
  const int cout = use_xyz ? 6 : 3;
  float* O = out + ((size_t)b * cout * n + i) * ns;
  float* L = O + (use_xyz ? 3 * plane : 0);
  const float* Q = new_xyz + ((size_t)b * n + i) * 3;
  const float nx = Q[0], ny = Q[1], nz = Q[2];
  // `x / self.r_lrf` with a Python scalar: ATen's CUDA division multiplies by the fp32 reciprocal
  const float inv_r = __fdiv_rn(1.0f, r);
  for (int j = lane; j < ns; j += 32) {
    const float px = gx[j], py = gy[j], pz = gz[j];
    if (use_xyz) {
      float ox = px - nx, oy = py - ny, oz = pz - nz;
      if (normalize_xyz) { ox *= inv_r; oy *= inv_r; oz *= inv_r; }
      O[j] = ox; O[plane + j] = oy; O[2 * plane + j] = oz;
    }
    const float qx = (px - cx) * inv_r, qy = (py - cy) * inv_r, qz = (pz - cz) * inv_r;
    L[j] = x0 * qx + x1 * qy + x2 * qz;
    L[plane + j] = y0 * qx + y1 * qy + y2 * qz;
    L[2 * plane + j] = z0 * qx + z1 * qy + z2 * qz;
  }
}

}  // namespace upk

using namespace upk;

extern "C" int upk_global_lrf(const float* pts, const float* radius, int b, int n, float eps, float* out,
                              float* frame_out, upk_stream_t stream) {
  if (b < 0 || n <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  if (!pts || !out) return UPK_ERR_INVALID_ARG;
  k_global_lrf<<<b, LRF_THREADS, 0, (cudaStream_t)stream>>>(pts, radius, n, eps, out, frame_out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

// Host-side evaluation of the frame solver (same source as the device code): lets the CPU test-suite pin the sign
// convention against (covariance, V) pairs recorded from torch.svd / cuSOLVER on a B200 (tests/golden/lrf_svd_sign.npz).
// cov: n x 6 doubles (xx, yy, zz, xy, xz, yz); z_out: n x 3.
extern "C" int upk_host_lrf_z_axis(const double* cov, int n, double* z_out) {
  if (n < 0 || (n > 0 && (!cov || !z_out))) return UPK_ERR_INVALID_ARG;
  for (int i = 0; i < n; ++i)
    sym3_min_eigenvector(cov[i * 6 + 0], cov[i * 6 + 1], cov[i * 6 + 2], cov[i * 6 + 3], cov[i * 6 + 4], cov[i * 6 + 5],
                         z_out + (size_t)i * 3);
  return UPK_OK;
}

extern "C" int upk_lrf_group(const float* centres, const float* new_xyz, const float* grouped, int b, int n, int ns,
                             float r_lrf, float eps, int use_xyz, int normalize_xyz, float* out, upk_stream_t stream) {
  if (b < 0 || n <= 0 || ns <= 0 || !(r_lrf > 0.f)) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  if (!centres || !new_xyz || !grouped || !out) return UPK_ERR_INVALID_ARG;
  k_lrf_group<<<dim3(ceil_div(n, LG_WARPS), b), LG_WARPS * 32, 0, (cudaStream_t)stream>>>(
      centres, new_xyz, grouped, n, ns, r_lrf, eps, use_xyz ? 1 : 0, normalize_xyz ? 1 : 0, out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}
