// In-register 3x3 Procrustes rotation solver (kernel family (3)).
//
// Replaces the reference's  U,S,V = torch.svd(H);  R = V diag(1,1,sign det(V U^T)) U^T
// (core/unopose/utils/model_utils.py:722-727; cuSOLVER batched SVD + batched LU
// for the determinant) by a branch-light closed form evaluated by ONE thread in
// fp64 registers:
//   M = H^T H  --cyclic Jacobi-->  eigenvectors v1,v2 of the two largest
//   eigenvalues;  u_k = H v_k / |H v_k| (Gram-Schmidt'ed);
//   R = v1 u1^T + v2 u2^T + (v1 x v2)(u1 x u2)^T.
// The third term carries the reflection fix implicitly: with v3 := v1 x v2 and
// u3 := u1 x u2 both bases are right-handed, so det R = +1, which is exactly
// what diag(1,1,sign det(V U^T)) does to the smallest singular direction.
// Only the two dominant singular directions are ever used, so rank-2 inputs
// (every 3-point hypothesis after centring) are handled without special cases;
// rank-1 / rank-0 inputs (repeated correspondences) get a valid
// arbitrary completion instead of NaNs.
#pragma once
#include <cuda_runtime.h>

namespace upk {

__host__ __device__ __forceinline__ void jacobi_rotate(double& app, double& aqq, double& apq, double& arp,
                                              double& arq, double* vp, double* vq) {
  // annihilate a[p][q] of the symmetric matrix; r is the third index
  if (apq == 0.0) return;
  double theta = (aqq - app) / (2.0 * apq);
  double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
  double c = 1.0 / sqrt(t * t + 1.0);
  double s = t * c;
  app -= t * apq;
  aqq += t * apq;
  apq = 0.0;
  double nrp = c * arp - s * arq;
  double nrq = s * arp + c * arq;
  arp = nrp;
  arq = nrq;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double a = vp[k], b = vq[k];
    vp[k] = c * a - s * b;
    vq[k] = s * a + c * b;
  }
}

// H row-major (H[a*3+b]); R row-major out.  R maximises tr(R H) over SO(3).
__host__ __device__ __forceinline__ void procrustes_rotation(const double* H, double* R) {
  // M = H^T H
  double m00 = H[0] * H[0] + H[3] * H[3] + H[6] * H[6];
  double m11 = H[1] * H[1] + H[4] * H[4] + H[7] * H[7];
  double m22 = H[2] * H[2] + H[5] * H[5] + H[8] * H[8];
  double m01 = H[0] * H[1] + H[3] * H[4] + H[6] * H[7];
  double m02 = H[0] * H[2] + H[3] * H[5] + H[6] * H[8];
  double m12 = H[1] * H[2] + H[4] * H[5] + H[7] * H[8];
  double v0[3] = {1, 0, 0}, v1[3] = {0, 1, 0}, v2[3] = {0, 0, 1};  // eigenvector columns
  const double scale = m00 + m11 + m22;
#pragma unroll 1
  for (int sweep = 0; sweep < 8; ++sweep) {
    double off = fabs(m01) + fabs(m02) + fabs(m12);
    if (off <= 1e-22 * scale) break;
    jacobi_rotate(m00, m11, m01, m02, m12, v0, v1);  // (p,q)=(0,1), r=2: a[r][p]=m02, a[r][q]=m12
    jacobi_rotate(m00, m22, m02, m01, m12, v0, v2);  // (0,2), r=1: a[r][p]=m01, a[r][q]=m12
    jacobi_rotate(m11, m22, m12, m01, m02, v1, v2);  // (1,2), r=0: a[r][p]=m01, a[r][q]=m02
  }
  // pick the two largest eigenvalues (register selects, no local arrays)
  const int i1 = (m00 >= m11) ? ((m00 >= m22) ? 0 : 2) : ((m11 >= m22) ? 1 : 2);
  const double l0 = i1 == 0 ? -1.0 : m00, l1 = i1 == 1 ? -1.0 : m11, l2 = i1 == 2 ? -1.0 : m22;
  const int i2 = (l0 >= l1) ? ((l0 >= l2) ? 0 : 2) : ((l1 >= l2) ? 1 : 2);
  double va[3], vb[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    va[k] = i1 == 0 ? v0[k] : (i1 == 1 ? v1[k] : v2[k]);
    vb[k] = i2 == 0 ? v0[k] : (i2 == 1 ? v1[k] : v2[k]);
  }
  // u1 = H v1
  double ua[3], ub[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    ua[r] = H[r * 3 + 0] * va[0] + H[r * 3 + 1] * va[1] + H[r * 3 + 2] * va[2];
    ub[r] = H[r * 3 + 0] * vb[0] + H[r * 3 + 1] * vb[1] + H[r * 3 + 2] * vb[2];
  }
  double na = sqrt(ua[0] * ua[0] + ua[1] * ua[1] + ua[2] * ua[2]);
  if (!(na > 1e-150)) {  // H == 0: any rotation is optimal
    R[0] = R[4] = R[8] = 1.0;
    R[1] = R[2] = R[3] = R[5] = R[6] = R[7] = 0.0;
    return;
  }
  double ia = 1.0 / na;
  ua[0] *= ia; ua[1] *= ia; ua[2] *= ia;
  double dp = ua[0] * ub[0] + ua[1] * ub[1] + ua[2] * ub[2];
  ub[0] -= dp * ua[0]; ub[1] -= dp * ua[1]; ub[2] -= dp * ua[2];
  double nb = sqrt(ub[0] * ub[0] + ub[1] * ub[1] + ub[2] * ub[2]);
  if (nb > 1e-13 * na) {
    double ib = 1.0 / nb;
    ub[0] *= ib; ub[1] *= ib; ub[2] *= ib;
  } else {
    // rank 1: complete u1 with any unit vector orthogonal to it
    int k = fabs(ua[0]) <= fabs(ua[1]) ? (fabs(ua[0]) <= fabs(ua[2]) ? 0 : 2)
                                        : (fabs(ua[1]) <= fabs(ua[2]) ? 1 : 2);
    double d2 = k == 0 ? ua[0] : (k == 1 ? ua[1] : ua[2]);
    ub[0] = (k == 0 ? 1.0 : 0.0) - d2 * ua[0];
    ub[1] = (k == 1 ? 1.0 : 0.0) - d2 * ua[1];
    ub[2] = (k == 2 ? 1.0 : 0.0) - d2 * ua[2];
    double ib = 1.0 / sqrt(ub[0] * ub[0] + ub[1] * ub[1] + ub[2] * ub[2]);
    ub[0] *= ib; ub[1] *= ib; ub[2] *= ib;
  }
  double uc[3] = {ua[1] * ub[2] - ua[2] * ub[1], ua[2] * ub[0] - ua[0] * ub[2], ua[0] * ub[1] - ua[1] * ub[0]};
  double vc[3] = {va[1] * vb[2] - va[2] * vb[1], va[2] * vb[0] - va[0] * vb[2], va[0] * vb[1] - va[1] * vb[0]};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int q = 0; q < 3; ++q) R[r * 3 + q] = va[r] * ua[q] + vb[r] * ub[q] + vc[r] * uc[q];
}

}  // namespace upk
