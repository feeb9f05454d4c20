// (1a) Query x reference feature similarity — compute_feature_similarity,
// core/unopose/utils/model_utils.py:260-282:
//     f1 <- F.normalize(f1); f2 <- F.normalize(f2);  atten = f1 @ f2^T / temp        ("cosine")
//     atten = sqrt(clamp(2 - 2 f1 @ f2^T, 0)) / temp                                  ("L2")
// fp32 SIMT path (exact-fp32 reference arithmetic: the reference runs cuBLAS SGEMM with TF32
// off).  128x128x16 tiles, 8x8 register blocking, 128-bit global loads along K (both operands
// are K-contiguous: this is an NT GEMM), transposed smem staging, division by `temp` fused in
// the epilogue.
#include <math.h>

#include "common.cuh"
#include "launch_count.h"
#include "pose_internal.h"
#include "../../include/unopose_b200.h"

namespace upk {

// F.normalize(x, p=2, dim=-1): x / max(||x||_2, 1e-12); one warp per row
__global__ void __launch_bounds__(256)
k_normalize_rows(const float* __restrict__ x, long long rows, int c, float* __restrict__ out) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* p = x + row * c;
  float* o = out + row * c;
  float s = 0.f;
  for (int k = lane; k < c; k += 32) {
    float v = p[k];
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  float nrm = fmaxf(sqrtf(s), 1e-12f);
  for (int k = lane; k < c; k += 32) o[k] = p[k] / nrm;
}

constexpr int BM = 128, BN = 128, BK = 16, GT = 256;

// C[b][i][j] = epilogue( sum_k A[b][i][k] * B[b][j][k] )
template <int MODE>  // 0: dot/temp   1: sqrt(clamp(2-2dot,0))/temp
__global__ void __launch_bounds__(GT)
k_sgemm_nt(const float* __restrict__ A, const float* __restrict__ B, int M, int N, int K, float temp,
           float* __restrict__ C) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  const int bz = blockIdx.z;
  A += (size_t)bz * M * K;
  B += (size_t)bz * N * K;
  C += (size_t)bz * M * N;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8x8 outputs (strided by 16... see below)
  // global->smem mapping: 128 rows x 16 k = 512 float4; 2 per thread
  const int lr = tid >> 2;        // 0..63
  const int lk = (tid & 3) * 4;   // 0,4,8,12
  const bool kvec = (K % 4 == 0);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int buf, int k0) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = lr + h * 64;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
      int gm = m0 + r, gn = n0 + r, gk = k0 + lk;
      if (gm < M) {
        const float* p = A + (size_t)gm * K + gk;
        if (kvec && gk + 3 < K) va = *reinterpret_cast<const float4*>(p);
        else { if (gk < K) va.x = p[0]; if (gk + 1 < K) va.y = p[1]; if (gk + 2 < K) va.z = p[2]; if (gk + 3 < K) va.w = p[3]; }
      }
      if (gn < N) {
        const float* p = B + (size_t)gn * K + gk;
        if (kvec && gk + 3 < K) vb = *reinterpret_cast<const float4*>(p);
        else { if (gk < K) vb.x = p[0]; if (gk + 1 < K) vb.y = p[1]; if (gk + 2 < K) vb.z = p[2]; if (gk + 3 < K) vb.w = p[3]; }
      }
      As[buf][lk + 0][r] = va.x; As[buf][lk + 1][r] = va.y; As[buf][lk + 2][r] = va.z; As[buf][lk + 3][r] = va.w;
      Bs[buf][lk + 0][r] = vb.x; Bs[buf][lk + 1][r] = vb.y; Bs[buf][lk + 2][r] = vb.z; Bs[buf][lk + 3][r] = vb.w;
    }
  };

  const int nk = ceil_div(K, BK);
  load_tile(0, 0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile(buf ^ 1, (kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      // each thread owns rows ty*4..+3 and 64+ty*4..+3, cols tx*4..+3 and 64+tx*4..+3
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (gn >= N) continue;
      float v = acc[i][j];
      if (MODE == 1) v = sqrtf(fmaxf(2.0f - 2.0f * v, 0.f));
      C[(size_t)gm * N + gn] = v / temp;
    }
  }
}

}  // namespace upk

using namespace upk;

extern "C" {

size_t upk_feature_similarity_workspace_bytes(int b, int n, int m, int c, int normalize) {
  if (b > 0 && similarity_tc_eligible(n, m, c)) return similarity_tc_workspace_bytes(b, n, m, c);
  if (!normalize || b <= 0) return 0;
  size_t a = (((size_t)b * n * c * sizeof(float)) + 255) & ~(size_t)255;
  size_t bb = (((size_t)b * m * c * sizeof(float)) + 255) & ~(size_t)255;
  return a + bb;
}

int upk_feature_similarity(const float* feat1, const float* feat2, int b, int n, int m, int c, float temp,
                           int normalize, int sim_type, void* workspace, size_t workspace_bytes,
                           float* atten_out, upk_stream_t stream) {
  if (b < 0 || n <= 0 || m <= 0 || c <= 0) return UPK_ERR_INVALID_ARG;
  if (sim_type != 0 && sim_type != 1) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(feat1) | reinterpret_cast<uintptr_t>(feat2)) & 15) == 0;
  // large problems: tcgen05 tensor-core path (3xTF32); its operand preparation reads the rows with 128-bit loads
  if (similarity_tc_eligible(n, m, c) && aligned16 && workspace &&
      workspace_bytes >= similarity_tc_workspace_bytes(b, n, m, c))
    return run_similarity_tc(feat1, feat2, b, n, m, c, temp, normalize, sim_type, workspace, workspace_bytes,
                             atten_out, st);
  const float* a = feat1;
  const float* bm = feat2;
  if (normalize) {
    size_t need = upk_feature_similarity_workspace_bytes(b, n, m, c, 1);
    if (!workspace || workspace_bytes < need) return UPK_ERR_INVALID_ARG;
    float* an = (float*)workspace;
    float* bn = (float*)((char*)workspace + ((((size_t)b * n * c * sizeof(float)) + 255) & ~(size_t)255));
    long long r1 = (long long)b * n, r2 = (long long)b * m;
    k_normalize_rows<<<(unsigned)((r1 + 7) / 8), 256, 0, st>>>(feat1, r1, c, an);
    k_normalize_rows<<<(unsigned)((r2 + 7) / 8), 256, 0, st>>>(feat2, r2, c, bn);
    count_launch(2);
    a = an;
    bm = bn;
  }
  dim3 grid(ceil_div(m, BN), ceil_div(n, BM), b);
  if (sim_type == 0) k_sgemm_nt<0><<<grid, GT, 0, st>>>(a, bm, n, m, c, temp, atten_out);
  else k_sgemm_nt<1><<<grid, GT, 0, st>>>(a, bm, n, m, c, temp, atten_out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

size_t upk_similarity_stats_bytes(int b, int n, int m) {
  if (b <= 0 || n <= 1 || m <= 1) return 0;
  return sim_stats_geom(b, n, m).total_bytes;
}

int upk_feature_similarity_stats(const float* feat1, const float* feat2, int b, int n, int m, int c, float temp,
                                 void* workspace, size_t workspace_bytes, float* atten_out, float* stats_out,
                                 size_t stats_bytes, upk_stream_t stream) {
  if (b < 0 || n <= 1 || m <= 1 || c <= 0 || !(temp > 0.f)) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(feat1) | reinterpret_cast<uintptr_t>(feat2)) & 15) == 0;
  if (!similarity_tc_eligible(n, m, c) || (similarity_mode() != 3 && similarity_mode() != 16) || !aligned16)
    return UPK_ERR_UNSUPPORTED;
  if (2.0f * 1.4426950408889634f / temp > 60.f) return UPK_ERR_UNSUPPORTED;   // exponent sums would underflow (see _ld)
  const SimStatsGeom sg = sim_stats_geom(b, n, m);
  if (!stats_out || stats_bytes < sg.total_bytes || !workspace ||
      workspace_bytes < similarity_tc_workspace_bytes(b, n, m, c))
    return UPK_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  UPK_CUDA_TRY(cudaMemsetAsync(stats_out, 0, sg.total_bytes, st));   // partials of ragged last tiles stay 0
  return run_similarity_tc(feat1, feat2, b, n, m, c, temp, /*normalize=*/1, /*cosine*/0, workspace, workspace_bytes,
                           atten_out, st, stats_out, stats_out + sg.col_off_floats, sim_stats_gref(temp));
}

int upk_feature_similarity_stats_ld(const float* feat1, const float* feat2, int b, int n, int m, int c, float temp,
                                    void* workspace, size_t workspace_bytes, float* atten_out, int atten_ld,
                                    float* stats_out, size_t stats_bytes, upk_stream_t stream) {
  if (b < 0 || n <= 1 || m <= 1 || c <= 0 || !(temp > 0.f) || atten_ld < m) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  const bool aligned16 = ((reinterpret_cast<uintptr_t>(feat1) | reinterpret_cast<uintptr_t>(feat2)) & 15) == 0;
  if (!similarity_tc_eligible(n, m, c) || (similarity_mode() != 3 && similarity_mode() != 16) || !aligned16)
    return UPK_ERR_UNSUPPORTED;
  if (!atten_out || !workspace || workspace_bytes < similarity_tc_workspace_bytes(b, n, m, c)) return UPK_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  float *srow = nullptr, *scol = nullptr;
  if (stats_out) {
    // one fixed reference exponent carries every exponent sum only while 2^(-2 log2e / temp) stays representable:
    // below temp ~ 0.03 the sums flush to zero -> the caller takes the exact three-pass path
    if (2.0f * 1.4426950408889634f / temp > 60.f) return UPK_ERR_UNSUPPORTED;
    const SimStatsGeom sg = sim_stats_geom(b, n, m);
    if (stats_bytes < sg.total_bytes) return UPK_ERR_INVALID_ARG;
    UPK_CUDA_TRY(cudaMemsetAsync(stats_out, 0, sg.total_bytes, st));   // partials of ragged last tiles stay 0
    srow = stats_out;
    scol = stats_out + sg.col_off_floats;
  }
  return run_similarity_tc(feat1, feat2, b, n, m, c, temp, /*normalize=*/1, /*cosine*/0, workspace, workspace_bytes,
                           atten_out, st, srow, scol, stats_out ? sim_stats_gref(temp) : 0.f, atten_ld);
}

}  // extern "C"
