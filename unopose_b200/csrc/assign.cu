// (1b) Background-token dual-softmax assignment  (reference: model_utils.py:443-457 coarse,
// :537-553 fine — ~12 ATen launches and ~20 full passes over the (N1+1)x(N2+1)
// matrix, two of them materialised `.repeat` broadcasts).
//
// Here the matrix is read by TILES (TR x TC, staged in shared memory, coalesced
// row segments); every pass produces row-wise AND column-wise partial results
// from the same tile, so each stage reads `atten` exactly once:
//   pass 1  stats  : (max, sum exp) per row and per column         -> merge
//   pass 2  labels : A = softmax_row * softmax_col * s1 * s2; is the background
//                    entry the arg-max of the row / column?          -> w1, w2
//   pass 3c coarse : P = (A w1 w2)^1.5 over the foreground block, row sums (fp64) -> CDF
//   pass 3f fine   : sum_j A_ij w2_j {p2_j, 1}                      -> soft correspondences
// HBM-bound for the fine shape (2049^2 fp32 = 16.8 MB / instance / pass).
#include <math.h>

#include "common.cuh"
#include "launch_count.h"
#include "pose_internal.h"

namespace upk {

constexpr int AT = 256;  // threads per CTA in every tile kernel

// Arithmetic mode per tile shape.  The small tiles (coarse, 197x197: the sampling CDF is built
// from these values and index parity is judged on it) use expf() and IEEE division exactly as
// torch.softmax does.  The large tiles (fine, 2049x2049: HBM-bound, only the R/t tolerance
// applies) use ex2.approx-based __expf and multiply by one reciprocal per row / column: the
// relative error of __expf grows as |x|*2^-24 with x = v - max <= 0, i.e. it is largest on the
// entries whose weight exp(x) is smallest.
template <int TR>
struct TileMath {
  static constexpr bool kFast = TR >= 64;
  static __device__ __forceinline__ float ex(float x) { return kFast ? __expf(x) : expf(x); }
};

AssignGeom assign_geom(int R, int C) {
  AssignGeom g;
  g.R = R;
  g.C = C;
  bool small = (long long)R * C <= 512LL * 512LL;
  g.TR = small ? 32 : 64;
  g.TC = small ? 128 : 256;
  g.ntr = ceil_div(R, g.TR);
  g.ntc = ceil_div(C, g.TC);
  return g;
}

void carve_assign(Carver& cv, int b, const AssignGeom& g, AssignWs& ws) {
  ws.rowpart = cv.take<float2>((size_t)b * g.R * g.ntc);
  ws.colpart = cv.take<float2>((size_t)b * g.C * g.ntr);
  ws.rmax = cv.take<float>((size_t)b * g.R);
  ws.rsum = cv.take<float>((size_t)b * g.R);
  ws.cmax = cv.take<float>((size_t)b * g.C);
  ws.csum = cv.take<float>((size_t)b * g.C);
  ws.rowpm = cv.take<float>((size_t)b * g.R * g.ntc);
  ws.colpm = cv.take<float>((size_t)b * g.C * g.ntr);
  ws.ai0 = cv.take<float>((size_t)b * g.R);
  ws.a0j = cv.take<float>((size_t)b * g.C);
  ws.flags = cv.take<int>((size_t)b);
  ws.bsum = cv.take<float>((size_t)b * 2);
}

template <int TR, int TC>
__device__ __forceinline__ void load_tile(const float* __restrict__ A, int C, int r0, int c0, int nr,
                                          int nc, float* tile) {
#pragma unroll 8
  for (int i = threadIdx.x; i < TR * TC; i += AT) {
    int r = i / TC, c = i % TC;
    float v = -INFINITY;
    if (r < nr && c < nc) v = __ldg(A + (size_t)(r0 + r) * C + c0 + c);
    tile[i] = v;
  }
}

// ------------------------------------------------------------------ pass 1
template <int TR, int TC>
__global__ void __launch_bounds__(AT)
k_stats_tile(const float* __restrict__ atten, int R, int C, int ntr, int ntc,
             float2* __restrict__ rowpart, float2* __restrict__ colpart) {
  extern __shared__ float tile[];
  const int b = blockIdx.z, tr = blockIdx.y, tc = blockIdx.x;
  const int r0 = tr * TR, c0 = tc * TC;
  const int nr = min(TR, R - r0), nc = min(TC, C - c0);
  load_tile<TR, TC>(atten + (size_t)b * R * C, C, r0, c0, nr, nc, tile);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < nr; r += AT / 32) {
    const float* row = tile + r * TC;
    float m = -INFINITY;
    for (int c = lane; c < nc; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < nc; c += 32) s += TileMath<TR>::ex(row[c] - m);
    s = warp_sum(s);
    if (lane == 0) rowpart[((size_t)b * R + r0 + r) * ntc + tc] = make_float2(m, s);
  }
  for (int c = threadIdx.x; c < nc; c += AT) {
    float m = -INFINITY;
    for (int r = 0; r < nr; ++r) m = fmaxf(m, tile[r * TC + c]);
    float s = 0.f;
    for (int r = 0; r < nr; ++r) s += TileMath<TR>::ex(tile[r * TC + c] - m);
    colpart[((size_t)b * C + c0 + c) * ntr + tr] = make_float2(m, s);
  }
}

__global__ void __launch_bounds__(256)
k_stats_merge(const float2* __restrict__ rowpart, const float2* __restrict__ colpart, int R, int C,
              int ntr, int ntc, float* __restrict__ rmax, float* __restrict__ rsum,
              float* __restrict__ cmax, float* __restrict__ csum) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float2* p;
  int n;
  float *om, *os;
  if (i < R) {
    p = rowpart + ((size_t)b * R + i) * ntc; n = ntc; om = rmax + (size_t)b * R + i; os = rsum + (size_t)b * R + i;
  } else if (i < R + C) {
    int c = i - R;
    p = colpart + ((size_t)b * C + c) * ntr; n = ntr; om = cmax + (size_t)b * C + c; os = csum + (size_t)b * C + c;
  } else {
    return;
  }
  float m = -INFINITY;
  for (int k = 0; k < n; ++k) m = fmaxf(m, p[k].x);
  float s = 0.f;
  for (int k = 0; k < n; ++k) s += p[k].y * expf(p[k].x - m);
  *om = m;
  *os = s;
}

// Row/column constants of a tile in shared memory + in-place A computation.
template <int TR, int TC>
struct TileConsts {
  float rm[TR], rs[TR], s1[TR];
  float cm[TC], cs[TC], s2[TC];
};

template <int TR, int TC>
__device__ __forceinline__ void load_consts(TileConsts<TR, TC>& k, int b, int R, int C, int r0, int c0,
                                            int nr, int nc, const float* rmax, const float* rsum,
                                            const float* cmax, const float* csum, const float* score1,
                                            int ld1, const float* score2, int ld2) {
  for (int r = threadIdx.x; r < TR; r += AT) {
    int gi = r0 + r;
    bool ok = r < nr;
    k.rm[r] = ok ? rmax[(size_t)b * R + gi] : 0.f;
    float rs = ok ? rsum[(size_t)b * R + gi] : 1.f;
    k.rs[r] = TileMath<TR>::kFast ? 1.f / rs : rs;  // fast mode keeps the reciprocal
    // background row/col carry score 1.0 (model_utils.py:443-445)
    k.s1[r] = (ok && gi > 0 && score1) ? score1[(size_t)b * ld1 + gi - 1] : 1.f;
  }
  for (int c = threadIdx.x; c < TC; c += AT) {
    int gj = c0 + c;
    bool ok = c < nc;
    k.cm[c] = ok ? cmax[(size_t)b * C + gj] : 0.f;
    float cs = ok ? csum[(size_t)b * C + gj] : 1.f;
    k.cs[c] = TileMath<TR>::kFast ? 1.f / cs : cs;
    k.s2[c] = (ok && gj > 0 && score2) ? score2[(size_t)b * ld2 + gj - 1] : 1.f;
  }
}

// A = softmax(atten,2) * softmax(atten,1) * score1 * score2   (model_utils.py:448-449 / :542-543)
template <int TR, int TC>
__device__ __forceinline__ void compute_A_inplace(float* tile, const TileConsts<TR, TC>& k, int nr, int nc) {
  for (int i = threadIdx.x; i < nr * TC; i += AT) {
    int r = i / TC, c = i % TC;
    float a = 0.f;
    if (c < nc) {
      float v = tile[i];
      float er, ec;
      if (TileMath<TR>::kFast) {
        er = __expf(v - k.rm[r]) * k.rs[r];
        ec = __expf(v - k.cm[c]) * k.cs[c];
      } else {
        er = expf(v - k.rm[r]) / k.rs[r];
        ec = expf(v - k.cm[c]) / k.cs[c];
      }
      a = ((er * ec) * k.s1[r]) * k.s2[c];
    }
    tile[i] = a;
  }
}

// ------------------------------------------------------------------ pass 2
template <int TR, int TC>
__global__ void __launch_bounds__(AT)
k_labels_tile(const float* __restrict__ atten, int R, int C, int ntr, int ntc,
              const float* __restrict__ rmax, const float* __restrict__ rsum,
              const float* __restrict__ cmax, const float* __restrict__ csum,
              const float* __restrict__ score1, int ld1, const float* __restrict__ score2, int ld2,
              float* __restrict__ rowpm, float* __restrict__ colpm, float* __restrict__ ai0,
              float* __restrict__ a0j) {
  extern __shared__ float tile[];
  __shared__ TileConsts<TR, TC> k;
  const int b = blockIdx.z, tr = blockIdx.y, tc = blockIdx.x;
  const int r0 = tr * TR, c0 = tc * TC;
  const int nr = min(TR, R - r0), nc = min(TC, C - c0);
  load_tile<TR, TC>(atten + (size_t)b * R * C, C, r0, c0, nr, nc, tile);
  load_consts<TR, TC>(k, b, R, C, r0, c0, nr, nc, rmax, rsum, cmax, csum, score1, ld1, score2, ld2);
  __syncthreads();
  compute_A_inplace<TR, TC>(tile, k, nr, nc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cstart = c0 == 0 ? 1 : 0;  // exclude the background column from the row max
  for (int r = warp; r < nr; r += AT / 32) {
    const float* row = tile + r * TC;
    float m = -INFINITY;
    for (int c = cstart + lane; c < nc; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    if (lane == 0) {
      rowpm[((size_t)b * R + r0 + r) * ntc + tc] = m;
      if (c0 == 0) ai0[(size_t)b * R + r0 + r] = row[0];
    }
  }
  const int rstart = r0 == 0 ? 1 : 0;
  for (int c = threadIdx.x; c < nc; c += AT) {
    float m = -INFINITY;
    for (int r = rstart; r < nr; ++r) m = fmaxf(m, tile[r * TC + c]);
    colpm[((size_t)b * C + c0 + c) * ntr + tr] = m;
    if (r0 == 0) a0j[(size_t)b * C + c0 + c] = tile[c];
  }
}

// w1[i-1] = (argmax_j A[i][:] > 0)  <=>  max_{j>=1} A[i][j] > A[i][0]  (torch.max returns the first maximum)
__global__ void __launch_bounds__(256)
k_labels_merge(const float* __restrict__ rowpm, const float* __restrict__ colpm,
               const float* __restrict__ ai0, const float* __restrict__ a0j, int R, int C, int ntr,
               int ntc, float* __restrict__ w1, float* __restrict__ w2) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 1 && i < R) {
    const float* p = rowpm + ((size_t)b * R + i) * ntc;
    float m = -INFINITY;
    for (int k = 0; k < ntc; ++k) m = fmaxf(m, p[k]);
    w1[(size_t)b * (R - 1) + i - 1] = m > ai0[(size_t)b * R + i] ? 1.f : 0.f;
  } else if (i >= R + 1 && i < R + C) {
    int c = i - R;
    const float* p = colpm + ((size_t)b * C + c) * ntr;
    float m = -INFINITY;
    for (int k = 0; k < ntr; ++k) m = fmaxf(m, p[k]);
    w2[(size_t)b * (C - 1) + c - 1] = m > a0j[(size_t)b * C + c] ? 1.f : 0.f;
  }
}

// ------------------------------------------------------------------ pass 3 (coarse)
// P = (A[1:,1:] * w1 (x) w2) ** 1.5   (model_utils.py:455-457); x^1.5 evaluated as x*sqrt(x)
template <int TR, int TC>
__global__ void __launch_bounds__(AT)
k_coarse_P_tile(const float* __restrict__ atten, int R, int C, int ntc,
                const float* __restrict__ rmax, const float* __restrict__ rsum,
                const float* __restrict__ cmax, const float* __restrict__ csum,
                const float* __restrict__ score1, int ld1, const float* __restrict__ score2, int ld2,
                const float* __restrict__ w1, const float* __restrict__ w2, float* __restrict__ pmat,
                double* __restrict__ prow) {
  extern __shared__ float tile[];
  __shared__ TileConsts<TR, TC> k;
  const int b = blockIdx.z, tr = blockIdx.y, tc = blockIdx.x;
  const int r0 = tr * TR, c0 = tc * TC;
  const int nr = min(TR, R - r0), nc = min(TC, C - c0);
  const int N1 = R - 1, N2 = C - 1;
  load_tile<TR, TC>(atten + (size_t)b * R * C, C, r0, c0, nr, nc, tile);
  load_consts<TR, TC>(k, b, R, C, r0, c0, nr, nc, rmax, rsum, cmax, csum, score1, ld1, score2, ld2);
  __syncthreads();
  compute_A_inplace<TR, TC>(tile, k, nr, nc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < nr; r += AT / 32) {
    int gi = r0 + r;
    if (gi == 0) continue;
    float wr = w1[(size_t)b * N1 + gi - 1];
    double acc = 0.0;
    for (int c = lane; c < nc; c += 32) {
      int gj = c0 + c;
      if (gj == 0) continue;
      float a = (tile[r * TC + c] * wr) * w2[(size_t)b * N2 + gj - 1];
      float p = a * sqrtf(a);
      pmat[((size_t)b * N1 + gi - 1) * N2 + gj - 1] = p;
      acc += (double)p;
    }
    acc = warp_sum(acc);
    if (lane == 0) prow[((size_t)b * N1 + gi - 1) * ntc + tc] = acc;
  }
}

// cdf = cumsum(P) / (cumsum(P)[-1] + 1e-8)   (model_utils.py:460-461).  fp64 running sums
// (CPU torch accumulates cumsum in double as well), rounded to fp32 per element.
constexpr int CDF_ROWS = 8;
__global__ void __launch_bounds__(CDF_ROWS * 32)
k_cdf(const float* __restrict__ pmat, const double* __restrict__ prow, int n1, int n2, int ntc,
      float* __restrict__ cdf) {
  extern __shared__ double rowtot[];  // n1
  __shared__ double s_warp[CDF_ROWS];
  __shared__ double s_offset0, s_total;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * CDF_ROWS;
  // row totals (identical in every CTA: fixed summation order)
  for (int i = threadIdx.x; i < n1; i += blockDim.x) {
    const double* p = prow + ((size_t)b * n1 + i) * ntc;
    double s = 0.0;
    for (int k = 0; k < ntc; ++k) s += p[k];
    rowtot[i] = s;
  }
  __syncthreads();
  // canonical sequential prefix (one thread; n1 <= a few thousand): offset of row0 and the total
  if (threadIdx.x == 0) {
    double acc = 0.0, off0 = 0.0;
    for (int i = 0; i < n1; ++i) {
      if (i == row0) off0 = acc;
      acc += rowtot[i];
    }
    s_offset0 = off0;
    s_total = acc;
  }
  __syncthreads();
  const float denom = (float)s_total + 1e-8f;
  // offsets of this CTA's rows: sequential from s_offset0
  if (threadIdx.x == 0) {
    double acc = s_offset0;
    for (int w = 0; w < CDF_ROWS; ++w) {
      s_warp[w] = acc;
      if (row0 + w < n1) acc += rowtot[row0 + w];
    }
  }
  __syncthreads();
  const int i = row0 + warp;
  if (i >= n1) return;
  double carry = s_warp[warp];
  const float* prow_i = pmat + ((size_t)b * n1 + i) * n2;
  float* out = cdf + ((size_t)b * n1 + i) * n2;
  for (int c0 = 0; c0 < n2; c0 += 32) {
    int c = c0 + lane;
    double v = c < n2 ? (double)prow_i[c] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      double t = __shfl_up_sync(kFull, v, o);
      if (lane >= o) v += t;
    }
    if (c < n2) out[c] = __fdiv_rn((float)(carry + v), denom);
    carry += __shfl_sync(kFull, v, 31);
  }
}

// ------------------------------------------------------------------ pass 3 (fine)
// per row i>=1:  sum_{j>=1} A_ij w2_j {x_j, y_j, z_j, 1}   (model_utils.py:549-555)
template <int TR, int TC>
__global__ void __launch_bounds__(AT)
k_fine_rows_tile(const float* __restrict__ atten, int R, int C, int ntc,
                 const float* __restrict__ rmax, const float* __restrict__ rsum,
                 const float* __restrict__ cmax, const float* __restrict__ csum,
                 const float* __restrict__ score1, int ld1, const float* __restrict__ score2, int ld2,
                 const float* __restrict__ w2, const float* __restrict__ pts2,
                 float4* __restrict__ rowpart4) {
  extern __shared__ float tile[];
  __shared__ TileConsts<TR, TC> k;
  __shared__ float4 colp[TC];  // (x,y,z,w2) of the tile's columns
  const int b = blockIdx.z, tr = blockIdx.y, tc = blockIdx.x;
  const int r0 = tr * TR, c0 = tc * TC;
  const int nr = min(TR, R - r0), nc = min(TC, C - c0);
  const int N1 = R - 1, N2 = C - 1;
  load_tile<TR, TC>(atten + (size_t)b * R * C, C, r0, c0, nr, nc, tile);
  load_consts<TR, TC>(k, b, R, C, r0, c0, nr, nc, rmax, rsum, cmax, csum, score1, ld1, score2, ld2);
  for (int c = threadIdx.x; c < TC; c += AT) {
    int gj = c0 + c;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < nc && gj > 0) {
      const float* p = pts2 + ((size_t)b * N2 + gj - 1) * 3;
      q = make_float4(p[0], p[1], p[2], w2[(size_t)b * N2 + gj - 1]);
    }
    colp[c] = q;
  }
  __syncthreads();
  compute_A_inplace<TR, TC>(tile, k, nr, nc);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < nr; r += AT / 32) {
    int gi = r0 + r;
    if (gi == 0) continue;
    float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
    for (int c = lane; c < nc; c += 32) {
      float4 q = colp[c];
      float a = tile[r * TC + c] * q.w;
      sx = fmaf(a, q.x, sx);
      sy = fmaf(a, q.y, sy);
      sz = fmaf(a, q.z, sz);
      sw += a;
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); sw = warp_sum(sw);
    if (lane == 0) rowpart4[((size_t)b * N1 + gi - 1) * ntc + tc] = make_float4(sx, sy, sz, sw);
  }
}

// asum_i = w1_i * sum_j A_ij w2_j ;  soft_i = (w1_i * sum_j A_ij w2_j p2_j) / (asum_i + 1e-6)
__global__ void __launch_bounds__(256)
k_fine_rows_merge(const float4* __restrict__ rowpart4, const float* __restrict__ w1, int n1, int ntc,
                  float* __restrict__ soft, float* __restrict__ asum) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n1) return;
  const float4* p = rowpart4 + ((size_t)b * n1 + i) * ntc;
  float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
  for (int k = 0; k < ntc; ++k) {
    float4 q = p[k];
    sx += q.x; sy += q.y; sz += q.z; sw += q.w;
  }
  float w = w1[(size_t)b * n1 + i];
  sx *= w; sy *= w; sz *= w; sw *= w;
  float d = sw + 1e-6f;
  float* o = soft + ((size_t)b * n1 + i) * 3;
  o[0] = sx / d; o[1] = sy / d; o[2] = sz / d;
  asum[(size_t)b * n1 + i] = sw;
}

// ------------------------------------------------------------------ host launchers
template <int TR, int TC>
static int stats_labels_t(const float* atten, const float* score1, int ld1, const float* score2, int ld2,
                          int b, const AssignGeom& g, const AssignWs& ws, float* w1, float* w2,
                          cudaStream_t st) {
  const size_t smem = (size_t)TR * TC * sizeof(float);
  auto k1 = k_stats_tile<TR, TC>;
  auto k2 = k_labels_tile<TR, TC>;
  if (smem + sizeof(TileConsts<TR, TC>) > 40 * 1024) {
    UPK_CUDA_TRY(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    UPK_CUDA_TRY(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  dim3 grid(g.ntc, g.ntr, b);
  k1<<<grid, AT, smem, st>>>(atten, g.R, g.C, g.ntr, g.ntc, ws.rowpart, ws.colpart);
  dim3 mg(ceil_div(g.R + g.C, 256), b);
  k_stats_merge<<<mg, 256, 0, st>>>(ws.rowpart, ws.colpart, g.R, g.C, g.ntr, g.ntc, ws.rmax, ws.rsum,
                                    ws.cmax, ws.csum);
  k2<<<grid, AT, smem, st>>>(atten, g.R, g.C, g.ntr, g.ntc, ws.rmax, ws.rsum, ws.cmax, ws.csum, score1,
                             ld1, score2, ld2, ws.rowpm, ws.colpm, ws.ai0, ws.a0j);
  k_labels_merge<<<mg, 256, 0, st>>>(ws.rowpm, ws.colpm, ws.ai0, ws.a0j, g.R, g.C, g.ntr, g.ntc, w1, w2);
  count_launch(4);
  UPK_RETURN_LAST_ERROR();
}

int run_assignment_labels(const float* atten, const float* score1, int ld1, const float* score2, int ld2,
                          int b, const AssignGeom& g, const AssignWs& ws, float* w1, float* w2,
                          cudaStream_t st, bool for_fine_solve, int atten_ld) {
  const int ld = atten_ld > 0 ? atten_ld : g.C;
  // large geometry, fine solve: the streaming passes of assign_fine.cu (the only ones that read a pitched atten)
  if (g.TR != 32 && for_fine_solve) return run_fine_labels2(atten, ld, score1, ld1, score2, ld2, b, g, ws, w1, w2, st);
  if (ld != g.C) return UPK_ERR_UNSUPPORTED;
  if (g.TR == 32) return stats_labels_t<32, 128>(atten, score1, ld1, score2, ld2, b, g, ws, w1, w2, st);
  // large geometry, coarse solve (never the case at the UNOPose shapes): the exact tile pipeline at 64 x 256
  return stats_labels_t<64, 256>(atten, score1, ld1, score2, ld2, b, g, ws, w1, w2, st);
}

int launch_labels_merge(const float* rowpm, const float* colpm, const float* ai0, const float* a0j, int b, int R,
                        int C, int ntr, int ntc, float* w1, float* w2, cudaStream_t st) {
  dim3 mg(ceil_div(R + C, 256), b);
  k_labels_merge<<<mg, 256, 0, st>>>(rowpm, colpm, ai0, a0j, R, C, ntr, ntc, w1, w2);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int launch_fine_rows_merge(const float4* rowpart4, const float* w1, int b, int n1, int ntc, float* soft,
                           float* asum, cudaStream_t st) {
  dim3 mg(ceil_div(n1, 256), b);
  k_fine_rows_merge<<<mg, 256, 0, st>>>(rowpart4, w1, n1, ntc, soft, asum);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

template <int TR, int TC>
static int coarse_P_t(const float* atten, const float* score1, int ld1, const float* score2, int ld2, int b,
                      const AssignGeom& g, const AssignWs& ws, const float* w1, const float* w2,
                      float* pmat, double* prow, cudaStream_t st) {
  const size_t smem = (size_t)TR * TC * sizeof(float);
  auto k = k_coarse_P_tile<TR, TC>;
  if (smem + sizeof(TileConsts<TR, TC>) > 40 * 1024)
    UPK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(g.ntc, g.ntr, b);
  k<<<grid, AT, smem, st>>>(atten, g.R, g.C, g.ntc, ws.rmax, ws.rsum, ws.cmax, ws.csum, score1, ld1,
                            score2, ld2, w1, w2, pmat, prow);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int run_coarse_P(const float* atten, const float* score1, int ld1, const float* score2, int ld2, int b,
                 const AssignGeom& g, const AssignWs& ws, const float* w1, const float* w2, float* pmat,
                 double* prow, cudaStream_t st) {
  if (g.TR == 32) return coarse_P_t<32, 128>(atten, score1, ld1, score2, ld2, b, g, ws, w1, w2, pmat, prow, st);
  return coarse_P_t<64, 256>(atten, score1, ld1, score2, ld2, b, g, ws, w1, w2, pmat, prow, st);
}

int run_cdf(const float* pmat, const double* prow, int b, int n1, int n2, int ntc, float* cdf,
            cudaStream_t st) {
  size_t smem = (size_t)n1 * sizeof(double);
  if (smem > 40 * 1024)
    UPK_CUDA_TRY(cudaFuncSetAttribute(k_cdf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(n1, CDF_ROWS), b);
  k_cdf<<<grid, CDF_ROWS * 32, smem, st>>>(pmat, prow, n1, n2, ntc, cdf);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

template <int TR, int TC>
static int fine_rows_t(const float* atten, const float* score1, int ld1, const float* score2, int ld2,
                       int b, const AssignGeom& g, const AssignWs& ws, const float* w1, const float* w2,
                       const float* pts2, float4* rowpart4, float* soft, float* asum, cudaStream_t st) {
  const size_t smem = (size_t)TR * TC * sizeof(float);
  auto k = k_fine_rows_tile<TR, TC>;
  if (smem + sizeof(TileConsts<TR, TC>) + TC * 16 > 40 * 1024)
    UPK_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(g.ntc, g.ntr, b);
  k<<<grid, AT, smem, st>>>(atten, g.R, g.C, g.ntc, ws.rmax, ws.rsum, ws.cmax, ws.csum, score1, ld1,
                            score2, ld2, w2, pts2, rowpart4);
  int n1 = g.R - 1;
  dim3 mg(ceil_div(n1, 256), b);
  k_fine_rows_merge<<<mg, 256, 0, st>>>(rowpart4, w1, n1, g.ntc, soft, asum);
  count_launch(2);
  UPK_RETURN_LAST_ERROR();
}

int run_fine_rowsums(const float* atten, const float* score1, int ld1, const float* score2, int ld2, int b,
                     const AssignGeom& g, const AssignWs& ws, const float* w1, const float* w2,
                     const float* pts2, float4* rowpart4, float* soft, float* asum, cudaStream_t st, int atten_ld) {
  const int ld = atten_ld > 0 ? atten_ld : g.C;
  if (g.TR == 32) {
    if (ld != g.C) return UPK_ERR_UNSUPPORTED;
    return fine_rows_t<32, 128>(atten, score1, ld1, score2, ld2, b, g, ws, w1, w2, pts2, rowpart4, soft, asum, st);
  }
  return run_fine_rows2(atten, ld, b, g, ws, w1, w2, pts2, rowpart4, soft, asum, st);
}

}  // namespace upk
