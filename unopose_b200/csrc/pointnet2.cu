// (5) pointnet2 ops for sm_100a: furthest point sampling, ball query,
// grouping / gathering (+ gradients), three_nn, three_interpolate (+ gradient).
//
// Behavioural contract = the reference extension
//   core/unopose/model/pointnet2/_ext_src/src/{sampling,ball_query,group_points,interpolate}_gpu.cu
// with bit-exact indices (see DESIGN.md §kernels and SURVEY.md Appendix A.1-A.4).
// The launch shapes are NOT the reference's (one CTA per instance): they are
// sized for 148 SMs, coalesced 128-bit accesses and register/smem residency.
#include "common.cuh"
#include "launch_count.h"
#include "../../include/unopose_b200.h"

namespace upk {

// ---------------------------------------------------------------------------
// Furthest point sampling
// ---------------------------------------------------------------------------
// Reference semantics (sampling_gpu.cu:74-178): block of BS = min(2^floor(log2 n), 512)
// threads; thread `t` owns points k = t, t+BS, ... and keeps the FIRST strict
// maximum of d2 = min(d, temp[k]); the smem tree (`__update`, :64-70) keeps the
// LOWER slot on ties.  Unrolling that tournament: among equal d2 the winner is
// the point with the smallest  (bitrev_{log2 BS}(k mod BS), k div BS).
// We therefore reduce on the pair
//     hi = float bits of d2 (d2 >= 0, so unsigned order == float order)   -> max
//     lo = bitrev(k mod BS) << 20 | (k div BS)                             -> min among hi ties
// which reproduces the reference index for every n, including duplicate
// points and the "all distances zero" tail, with any thread count that is a
// multiple of BS (then k mod BS is constant per thread and ascending j is
// ascending `lo`, so a strict `>` scan per thread keeps the right candidate).
//
// Layout: one CTA per instance; each thread keeps its PPT points AND their
// running min-distances in registers for the whole kernel (the reference
// round-trips `temp` through global memory every iteration); the cloud is
// also staged once in smem (SoA) so every thread can fetch the winner's
// coordinates with one broadcast LDS.  One __syncthreads per iteration
// (double-buffered per-warp slots), warp stage = 2x REDUX.

__device__ __forceinline__ unsigned bitrev_n(unsigned v, int nbits) {
  return nbits == 0 ? 0u : (__brev(v) >> (32 - nbits));
}

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(const float* __restrict__ xyz, int n, int m, int bs_log2,
           int* __restrict__ idx_out) {
  constexpr int NW = THREADS / 32;
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sy = sx + n;
  float* sz = sy + n;
  __shared__ uint2 sred[2][32];

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  xyz += (size_t)blockIdx.x * n * 3;
  idx_out += (size_t)blockIdx.x * m;

  // stage the cloud (coalesced AoS read -> SoA smem)
  for (int i = tid; i < n * 3; i += THREADS) {
    float v = xyz[i];
    int k = i / 3, c = i - k * 3;
    (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
  }
  __syncthreads();

  float px[PPT], py[PPT], pz[PPT], td[PPT];
  int cnt = 0;
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    int k = tid + p * THREADS;
    bool ok = k < n;
    px[p] = ok ? sx[k] : 0.f;
    py[p] = ok ? sy[k] : 0.f;
    pz[p] = ok ? sz[k] : 0.f;
    td[p] = 1e10f;  // sampling.cpp:78-80
    cnt += ok ? 1 : 0;
  }
  const unsigned bs_mask = (1u << bs_log2) - 1u;
  const unsigned my_rev = bitrev_n((unsigned)tid & bs_mask, bs_log2) << 20;

  if (tid == 0) idx_out[0] = 0;
  float x1 = sx[0], y1 = sy[0], z1 = sz[0];
  int buf = 0;
  for (int j = 1; j < m; ++j) {
    float best = -1.f;
    int bestp = 0;
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
      if (p < cnt) {
        float d = sqdist_ref(px[p] - x1, py[p] - y1, pz[p] - z1);
        float d2 = fminf(d, td[p]);
        td[p] = d2;
        bool g = d2 > best;
        bestp = g ? p : bestp;
        best = g ? d2 : best;
      }
    }
    unsigned hi = cnt > 0 ? __float_as_uint(best) : 0u;
    unsigned lo = cnt > 0 ? (my_rev | ((unsigned)(tid + bestp * THREADS) >> bs_log2)) : 0xffffffffu;
    unsigned whi = __reduce_max_sync(kFull, hi);
    unsigned wlo = __reduce_min_sync(kFull, hi == whi ? lo : 0xffffffffu);
    if (lane == 0) sred[buf][warp] = make_uint2(whi, wlo);
    __syncthreads();
    uint2 e = lane < NW ? sred[buf][lane] : make_uint2(0u, 0xffffffffu);
    unsigned bhi = __reduce_max_sync(kFull, e.x);
    unsigned blo = __reduce_min_sync(kFull, e.x == bhi ? e.y : 0xffffffffu);
    int old = (int)(((blo & 0xfffffu) << bs_log2) | bitrev_n(blo >> 20, bs_log2));
    x1 = sx[old];
    y1 = sy[old];
    z1 = sz[old];
    if (tid == 0) idx_out[j] = old;
    buf ^= 1;
  }
}

// Large-n fallback: running distances in shared memory, coordinates re-read
// from global (L1/L2).  Same selection rule.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel_large(const float* __restrict__ xyz, int n, int m, int bs_log2,
                 int* __restrict__ idx_out) {
  constexpr int NW = THREADS / 32;
  extern __shared__ float smem_f[];
  float* td = smem_f;  // n
  __shared__ uint2 sred[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  xyz += (size_t)blockIdx.x * n * 3;
  idx_out += (size_t)blockIdx.x * m;
  for (int k = tid; k < n; k += THREADS) td[k] = 1e10f;
  const unsigned bs_mask = (1u << bs_log2) - 1u;
  const unsigned my_rev = bitrev_n((unsigned)tid & bs_mask, bs_log2) << 20;
  if (tid == 0) idx_out[0] = 0;
  float x1 = xyz[0], y1 = xyz[1], z1 = xyz[2];
  int buf = 0;
  for (int j = 1; j < m; ++j) {
    float best = -1.f;
    int bestk = 0;
    for (int k = tid; k < n; k += THREADS) {
      float d = sqdist_ref(xyz[k * 3 + 0] - x1, xyz[k * 3 + 1] - y1, xyz[k * 3 + 2] - z1);
      float d2 = fminf(d, td[k]);
      td[k] = d2;
      bool g = d2 > best;
      bestk = g ? k : bestk;
      best = g ? d2 : best;
    }
    bool has = tid < n;
    unsigned hi = has ? __float_as_uint(best) : 0u;
    unsigned lo = has ? (my_rev | ((unsigned)bestk >> bs_log2)) : 0xffffffffu;
    unsigned whi = __reduce_max_sync(kFull, hi);
    unsigned wlo = __reduce_min_sync(kFull, hi == whi ? lo : 0xffffffffu);
    if (lane == 0) sred[buf][warp] = make_uint2(whi, wlo);
    __syncthreads();
    uint2 e = lane < NW ? sred[buf][lane] : make_uint2(0u, 0xffffffffu);
    unsigned bhi = __reduce_max_sync(kFull, e.x);
    unsigned blo = __reduce_min_sync(kFull, e.x == bhi ? e.y : 0xffffffffu);
    int old = (int)(((blo & 0xfffffu) << bs_log2) | bitrev_n(blo >> 20, bs_log2));
    x1 = xyz[old * 3 + 0];
    y1 = xyz[old * 3 + 1];
    z1 = xyz[old * 3 + 2];
    if (tid == 0) idx_out[j] = old;
    buf ^= 1;
  }
}

template <int THREADS, int PPT>
static int launch_fps(const float* xyz, int b, int n, int m, int bs_log2, int* out,
                      cudaStream_t st) {
  size_t smem = (size_t)n * 3 * sizeof(float);
  auto kern = fps_kernel<THREADS, PPT>;
  if (smem > 40 * 1024) {  // dynamic + the 512 B of static smem must stay within the default 48 KB
    UPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  kern<<<b, THREADS, smem, st>>>(xyz, n, m, bs_log2, out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------
// Ball query
// ---------------------------------------------------------------------------
// Reference (ball_query_gpu.cu:14-49): ONE CTA per instance, a thread walks all
// n points serially for each of its queries.  Here: one warp per query scans
// 32 points per step from an SoA smem tile; __ballot + popc prefix keeps the
// ascending-k slot order; rows are completed (first-hit fill / zeros) by
// coalesced warp stores.  grid = (ceil(m/QPB), b).
constexpr int BQ_WARPS = 8;
constexpr int BQ_QPW = 4;                    // queries per warp
constexpr int BQ_QPB = BQ_WARPS * BQ_QPW;    // queries per block
constexpr int BQ_TILE = 4096;                // points per smem tile (48 KB)

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                  int n, int m, float radius2, int nsample, int* __restrict__ idx) {
  __shared__ float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * m * 3;
  idx += (size_t)b * m * nsample;

  const int q0 = blockIdx.x * BQ_QPB + warp * BQ_QPW;
  float qx[BQ_QPW], qy[BQ_QPW], qz[BQ_QPW];
  int cnt[BQ_QPW], first[BQ_QPW];
#pragma unroll
  for (int q = 0; q < BQ_QPW; ++q) {
    int j = q0 + q;
    bool ok = j < m;
    qx[q] = ok ? new_xyz[j * 3 + 0] : 0.f;
    qy[q] = ok ? new_xyz[j * 3 + 1] : 0.f;
    qz[q] = ok ? new_xyz[j * 3 + 2] : 0.f;
    cnt[q] = ok ? 0 : nsample;  // out-of-range queries are "done"
    first[q] = 0;
  }

  for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
    const int tn = min(BQ_TILE, n - t0);
    __syncthreads();
    for (int i = tid; i < tn * 3; i += BQ_WARPS * 32) {
      float v = xyz[(size_t)t0 * 3 + i];
      int k = i / 3, c = i - k * 3;
      (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BQ_QPW; ++q) {
      if (cnt[q] >= nsample) continue;
      int* row = idx + (size_t)(q0 + q) * nsample;
      for (int base = 0; base < tn; base += 32) {
        int k = base + lane;
        bool hit = false;
        if (k < tn) {
          // (new_x - x)^2 + ... with the reference's FMA contraction
          float d2 = sqdist_ref(qx[q] - sx[k], qy[q] - sy[k], qz[q] - sz[k]);
          hit = d2 < radius2;
        }
        unsigned mask = __ballot_sync(kFull, hit);
        if (mask) {
          if (cnt[q] == 0) first[q] = t0 + base + __ffs(mask) - 1;
          int slot = cnt[q] + __popc(mask & ((1u << lane) - 1u));
          if (hit && slot < nsample) row[slot] = t0 + k;
          cnt[q] += __popc(mask);
          if (cnt[q] >= nsample) break;
        }
      }
    }
  }
  // complete the rows: slots [cnt, nsample) <- first hit (or 0 when no hit)
#pragma unroll
  for (int q = 0; q < BQ_QPW; ++q) {
    int j = q0 + q;
    if (j >= m) continue;
    int* row = idx + (size_t)j * nsample;
    int c = min(cnt[q], nsample);
    int fillv = cnt[q] > 0 ? first[q] : 0;
    for (int s = c + lane; s < nsample; s += 32) row[s] = fillv;
  }
}

// ---------------------------------------------------------------------------
// Grouping / gathering:  out[b,c,l] = points[b,c,idx[b,l]],  l in [0, L)
//   group_points: L = npoints*nsample   gather_points: L = m
// One thread handles 4 consecutive l (one 128-bit idx load, one 128-bit store
// per channel); the idx vector is reused for the CPB channels of the block.
// ---------------------------------------------------------------------------
constexpr int GP_THREADS = 256;
constexpr int GP_CPB = 4;

template <bool VEC>
__global__ void __launch_bounds__(GP_THREADS)
group_kernel(const float* __restrict__ points, const int* __restrict__ idx,
             int c, int n, int L, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * GP_CPB;
  const int cend = min(c0 + GP_CPB, c);
  idx += (size_t)b * L;
  points += (size_t)b * c * n;
  out += (size_t)b * c * L;
  if (VEC) {
    int l = (blockIdx.x * GP_THREADS + threadIdx.x) * 4;
    if (l >= L) return;
    int4 ii = ld_stream_i4(reinterpret_cast<const int4*>(idx + l));
    for (int cc = c0; cc < cend; ++cc) {
      const float* p = points + (size_t)cc * n;
      float4 v = make_float4(__ldg(p + ii.x), __ldg(p + ii.y), __ldg(p + ii.z), __ldg(p + ii.w));
      __stcs(reinterpret_cast<float4*>(out + (size_t)cc * L + l), v);
    }
  } else {
    int l = blockIdx.x * GP_THREADS + threadIdx.x;
    if (l >= L) return;
    int ii = idx[l];
    for (int cc = c0; cc < cend; ++cc)
      out[(size_t)cc * L + l] = __ldg(points + (size_t)cc * n + ii);
  }
}

__global__ void __launch_bounds__(GP_THREADS)
group_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx,
                  int c, int n, int L, float* __restrict__ grad_points) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * GP_CPB;
  const int cend = min(c0 + GP_CPB, c);
  int l = blockIdx.x * GP_THREADS + threadIdx.x;
  if (l >= L) return;
  int ii = idx[(size_t)b * L + l];
  for (int cc = c0; cc < cend; ++cc)
    atomicAdd(grad_points + ((size_t)b * c + cc) * n + ii,
              grad_out[((size_t)b * c + cc) * L + l]);
}

static int launch_group(const float* points, const int* idx, int b, int c, int n, int L,
                        float* out, cudaStream_t st) {
  if (b <= 0 || c <= 0 || L <= 0) return UPK_OK;
  bool vec = (L % 4 == 0) && (((uintptr_t)idx | (uintptr_t)out) % 16 == 0);
  if (vec) {
    dim3 grid(ceil_div(L / 4, GP_THREADS), ceil_div(c, GP_CPB), b);
    group_kernel<true><<<grid, GP_THREADS, 0, st>>>(points, idx, c, n, L, out);
  } else {
    dim3 grid(ceil_div(L, GP_THREADS), ceil_div(c, GP_CPB), b);
    group_kernel<false><<<grid, GP_THREADS, 0, st>>>(points, idx, c, n, L, out);
  }
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

static int launch_group_grad(const float* grad_out, const int* idx, int b, int c, int n,
                             int L, float* grad_points, cudaStream_t st) {
  if (b <= 0 || c <= 0 || n <= 0) return UPK_OK;
  UPK_CUDA_TRY(cudaMemsetAsync(grad_points, 0, (size_t)b * c * n * sizeof(float), st));
  if (L <= 0) return UPK_OK;
  dim3 grid(ceil_div(L, GP_THREADS), ceil_div(c, GP_CPB), b);
  group_grad_kernel<<<grid, GP_THREADS, 0, st>>>(grad_out, idx, c, n, L, grad_points);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

// ---------------------------------------------------------------------------
// three_nn / three_interpolate (not on the UNOPose forward path; kept because
// the reference extension exports them, bindings.cpp:16-18)
// ---------------------------------------------------------------------------
constexpr int NN_THREADS = 128;
constexpr int NN_TILE = 1024;

__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known,
                int n, int m, float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ float sk[NN_TILE * 3];
  const int b = blockIdx.y;
  unknown += (size_t)b * n * 3;
  known += (size_t)b * m * 3;
  dist2 += (size_t)b * n * 3;
  idx += (size_t)b * n * 3;
  const int j = blockIdx.x * NN_THREADS + threadIdx.x;
  const bool ok = j < n;
  float ux = ok ? unknown[j * 3 + 0] : 0.f;
  float uy = ok ? unknown[j * 3 + 1] : 0.f;
  float uz = ok ? unknown[j * 3 + 2] : 0.f;
  // interpolate_gpu.cu:32: the running bests are double, the candidate is float
  double best1 = 1e40, best2 = 1e40, best3 = 1e40;
  int besti1 = 0, besti2 = 0, besti3 = 0;
  for (int t0 = 0; t0 < m; t0 += NN_TILE) {
    int tn = min(NN_TILE, m - t0);
    __syncthreads();
    for (int i = threadIdx.x; i < tn * 3; i += NN_THREADS) sk[i] = known[(size_t)t0 * 3 + i];
    __syncthreads();
    if (ok) {
      for (int k = 0; k < tn; ++k) {
        float d = sqdist_ref(ux - sk[k * 3 + 0], uy - sk[k * 3 + 1], uz - sk[k * 3 + 2]);
        int kk = t0 + k;
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = kk;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = kk;
        } else if (d < best3) {
          best3 = d; besti3 = kk;
        }
      }
    }
  }
  if (ok) {
    dist2[j * 3 + 0] = (float)best1;
    dist2[j * 3 + 1] = (float)best2;
    dist2[j * 3 + 2] = (float)best3;
    idx[j * 3 + 0] = besti1;
    idx[j * 3 + 1] = besti2;
    idx[j * 3 + 2] = besti3;
  }
}

__global__ void __launch_bounds__(256)
three_interpolate_kernel(const float* __restrict__ points, const int* __restrict__ idx,
                         const float* __restrict__ weight, int c, int m, int n,
                         float* __restrict__ out) {
  const int b = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  const float* p = points + ((size_t)b * c + l) * m;
  const int* ix = idx + ((size_t)b * n + j) * 3;
  const float* w = weight + ((size_t)b * n + j) * 3;
  // interpolate_gpu.cu:103-104 under -fmad=true (reference SASS): fma(p3,w3, fma(p1,w1, p2*w2))
  float v = __fmaf_rn(__ldg(p + ix[2]), w[2],
                      __fmaf_rn(__ldg(p + ix[0]), w[0], __fmul_rn(__ldg(p + ix[1]), w[1])));
  out[((size_t)b * c + l) * n + j] = v;
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx,
                              const float* __restrict__ weight, int c, int n, int m,
                              float* __restrict__ grad_points) {
  const int b = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= n) return;
  float g = grad_out[((size_t)b * c + l) * n + j];
  const int* ix = idx + ((size_t)b * n + j) * 3;
  const float* w = weight + ((size_t)b * n + j) * 3;
  float* gp = grad_points + ((size_t)b * c + l) * m;
  atomicAdd(gp + ix[0], g * w[0]);
  atomicAdd(gp + ix[1], g * w[1]);
  atomicAdd(gp + ix[2], g * w[2]);
}

}  // namespace upk

using namespace upk;

extern "C" {

int upk_furthest_point_sampling(const float* xyz, int b, int n, int m, int* idx_out,
                                upk_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || m == 0) return UPK_OK;
  if (n == 0 || !xyz || !idx_out) return UPK_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  // reference block size: opt_n_threads(n) = clamp(2^floor(log2 n), 1, 512)  (cuda_utils.h:20-24)
  int bs_log2 = 0;
  while ((2 << bs_log2) <= n && bs_log2 < 9) ++bs_log2;
  if (((long long)n >> bs_log2) >= (1 << 20)) return UPK_ERR_UNSUPPORTED;
  // THREADS must be a multiple of the reference block size BS = 2^bs_log2 (see fps_kernel)
  if (n <= 128) return launch_fps<128, 1>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n < 256) return launch_fps<128, 2>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n < 512) return launch_fps<256, 2>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 1024) return launch_fps<512, 2>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 2048) return launch_fps<512, 4>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 3072) return launch_fps<1024, 3>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 4096) return launch_fps<1024, 4>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 5120) return launch_fps<1024, 5>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 8192) return launch_fps<512, 16>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 12288) return launch_fps<512, 24>(xyz, b, n, m, bs_log2, idx_out, st);
  size_t smem = (size_t)n * sizeof(float);
  if (smem > 200 * 1024) return UPK_ERR_UNSUPPORTED;
  auto kern = fps_kernel_large<1024>;
  UPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<b, 1024, smem, st>>>(xyz, n, m, bs_log2, idx_out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_gather_points(const float* points, const int* idx, int b, int c, int n, int m,
                      float* out, upk_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0) return UPK_ERR_INVALID_ARG;
  return launch_group(points, idx, b, c, n, m, out, (cudaStream_t)stream);
}

int upk_gather_points_grad(const float* grad_out, const int* idx, int b, int c, int n, int m,
                           float* grad_points, upk_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0) return UPK_ERR_INVALID_ARG;
  return launch_group_grad(grad_out, idx, b, c, n, m, grad_points, (cudaStream_t)stream);
}

int upk_ball_query(const float* new_xyz, const float* xyz, int b, int n, int m, float radius,
                   int nsample, int* idx_out, upk_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || nsample < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || m == 0 || nsample == 0) return UPK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  float r2 = radius * radius;  // fp32 product, ball_query_gpu.cu:27
  dim3 grid(ceil_div(m, BQ_QPB), b);
  ball_query_kernel<<<grid, BQ_WARPS * 32, 0, st>>>(new_xyz, xyz, n, m, r2, nsample, idx_out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_group_points(const float* points, const int* idx, int b, int c, int n, int npoints,
                     int nsample, float* out, upk_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0) return UPK_ERR_INVALID_ARG;
  long long L = (long long)npoints * nsample;
  if (L > 0x7fffffffLL) return UPK_ERR_UNSUPPORTED;
  return launch_group(points, idx, b, c, n, (int)L, out, (cudaStream_t)stream);
}

int upk_group_points_grad(const float* grad_out, const int* idx, int b, int c, int n,
                          int npoints, int nsample, float* grad_points, upk_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0) return UPK_ERR_INVALID_ARG;
  long long L = (long long)npoints * nsample;
  if (L > 0x7fffffffLL) return UPK_ERR_UNSUPPORTED;
  return launch_group_grad(grad_out, idx, b, c, n, (int)L, grad_points, (cudaStream_t)stream);
}

int upk_three_nn(const float* unknown, const float* known, int b, int n, int m,
                 float* dist2_out, int* idx_out, upk_stream_t stream) {
  if (b < 0 || n < 0 || m < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || n == 0) return UPK_OK;
  dim3 grid(ceil_div(n, NN_THREADS), b);
  three_nn_kernel<<<grid, NN_THREADS, 0, (cudaStream_t)stream>>>(unknown, known, n, m,
                                                                 dist2_out, idx_out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_three_interpolate(const float* points, const int* idx, const float* weight, int b,
                          int c, int m, int n, float* out, upk_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || c == 0 || n == 0) return UPK_OK;
  dim3 grid(ceil_div(n, 256), c, b);
  three_interpolate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(points, idx, weight, c, m, n, out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_three_interpolate_grad(const float* grad_out, const int* idx, const float* weight,
                               int b, int c, int n, int m, float* grad_points,
                               upk_stream_t stream) {
  if (b < 0 || c < 0 || n < 0 || m < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || c == 0 || m == 0) return UPK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  UPK_CUDA_TRY(cudaMemsetAsync(grad_points, 0, (size_t)b * c * m * sizeof(float), st));
  if (n == 0) return UPK_OK;
  dim3 grid(ceil_div(n, 256), c, b);
  three_interpolate_grad_kernel<<<grid, 256, 0, st>>>(grad_out, idx, weight, c, n, m, grad_points);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

}  // extern "C"
