// (5) pointnet2 ops for sm_100a: furthest point sampling, ball query,
// grouping / gathering (+ gradients), three_nn, three_interpolate (+ gradient).
//
// Behavioural contract = the reference extension
//   core/unopose/model/pointnet2/_ext_src/src/{sampling,ball_query,group_points,interpolate}_gpu.cu
// with bit-exact indices (see DESIGN.md §kernels and SURVEY.md Appendix A.1-A.4).
// The launch shapes are NOT the reference's (one CTA per instance): they are
// sized for 148 SMs, coalesced 128-bit accesses and register/smem residency.
#include <stdlib.h>

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

// Offset of a thread's p-th point from its thread index.  Default: k = tid + p * THREADS; with THREADS a multiple of the
// reference block size BS every point of a thread has the same k mod BS, so ascending p is ascending tie key and the
// strict ">" of the scan keeps the reference's winner among equal distances.  GENKEY (THREADS = BS / S, BS = 512): a thread
// owns S residues r_s = tid + s * THREADS; the key orders them by bitrev(r_s), i.e. by the bit-reversed s, so the points
// are laid out residue by residue in THAT order (J = PPT / S blocks each) and ascending p is again ascending key.
template <int THREADS, int PPT, bool GENKEY>
__host__ __device__ constexpr int fps_koff(int p) {
  if (!GENKEY) return p * THREADS;
  constexpr int S = GENKEY ? 512 / THREADS : 1, J = PPT / S;
  const int si = p / J, j = p % J;
  int s = 0;
  for (int bit = 1, rb = S >> 1; bit < S; bit <<= 1, rb >>= 1)
    if (si & bit) s |= rb;
  return s * THREADS + j * 512;
}

template <int THREADS, int PPT, bool X2 = false, bool GENKEY = false>
__global__ void __launch_bounds__(THREADS, 1)
fps_kernel(const float* __restrict__ xyz, int n, int m, int bs_log2,
           int* __restrict__ idx_out) {
  static_assert(!X2 || PPT % 2 == 0, "packed variant needs an even number of points per thread");
  static_assert(!GENKEY || (512 % THREADS == 0 && PPT % (512 / THREADS) == 0), "GENKEY: BS = 512 split over S residues per thread");
  extern __shared__ float smem_f[];
  float* sx = smem_f;
  float* sy = sx + n;
  float* sz = sy + n;
  constexpr int NW = THREADS / 32;
  __shared__ int2 sred[2][32];  // per-warp (hi, lo) candidates, double-buffered -> one barrier / iteration

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  xyz += (size_t)blockIdx.x * n * 3;
  idx_out += (size_t)blockIdx.x * m;

  // stage the cloud (coalesced AoS read -> SoA smem)
  for (int i = tid; i < n * 3; i += THREADS) {
    float v = xyz[i];
    int k = i / 3, c = i - k * 3;
    (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
  }
  __syncthreads();

  // Points past the end of the cloud carry a running distance of -1: min(d, -1) = -1 never beats
  // `best` (initialised to -1, strict >), so the scan below needs no bounds checks.
  float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    int k = tid + fps_koff<THREADS, PPT, GENKEY>(p);
    bool ok = k < n;
    px[p] = ok ? sx[k] : 0.f;
    py[p] = ok ? sy[k] : 0.f;
    pz[p] = ok ? sz[k] : 0.f;
    td[p] = ok ? 1e10f : -1.f;  // sampling.cpp:78-80
  }
  const unsigned bs_mask = (1u << bs_log2) - 1u;
  const unsigned my_rev = bitrev_n((unsigned)tid & bs_mask, bs_log2) << 20;

  if (tid == 0) idx_out[0] = 0;
  float x1 = sx[0], y1 = sy[0], z1 = sz[0];
  const int warp = tid >> 5;
  int buf = 0;
  for (int j = 1; j < m; ++j) {
    float best = -1.f;
    int bestp = 0;
    if (X2) {
      // (x2 - x1) == x2 + (-x1) exactly; d = fma(dz,dz, fma(dx,dx, dy*dy)) on two points per instruction
      const unsigned long long nx = pack2(-x1, -x1), ny = pack2(-y1, -y1), nz = pack2(-z1, -z1);
#pragma unroll
      for (int p = 0; p < PPT; p += 2) {
        unsigned long long dx = add2(pack2(px[p], px[p + 1]), nx);
        unsigned long long dy = add2(pack2(py[p], py[p + 1]), ny);
        unsigned long long dz = add2(pack2(pz[p], pz[p + 1]), nz);
        unsigned long long dd = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
        float da, db;
        unpack2(dd, da, db);
        float d2a = fminf(da, td[p]), d2b = fminf(db, td[p + 1]);
        td[p] = d2a;
        td[p + 1] = d2b;
        bool ga = d2a > best;
        bestp = ga ? fps_koff<THREADS, PPT, GENKEY>(p) : bestp;
        best = ga ? d2a : best;
        bool gb = d2b > best;
        bestp = gb ? fps_koff<THREADS, PPT, GENKEY>(p + 1) : bestp;
        best = gb ? d2b : best;
      }
    } else {
#pragma unroll
      for (int p = 0; p < PPT; ++p) {
        float d = sqdist_ref(px[p] - x1, py[p] - y1, pz[p] - z1);
        float d2 = fminf(d, td[p]);
        td[p] = d2;
        bool g = d2 > best;
        bestp = g ? fps_koff<THREADS, PPT, GENKEY>(p) : bestp;
        best = g ? d2 : best;
      }
    }
    // d2 >= 0 for real points, so signed-int order of the bit patterns == float order and the
    // sentinel -1.0f (negative as int) loses against everything
    const int hi = __float_as_int(best);
    // tie key of point k: (bitrev(k mod BS), k div BS).  With THREADS a multiple of BS, k mod BS == tid mod BS for all of
    // a thread's points; the configurations with fewer warps than BS / 32 (GENKEY) evaluate it for the winning point
    const unsigned kb = (unsigned)(tid + bestp);   // bestp = offset of the thread's winning point (fps_koff)
    const unsigned lo = (GENKEY ? bitrev_n(kb & bs_mask, bs_log2) << 20 : my_rev) | (kb >> bs_log2);
    const int whi = __reduce_max_sync(kFull, hi);
    const unsigned wlo = __reduce_min_sync(kFull, hi == whi ? lo : 0xffffffffu);
    if (lane == 0) sred[buf][warp] = make_int2(whi, (int)wlo);
    __syncthreads();
    const int2 e = lane < NW ? sred[buf][lane] : make_int2(-1, -1);
    const int bhi = __reduce_max_sync(kFull, e.x);
    const unsigned blo = __reduce_min_sync(kFull, e.x == bhi ? (unsigned)e.y : 0xffffffffu);
    const int old = (int)(((blo & 0xfffffu) << bs_log2) | bitrev_n(blo >> 20, bs_log2));
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

// UPK_FPS_CFG (dev knob): 1 = 1024-thread variants, 2 = 512-thread scalar, 4 = 512-thread packed f32x2;
// default = tuned choice
static int fps_cfg() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UPK_FPS_CFG");
    v = e ? atoi(e) : 0;
  }
  return v;
}

static int fps_exclusive() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("UPK_FPS_EXCLUSIVE");
    v = e ? atoi(e) : 0;  // measured: no gain on the step median, kept as an opt-in knob
  }
  return v;
}

template <int THREADS, int PPT, bool X2 = false, bool GENKEY = false>
static int launch_fps(const float* xyz, int b, int n, int m, int bs_log2, int* out,
                      cudaStream_t st) {
  size_t smem = (size_t)n * 3 * sizeof(float);
  // A long FPS is a serial latency chain: when it overlaps other kernels on a second stream, CTAs that
  // co-reside on its SM steal issue slots and stretch every one of its m-1 iterations.  Claiming most of
  // the SM's shared memory keeps the SM exclusive (opt-in: UPK_FPS_EXCLUSIVE=1).
  if (fps_exclusive() && (long long)n * m >= 4096LL * 1024LL && smem < 200 * 1024) smem = 200 * 1024;
  auto kern = fps_kernel<THREADS, PPT, X2, GENKEY>;
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
constexpr int BQ_TILE = 3968;                // points per smem tile (46.5 KB), multiple of 128

// One ballot step over 64 points (two per lane, packed f32x2 arithmetic): lane l tests points
// k0 + 2l and k0 + 2l + 1.  Returns the two hit masks (bit l of `even`/`odd` = lane l's first/second point).
__device__ __forceinline__ void bq_hits2(const float* sx, const float* sy, const float* sz, int k,
                                         unsigned long long nqx, unsigned long long nqy, unsigned long long nqz,
                                         float r2, unsigned& even, unsigned& odd) {
  // (new_x - x)^2 == (x - new_x)^2 exactly; d2 = fma(dz,dz, fma(dx,dx, dy*dy)) as the reference's SASS
  const float2 x = *reinterpret_cast<const float2*>(sx + k);
  const float2 y = *reinterpret_cast<const float2*>(sy + k);
  const float2 z = *reinterpret_cast<const float2*>(sz + k);
  unsigned long long dx = add2(pack2(x.x, x.y), nqx);
  unsigned long long dy = add2(pack2(y.x, y.y), nqy);
  unsigned long long dz = add2(pack2(z.x, z.y), nqz);
  unsigned long long dd = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
  float d0, d1;
  unpack2(dd, d0, d1);
  even = __ballot_sync(kFull, d0 < r2);
  odd = __ballot_sync(kFull, d1 < r2);
}

// Append the hits of one 64-point step in ascending k: point order is lane0.even, lane0.odd, lane1.even, ...
__device__ __forceinline__ void bq_append2(unsigned even, unsigned odd, int kbase, int lane, int nsample, int& cnt,
                                           int& first, int* __restrict__ row) {
  if (even | odd) {
    if (cnt == 0) {
      const int fe = even ? __ffs(even) - 1 : 64, fo = odd ? __ffs(odd) - 1 : 64;
      first = kbase + (fe <= fo ? 2 * fe : 2 * fo + 1);
    }
    const unsigned lt = (1u << lane) - 1u;
    const int before = __popc(even & lt) + __popc(odd & lt);
    const int he = (even >> lane) & 1u, ho = (odd >> lane) & 1u;
    int slot = cnt + before;
    if (he && slot < nsample) row[slot] = kbase + 2 * lane;
    slot += he;
    if (ho && slot < nsample) row[slot] = kbase + 2 * lane + 1;
    cnt += __popc(even) + __popc(odd);
  }
}

__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                  int n, int m, float radius2, int nsample, int* __restrict__ idx) {
  __shared__ __align__(8) float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE];
  __shared__ int s_cnt[BQ_QPB], s_first[BQ_QPB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * m * 3;
  idx += (size_t)b * m * nsample;
  const int q0 = blockIdx.x * BQ_QPB + warp * BQ_QPW;
  if (lane < BQ_QPW) {
    s_cnt[warp * BQ_QPW + lane] = (q0 + lane < m) ? 0 : nsample;  // out-of-range queries are "done"
    s_first[warp * BQ_QPW + lane] = 0;
  }
  for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
    const int tn = min(BQ_TILE, n - t0);
    const int tn_pad = (tn + 127) & ~127;  // sentinel padding: far-away points never hit
    __syncthreads();
    for (int i = tid; i < tn * 3; i += BQ_WARPS * 32) {
      float v = xyz[(size_t)t0 * 3 + i];
      int k = i / 3, c = i - k * 3;
      (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
    }
    for (int k = tn + tid; k < tn_pad; k += BQ_WARPS * 32) sx[k] = sy[k] = sz[k] = 1e30f;
    __syncthreads();
    for (int q = 0; q < BQ_QPW; ++q) {
      int cnt = s_cnt[warp * BQ_QPW + q];
      if (cnt >= nsample) continue;
      int first = s_first[warp * BQ_QPW + q];
      const int j = q0 + q;
      const float qx = new_xyz[j * 3 + 0], qy = new_xyz[j * 3 + 1], qz = new_xyz[j * 3 + 2];
      int* row = idx + (size_t)j * nsample;
      const unsigned long long nqx = pack2(-qx, -qx), nqy = pack2(-qy, -qy), nqz = pack2(-qz, -qz);
      for (int base = 0; base < tn_pad; base += 128) {
        const int k = base + 2 * lane;
        unsigned e0, o0, e1, o1;
        bq_hits2(sx, sy, sz, k, nqx, nqy, nqz, radius2, e0, o0);
        bq_hits2(sx, sy, sz, k + 64, nqx, nqy, nqz, radius2, e1, o1);
        if (e0 | o0 | e1 | o1) {
          bq_append2(e0, o0, t0 + base, lane, nsample, cnt, first, row);
          bq_append2(e1, o1, t0 + base + 64, lane, nsample, cnt, first, row);
          if (cnt >= nsample) break;
        }
      }
      if (lane == 0) {
        s_cnt[warp * BQ_QPW + q] = cnt;
        s_first[warp * BQ_QPW + q] = first;
      }
      __syncwarp();
    }
  }
  // complete the rows: slots [cnt, nsample) <- first hit (or 0 when no hit)
  for (int q = 0; q < BQ_QPW; ++q) {
    const int j = q0 + q;
    if (j >= m) continue;
    int* row = idx + (size_t)j * nsample;
    const int cnt = s_cnt[warp * BQ_QPW + q];
    const int fillv = cnt > 0 ? s_first[warp * BQ_QPW + q] : 0;
    for (int s = min(cnt, nsample) + lane; s < nsample; s += 32) row[s] = fillv;
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
// Row gather (channel-last):  out[b,j,:] = x[b,idx[b,j],:]   x:(b,n,c)  idx:(b,m)
// What sample_pts_feats / gather_pts_feats need (model_utils.py:137-212).  The reference reaches
// it through transpose -> channel-first gather_points -> transpose, i.e. two full copies of the
// (B,N,C) tensor per call; here each output row is one coalesced 128-bit-vectorised row copy.
// ---------------------------------------------------------------------------
constexpr int GR_THREADS = 256;

template <bool VEC>
__global__ void __launch_bounds__(GR_THREADS)
gather_rows_kernel(const float* __restrict__ x, const int* __restrict__ idx, int n, int m, int c,
                   float* __restrict__ out) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * (GR_THREADS / 32) + warp;
  if (j >= m) return;
  const int src = __ldg(idx + (size_t)b * m + j);
  const float* p = x + ((size_t)b * n + src) * c;
  float* o = out + ((size_t)b * m + j) * c;
  if (VEC) {
    const float4* p4 = reinterpret_cast<const float4*>(p);
    float4* o4 = reinterpret_cast<float4*>(o);
    for (int k = lane; k < c / 4; k += 32) o4[k] = __ldg(p4 + k);
  } else {
    for (int k = lane; k < c; k += 32) o[k] = __ldg(p + k);
  }
}

__global__ void __launch_bounds__(GR_THREADS)
gather_rows_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx, int n, int m, int c,
                        float* __restrict__ grad_x) {
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * (GR_THREADS / 32) + warp;
  if (j >= m) return;
  const int src = __ldg(idx + (size_t)b * m + j);
  const float* g = grad_out + ((size_t)b * m + j) * c;
  float* o = grad_x + ((size_t)b * n + src) * c;
  for (int k = lane; k < c; k += 32) atomicAdd(o + k, g[k]);
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
  const int cfg = fps_cfg();
  if (n <= 2048) {
    if (cfg == 1) return launch_fps<1024, 2>(xyz, b, n, m, bs_log2, idx_out, st);
    if (cfg == 2) return launch_fps<512, 4>(xyz, b, n, m, bs_log2, idx_out, st);
    if (cfg == 4) return launch_fps<512, 4, true>(xyz, b, n, m, bs_log2, idx_out, st);
    if (cfg == 5) return launch_fps<256, 8, true, true>(xyz, b, n, m, bs_log2, idx_out, st);
    if (cfg == 6) return launch_fps<128, 16, true, true>(xyz, b, n, m, bs_log2, idx_out, st);
    return launch_fps<512, 4, true>(xyz, b, n, m, bs_log2, idx_out, st);
  }
  if (n <= 3072) return launch_fps<1024, 3>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 4096) return launch_fps<1024, 4>(xyz, b, n, m, bs_log2, idx_out, st);
  if (n <= 5120) {
    if (cfg == 1) return launch_fps<1024, 5>(xyz, b, n, m, bs_log2, idx_out, st);
    if (cfg == 2) return launch_fps<512, 10>(xyz, b, n, m, bs_log2, idx_out, st);
    if (cfg == 4) return launch_fps<512, 10, true>(xyz, b, n, m, bs_log2, idx_out, st);
    if (cfg == 5) return launch_fps<256, 20, true, true>(xyz, b, n, m, bs_log2, idx_out, st);
    if (cfg == 6) return launch_fps<128, 40, true, true>(xyz, b, n, m, bs_log2, idx_out, st);
    return launch_fps<512, 10, true>(xyz, b, n, m, bs_log2, idx_out, st);
  }
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

int upk_gather_rows(const float* x, const int* idx, int b, int n, int m, int c, float* out,
                    upk_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || c < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || m == 0 || c == 0) return UPK_OK;
  dim3 grid(ceil_div(m, GR_THREADS / 32), b);
  bool vec = (c % 4 == 0) && (((uintptr_t)x | (uintptr_t)out) % 16 == 0);
  if (vec) gather_rows_kernel<true><<<grid, GR_THREADS, 0, (cudaStream_t)stream>>>(x, idx, n, m, c, out);
  else gather_rows_kernel<false><<<grid, GR_THREADS, 0, (cudaStream_t)stream>>>(x, idx, n, m, c, out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_gather_rows_grad(const float* grad_out, const int* idx, int b, int n, int m, int c,
                         float* grad_x, upk_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || c < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || n == 0 || c == 0) return UPK_OK;
  cudaStream_t st = (cudaStream_t)stream;
  UPK_CUDA_TRY(cudaMemsetAsync(grad_x, 0, (size_t)b * n * c * sizeof(float), st));
  if (m == 0) return UPK_OK;
  dim3 grid(ceil_div(m, GR_THREADS / 32), b);
  gather_rows_grad_kernel<<<grid, GR_THREADS, 0, st>>>(grad_out, idx, n, m, c, grad_x);
  count_launch();
  UPK_RETURN_LAST_ERROR();
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
