// (5) Fused ball query (+ grouping of the xyz channels) for one or two radii in ONE scan of the cloud.
//
// Reference: QueryAndGroup / QueryAndLRFGroup (pointnet2_utils.py:292-378, :484-584) run, per scale,
// ball_query (ball_query_gpu.cu:14-49; one CTA per instance, thread-serial scan) and then
// grouping_operation on the transposed cloud (group_points_gpu.cu:13-33).  PositionalEncoding calls this
// for two scales on the same cloud (fine module :159-178), i.e. four launches and two full scans.
//
// Here: thread-per-query scan with the query in registers and the cloud broadcast from shared memory
// (SoA, LDS.128 = 4 points per load, packed f32x2 distance arithmetic, bit-identical to the reference's
// fma(dz,dz, fma(dx,dx, dy*dy))).  Hits are recorded as BIT MASKS (one 32-point word per register), so
// the ascending-index order of the reference falls out of the bit order.  Only the OUTER radius is
// tested in the scan; the inner radius is re-tested on the few outer hits during emission (the same
// expression on the same operands gives the same bits).  Emission is warp-per-query: popc + warp prefix
// scan give every hit its slot; rows are completed (first-hit padding / zero rows) and the grouped xyz
// rows (b,3,m,nsample) are written with coalesced warp stores from the smem copy of the cloud.
#include "common.cuh"
#include "launch_count.h"
#include "../../include/unopose_b200.h"

namespace upk {

constexpr int BG_THREADS = 256;
constexpr int BG_WARPS = BG_THREADS / 32;
constexpr int BG_QPB = 64;                 // queries per CTA (two warps of queries)
constexpr int BG_PARTS = BG_WARPS / 2;     // the tile's points are split over 4 warp pairs
constexpr int BG_TILE = 2048;              // points per smem tile
constexpr int BG_WORDS = BG_TILE / 32;     // 64 mask words per query per tile
constexpr int BG_WPP = BG_WORDS / BG_PARTS;  // 16 words per thread per tile
constexpr int BG_MPITCH = BG_WORDS + 1;    // +1: a warp stores one word index for 32 queries -> no bank conflicts

struct BgScale {
  float r2;      // radius*radius (fp32 product, ball_query_gpu.cu:27)
  int ns;        // nsample
  int* idx;      // [b,m,ns]
  float* grp;    // [b,3,m,ns] or nullptr
};

__device__ __forceinline__ float bg_d2(float px, float py, float pz, float nqx, float nqy, float nqz) {
  return sqdist_ref(px + nqx, py + nqy, pz + nqz);
}

// 32 points -> hit mask against r2 (packed arithmetic, two points per instruction)
__device__ __forceinline__ unsigned bg_scan_word(const float* __restrict__ sx, const float* __restrict__ sy,
                                                 const float* __restrict__ sz, int k0, unsigned long long nqx,
                                                 unsigned long long nqy, unsigned long long nqz, float r2) {
  unsigned mask = 0u;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(sx + k0 + 4 * g);
    const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(sy + k0 + 4 * g);
    const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(sz + k0 + 4 * g);
    unsigned long long dx = add2(X.x, nqx), dy = add2(Y.x, nqy), dz = add2(Z.x, nqz);
    unsigned long long da = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    dx = add2(X.y, nqx); dy = add2(Y.y, nqy); dz = add2(Z.y, nqz);
    unsigned long long db = fma2(dz, dz, fma2(dx, dx, mul2(dy, dy)));
    float d0, d1, d2, d3;
    unpack2(da, d0, d1);
    unpack2(db, d2, d3);
    if (d0 < r2) mask |= 1u << (4 * g);
    if (d1 < r2) mask |= 1u << (4 * g + 1);
    if (d2 < r2) mask |= 1u << (4 * g + 2);
    if (d3 < r2) mask |= 1u << (4 * g + 3);
  }
  return mask;
}

// inclusive warp scan of a packed pair of 16-bit counters
__device__ __forceinline__ unsigned bg_incl_scan(unsigned v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

template <bool TWO>
__global__ void __launch_bounds__(BG_THREADS)
ball_group_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz, int n, int m,
                  BgScale in, BgScale out) {
  // `out` = the scan radius (the larger one when TWO); `in` = the inner radius, a subset of `out`'s hits
  __shared__ __align__(16) float sx[BG_TILE];
  __shared__ __align__(16) float sy[BG_TILE];
  __shared__ __align__(16) float sz[BG_TILE];
  __shared__ unsigned s_mask[BG_QPB * BG_MPITCH];
  __shared__ int s_cnt[2][BG_QPB], s_first[2][BG_QPB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * m * 3;
  const int q0 = blockIdx.x * BG_QPB;
  int* const idx_o = out.idx + (size_t)b * m * out.ns;
  int* const idx_i = TWO ? in.idx + (size_t)b * m * in.ns : nullptr;

  if (tid < BG_QPB) {
    s_cnt[0][tid] = s_cnt[1][tid] = 0;
    s_first[0][tid] = s_first[1][tid] = 0;
  }
  // scan role: query (warp & 1) * 32 + lane, point part warp >> 1
  const int sq = (warp & 1) * 32 + lane, part = warp >> 1;
  const int sj = min(q0 + sq, m - 1);
  const unsigned long long nqx = pack2(-new_xyz[sj * 3 + 0], -new_xyz[sj * 3 + 0]);
  const unsigned long long nqy = pack2(-new_xyz[sj * 3 + 1], -new_xyz[sj * 3 + 1]);
  const unsigned long long nqz = pack2(-new_xyz[sj * 3 + 2], -new_xyz[sj * 3 + 2]);

  for (int t0 = 0; t0 < n; t0 += BG_TILE) {
    const int tn = min(BG_TILE, n - t0);
    __syncthreads();  // previous tile fully consumed (and the counters initialised)
    for (int i = tid; i < tn * 3; i += BG_THREADS) {
      const float v = xyz[(size_t)t0 * 3 + i];
      const int k = i / 3, c = i - k * 3;
      (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
    }
    for (int k = tn + tid; k < BG_TILE; k += BG_THREADS) sx[k] = sy[k] = sz[k] = 1e30f;  // never hit
    __syncthreads();
    // ---- scan: 16 words of 32 points per thread
    const int nwords = (tn + 31) >> 5;
#pragma unroll 1
    for (int w = 0; w < BG_WPP; ++w) {
      const int word = part * BG_WPP + w;
      unsigned mask = 0u;
      if (word < nwords) mask = bg_scan_word(sx, sy, sz, word * 32, nqx, nqy, nqz, out.r2);
      s_mask[sq * BG_MPITCH + word] = mask;
    }
    __syncthreads();
    // ---- emission: warp per query
    for (int qi = warp; qi < BG_QPB; qi += BG_WARPS) {
      const int j = q0 + qi;
      if (j >= m) break;
      int cnt_o = s_cnt[0][qi], cnt_i = s_cnt[1][qi];
      if (cnt_o >= out.ns && (!TWO || cnt_i >= in.ns)) continue;
      const float qx = -new_xyz[j * 3 + 0], qy = -new_xyz[j * 3 + 1], qz = -new_xyz[j * 3 + 2];
      int* const row_o = idx_o + (size_t)j * out.ns;
      int* const row_i = TWO ? idx_i + (size_t)j * in.ns : nullptr;
#pragma unroll 1
      for (int half = 0; half < BG_WORDS / 32; ++half) {
        const unsigned w = s_mask[qi * BG_MPITCH + half * 32 + lane];
        const unsigned any = __ballot_sync(kFull, w != 0u);
        if (any == 0u) continue;
        const int kb = half * 1024 + lane * 32;  // tile-local index of bit 0 of my word
        unsigned win = 0u;
        if (TWO) {
          unsigned u = w;
          while (u) {
            const int bit = __ffs(u) - 1;
            u &= u - 1;
            const int k = kb + bit;
            if (bg_d2(sx[k], sy[k], sz[k], qx, qy, qz) < in.r2) win |= 1u << bit;
          }
        }
        const unsigned v = (unsigned)__popc(w) | ((unsigned)__popc(win) << 16);
        const unsigned incl = bg_incl_scan(v, lane);
        const unsigned tot = __shfl_sync(kFull, incl, 31);
        const unsigned excl = incl - v;
        if (cnt_o == 0) {  // first hit overall = lowest bit of the lowest lane that has one
          const int src = __ffs(any) - 1;
          const int f = __shfl_sync(kFull, t0 + kb + __ffs(w) - 1, src);
          if (lane == 0) s_first[0][qi] = f;
        }
        if (TWO && cnt_i == 0) {
          const unsigned anyi = __ballot_sync(kFull, win != 0u);
          if (anyi) {
            const int src = __ffs(anyi) - 1;
            const int f = __shfl_sync(kFull, t0 + kb + __ffs(win) - 1, src);
            if (lane == 0) s_first[1][qi] = f;
          }
        }
        int so = cnt_o + (int)(excl & 0xffffu), si = cnt_i + (int)(excl >> 16);
        unsigned u = w;
        while (u) {
          const int bit = __ffs(u) - 1;
          u &= u - 1;
          if (so < out.ns) row_o[so] = t0 + kb + bit;
          ++so;
          if (TWO && ((win >> bit) & 1u)) {
            if (si < in.ns) row_i[si] = t0 + kb + bit;
            ++si;
          }
        }
        cnt_o += (int)(tot & 0xffffu);
        cnt_i += (int)(tot >> 16);
      }
      if (lane == 0) {
        s_cnt[0][qi] = cnt_o;
        s_cnt[1][qi] = cnt_i;
      }
    }
  }
  __syncthreads();  // every row's hits are in global memory (same-CTA visibility), counters final
  // ---- completion: pad the rows with the first hit (zero rows when no hit) + grouped xyz
  const bool tile_resident = n <= BG_TILE;  // the smem tile still holds the whole cloud
#pragma unroll 1
  for (int sc = 0; sc < (TWO ? 2 : 1); ++sc) {
    const BgScale S = sc == 0 ? out : in;
    int* const idx_s = S.idx + (size_t)b * m * S.ns;
    float* const g = S.grp ? S.grp + (size_t)b * 3 * m * S.ns : nullptr;
    const size_t cstride = (size_t)m * S.ns;
    for (int qi = warp; qi < BG_QPB; qi += BG_WARPS) {
      const int j = q0 + qi;
      if (j >= m) break;
      const int cnt = min(s_cnt[sc][qi], S.ns);
      const int fillv = cnt > 0 ? s_first[sc][qi] : 0;
      int* const row = idx_s + (size_t)j * S.ns;
      for (int s = lane; s < S.ns; s += 32) {
        int k = fillv;
        if (s < cnt) k = row[s];
        else row[s] = fillv;
        if (g) {
          float x, y, z;
          if (tile_resident) { x = sx[k]; y = sy[k]; z = sz[k]; }
          else { x = __ldg(xyz + k * 3 + 0); y = __ldg(xyz + k * 3 + 1); z = __ldg(xyz + k * 3 + 2); }
          float* o = g + (size_t)j * S.ns + s;
          __stcs(o, x);
          __stcs(o + cstride, y);
          __stcs(o + 2 * cstride, z);
        }
      }
    }
  }
}

}  // namespace upk

using namespace upk;

extern "C" {

int upk_ball_query_group(const float* new_xyz, const float* xyz, int b, int n, int m,
                         float radius0, int nsample0, int* idx0, float* grouped0,
                         float radius1, int nsample1, int* idx1, float* grouped1,
                         upk_stream_t stream) {
  if (b < 0 || n < 0 || m < 0 || nsample0 < 0 || nsample1 < 0) return UPK_ERR_INVALID_ARG;
  if (b == 0 || m == 0) return UPK_OK;
  if (nsample0 > 0xffff || nsample1 > 0xffff || n >= (1 << 30)) return UPK_ERR_UNSUPPORTED;
  if (nsample0 == 0 && nsample1 == 0) return UPK_OK;
  if ((nsample0 > 0 && !idx0) || (nsample1 > 0 && !idx1) || !new_xyz || (n > 0 && !xyz)) return UPK_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  BgScale s0{radius0 * radius0, nsample0, idx0, grouped0};
  BgScale s1{radius1 * radius1, nsample1, idx1, grouped1};
  dim3 grid(ceil_div(m, BG_QPB), b);
  if (nsample0 > 0 && nsample1 > 0) {
    // the scan runs on the larger radius; the other one is a subset of its hits
    const bool swap = s0.r2 > s1.r2;
    ball_group_kernel<true><<<grid, BG_THREADS, 0, st>>>(new_xyz, xyz, n, m, swap ? s1 : s0, swap ? s0 : s1);
  } else {
    const BgScale s = nsample0 > 0 ? s0 : s1;
    ball_group_kernel<false><<<grid, BG_THREADS, 0, st>>>(new_xyz, xyz, n, m, s, s);
  }
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

}  // extern "C"
