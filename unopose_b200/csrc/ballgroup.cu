// (5) Fused ball query (+ grouping of the xyz channels) for one or two radii in ONE scan of the cloud.
//
// Reference: QueryAndGroup / QueryAndLRFGroup (pointnet2_utils.py:292-378, :484-584) run, per scale,
// ball_query (ball_query_gpu.cu:14-49; one CTA per instance, thread-serial scan) and then
// grouping_operation on the transposed cloud (group_points_gpu.cu:13-33).  PositionalEncoding calls this
// for two scales on the same cloud (fine module :159-178), i.e. four launches and two full scans.
//
// Here: thread-per-query scan with the query in registers and the cloud broadcast from shared memory
// (SoA, LDS.128 = 4 points per load, packed f32x2 distance arithmetic, bit-identical to the reference's
// fma(dz,dz, fma(dx,dx, dy*dy))).  Hits are recorded as BIT MASKS: the sign bit of (d2 - r2) is shifted
// into a 32-point word (one packed subtract per two points + one funnel shift per point), so the
// ascending-index order of the reference falls out of the bit order.  Only the OUTER radius is tested in
// the scan; the inner radius is re-tested on the few outer hits during emission (the same expression on the
// same operands gives the same bits).  Emission is warp-per-query: popc + one warp prefix scan give every
// hit its slot.  When the whole cloud fits one tile (n <= 2048: the UNOPose shapes) the row is staged in
// shared memory, completed (first-hit padding / zero rows) and written together with the grouped xyz rows
// (b,3,m,nsample) by 128-bit coalesced stores; longer clouds accumulate their rows in global memory.
#include "common.cuh"
#include "launch_count.h"
#include "../../include/unopose_b200.h"

namespace upk {

constexpr int BG_THREADS = 256;
constexpr int BG_WARPS = BG_THREADS / 32;
constexpr int BG_QPB = 64;                 // queries per CTA (every scan thread takes queries lane and 32 + lane)
constexpr int BG_TILE = 2048;              // points per smem tile
constexpr int BG_WORDS = BG_TILE / 32;     // 64 mask words per query per tile
constexpr int BG_WPP2 = BG_WORDS / BG_WARPS;  // 8 words per thread per tile, two queries per thread
constexpr int BG_MPITCH = BG_WORDS + 2;    // even (64-bit reads of word pairs), 2-way conflicts on the 16 stores only
constexpr int BG_ROWCAP = 512;             // staged path: nsample0 + nsample1 <= 512 ints per warp

struct BgScale {
  float r2;      // radius*radius (fp32 product, ball_query_gpu.cu:27)
  int ns;        // nsample
  int* idx;      // [b,m,ns]
  float* grp;    // [b,3,m,ns] or nullptr
};

__device__ __forceinline__ float bg_d2(float px, float py, float pz, float nqx, float nqy, float nqz) {
  return sqdist_ref(px + nqx, py + nqy, pz + nqz);
}

// 32 points -> hit masks against r2 for TWO queries: the three LDS.128 of a step serve both.  (A uniform-address LDS.128
// costs two shared-memory wavefronts; with one query per thread the scan asked for 24 wavefront cycles per 21 issue
// cycles of the SM.)  d2 < r2  <=>  sign(d2 - r2) for every finite or infinite d2 (round to nearest never turns a
// non-zero difference into zero, x - x = +0, NaN results are the canonical positive NaN).
__device__ __forceinline__ void bg_scan_word2(const float* __restrict__ sx, const float* __restrict__ sy,
                                              const float* __restrict__ sz, int k0, const unsigned long long (&nqa)[3],
                                              const unsigned long long (&nqb)[3], unsigned long long nr2,
                                              unsigned& mask_a, unsigned& mask_b) {
  unsigned ma = 0u, mb = 0u;
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(sx + k0 + 4 * g);
    const ulonglong2 Y = *reinterpret_cast<const ulonglong2*>(sy + k0 + 4 * g);
    const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(sz + k0 + 4 * g);
    unsigned long long dx = add2(X.x, nqa[0]), dy = add2(Y.x, nqa[1]), dz = add2(Z.x, nqa[2]);
    const unsigned long long sa = add2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), nr2);
    dx = add2(X.y, nqa[0]); dy = add2(Y.y, nqa[1]); dz = add2(Z.y, nqa[2]);
    const unsigned long long sb = add2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), nr2);
    dx = add2(X.x, nqb[0]); dy = add2(Y.x, nqb[1]); dz = add2(Z.x, nqb[2]);
    const unsigned long long ta = add2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), nr2);
    dx = add2(X.y, nqb[0]); dy = add2(Y.y, nqb[1]); dz = add2(Z.y, nqb[2]);
    const unsigned long long tb = add2(fma2(dz, dz, fma2(dx, dx, mul2(dy, dy))), nr2);
    ma = __funnelshift_l((unsigned)sa, ma, 1);
    ma = __funnelshift_l((unsigned)(sa >> 32), ma, 1);
    ma = __funnelshift_l((unsigned)sb, ma, 1);
    ma = __funnelshift_l((unsigned)(sb >> 32), ma, 1);
    mb = __funnelshift_l((unsigned)ta, mb, 1);
    mb = __funnelshift_l((unsigned)(ta >> 32), mb, 1);
    mb = __funnelshift_l((unsigned)tb, mb, 1);
    mb = __funnelshift_l((unsigned)(tb >> 32), mb, 1);
  }
  mask_a = __brev(ma);
  mask_b = __brev(mb);
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

// Stage one tile of the cloud (AoS global -> SoA shared), sentinel-pad the rest of the tile.
__device__ __forceinline__ void bg_load_tile(const float* __restrict__ xyz, int t0, int tn, float* sx, float* sy,
                                             float* sz) {
  const int tid = threadIdx.x;
  const float* src = xyz + (size_t)t0 * 3;
  int done = 0;
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const int nv = (tn * 3) >> 2;
    for (int q = tid; q < nv; q += BG_THREADS) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + q);
      const float vv[4] = {v.x, v.y, v.z, v.w};
      const int e0 = 4 * q, k = e0 / 3, c = e0 - 3 * k;  // element e0 + i: point k + (c + i) / 3, component (c + i) % 3
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ci = c + i;
        const int wrap = (ci >= 3 ? 1 : 0) + (ci >= 6 ? 1 : 0);
        const int cc = ci - 3 * wrap;
        (cc == 0 ? sx : (cc == 1 ? sy : sz))[k + wrap] = vv[i];
      }
    }
    done = 4 * nv;
  }
  for (int e = done + tid; e < tn * 3; e += BG_THREADS) {
    const int k = e / 3, c = e - 3 * k;
    (c == 0 ? sx : (c == 1 ? sy : sz))[k] = src[e];
  }
  for (int k = tn + tid; k < BG_TILE; k += BG_THREADS) sx[k] = sy[k] = sz[k] = 1e30f;  // never hit
}

// STAGED: the cloud is a single tile; rows live in shared memory until they are complete.
template <bool TWO, bool STAGED>
__global__ void __launch_bounds__(BG_THREADS)
ball_group_kernel(const float* __restrict__ new_xyz, const float* __restrict__ xyz, int n, int m,
                  BgScale in, BgScale out) {
  // `out` = the scan radius (the larger one when TWO); `in` = the inner radius, a subset of `out`'s hits
  extern __shared__ __align__(16) unsigned char bg_smem[];
  float* sx = reinterpret_cast<float*>(bg_smem);
  float* sy = sx + BG_TILE;
  float* sz = sy + BG_TILE;
  unsigned* s_mask = reinterpret_cast<unsigned*>(sz + BG_TILE);          // BG_QPB * BG_MPITCH
  int* s_rows = reinterpret_cast<int*>(s_mask + BG_QPB * BG_MPITCH);     // STAGED: BG_WARPS * BG_ROWCAP
  __shared__ int s_cnt[2][BG_QPB], s_first[2][BG_QPB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = blockIdx.y;
  xyz += (size_t)b * n * 3;
  new_xyz += (size_t)b * m * 3;
  const int q0 = blockIdx.x * BG_QPB;
  int* const idx_o = out.idx + (size_t)b * m * out.ns;
  int* const idx_i = TWO ? in.idx + (size_t)b * m * in.ns : nullptr;

  if (!STAGED && tid < BG_QPB) {
    s_cnt[0][tid] = s_cnt[1][tid] = 0;
    s_first[0][tid] = s_first[1][tid] = 0;
  }
  // scan role: queries lane and 32 + lane of the CTA, words [warp * BG_WPP2, (warp + 1) * BG_WPP2) of the tile
  unsigned long long nqa[3], nqb[3];
  {
    const int ja = min(q0 + lane, m - 1), jb = min(q0 + 32 + lane, m - 1);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      nqa[a] = pack2(-new_xyz[ja * 3 + a], -new_xyz[ja * 3 + a]);
      nqb[a] = pack2(-new_xyz[jb * 3 + a], -new_xyz[jb * 3 + a]);
    }
  }
  const unsigned long long nr2 = pack2(-out.r2, -out.r2);

  for (int t0 = 0; t0 < n; t0 += BG_TILE) {
    const int tn = min(BG_TILE, n - t0);
    __syncthreads();  // previous tile fully consumed (and the counters initialised)
    bg_load_tile(xyz, t0, tn, sx, sy, sz);
    __syncthreads();
    // ---- scan: 8 words of 32 points for two queries per thread
    const int nwords = (tn + 31) >> 5;
#pragma unroll 1
    for (int w = 0; w < BG_WPP2; ++w) {
      const int word = warp * BG_WPP2 + w;
      unsigned mask_a = 0u, mask_b = 0u;
      if (word < nwords) bg_scan_word2(sx, sy, sz, word * 32, nqa, nqb, nr2, mask_a, mask_b);
      s_mask[lane * BG_MPITCH + word] = mask_a;
      s_mask[(32 + lane) * BG_MPITCH + word] = mask_b;
    }
    __syncthreads();
    // ---- emission: warp per query; lane l owns points [64 l, 64 l + 64) of the tile
    for (int qi = warp; qi < BG_QPB; qi += BG_WARPS) {
      const int j = q0 + qi;
      if (j >= m) break;
      int cnt_o = 0, cnt_i = 0;
      if (!STAGED) {
        cnt_o = s_cnt[0][qi];
        cnt_i = s_cnt[1][qi];
        if (cnt_o >= out.ns && (!TWO || cnt_i >= in.ns)) continue;
      }
      int* const row_o = STAGED ? s_rows + warp * BG_ROWCAP : idx_o + (size_t)j * out.ns;
      int* const row_i = STAGED ? row_o + ((out.ns + 3) & ~3) : (TWO ? idx_i + (size_t)j * in.ns : nullptr);
      const uint2 w2 = *reinterpret_cast<const uint2*>(s_mask + qi * BG_MPITCH + 2 * lane);
      const int kb = 64 * lane;  // tile-local index of bit 0 of my first word
      // outer hits: slots from a warp prefix scan of the per-lane bit counts, emitted in ascending index;
      // the inner radius is re-tested on each outer hit in the same loop (same expression, same operands)
      const unsigned vo = (unsigned)(__popc(w2.x) + __popc(w2.y));
      const unsigned incl_o = bg_incl_scan(vo, lane);
      const unsigned tot_o = __shfl_sync(kFull, incl_o, 31);
      uint2 win = make_uint2(0u, 0u);
      float qx = 0.f, qy = 0.f, qz = 0.f;
      if (TWO) { qx = -new_xyz[j * 3 + 0]; qy = -new_xyz[j * 3 + 1]; qz = -new_xyz[j * 3 + 2]; }
      int so = cnt_o + (int)(incl_o - vo);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        unsigned u = h ? w2.y : w2.x;
        unsigned wi = 0u;
        while (u) {
          const int bit = __ffs(u) - 1;
          u &= u - 1;
          const int kl = kb + 32 * h + bit;
          if (so < out.ns) row_o[so] = t0 + kl;
          ++so;
          if (TWO && bg_d2(sx[kl], sy[kl], sz[kl], qx, qy, qz) < in.r2) wi |= 1u << bit;
        }
        if (h) win.y = wi; else win.x = wi;
      }
      unsigned tot_i = 0u;
      if (TWO) {
        const unsigned vi = (unsigned)(__popc(win.x) + __popc(win.y));
        const unsigned incl_i = bg_incl_scan(vi, lane);
        tot_i = __shfl_sync(kFull, incl_i, 31);
        int si = cnt_i + (int)(incl_i - vi);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          unsigned u = h ? win.y : win.x;
          while (u) {
            const int bit = __ffs(u) - 1;
            u &= u - 1;
            if (si < in.ns) row_i[si] = t0 + kb + 32 * h + bit;
            ++si;
          }
        }
      }
      const unsigned tot = tot_o | (tot_i << 16);
      int first_o = 0, first_i = 0;
      if (!STAGED) {   // first hit overall = lowest bit of the lowest lane that has one (kept across tiles)
        if (cnt_o == 0 && tot_o) {
          const unsigned any = __ballot_sync(kFull, (w2.x | w2.y) != 0u);
          const int mine = t0 + kb + (w2.x ? __ffs(w2.x) - 1 : 32 + __ffs(w2.y) - 1);
          first_o = __shfl_sync(kFull, mine, __ffs(any) - 1);
          if (lane == 0) s_first[0][qi] = first_o;
        }
        if (TWO && cnt_i == 0 && tot_i) {
          const unsigned any = __ballot_sync(kFull, (win.x | win.y) != 0u);
          const int mine = t0 + kb + (win.x ? __ffs(win.x) - 1 : 32 + __ffs(win.y) - 1);
          first_i = __shfl_sync(kFull, mine, __ffs(any) - 1);
          if (lane == 0) s_first[1][qi] = first_i;
        }
      }
      cnt_o += (int)(tot & 0xffffu);
      cnt_i += (int)(tot >> 16);
      if (!STAGED) {
        if (lane == 0) {
          s_cnt[0][qi] = cnt_o;
          s_cnt[1][qi] = cnt_i;
        }
        continue;
      }
      // ---- STAGED completion of this query: pad, write idx and the grouped rows with 128-bit stores
      __syncwarp();
      first_o = row_o[0];   // slot 0 holds the first hit in ascending index (only read when cnt > 0)
      if (TWO) first_i = row_i[0];
#pragma unroll 1
      for (int sc = 0; sc < (TWO ? 2 : 1); ++sc) {
        const BgScale S = sc == 0 ? out : in;
        const int* row = sc == 0 ? row_o : row_i;
        const int cnt = min(sc == 0 ? cnt_o : cnt_i, S.ns);
        const int fillv = cnt > 0 ? (sc == 0 ? first_o : first_i) : 0;
        int* const grow = S.idx + ((size_t)b * m + j) * S.ns;
        float* const g = S.grp ? S.grp + (size_t)b * 3 * m * S.ns + (size_t)j * S.ns : nullptr;
        const size_t cstride = (size_t)m * S.ns;
        if ((S.ns & 3) == 0) {  // rows are 16-byte aligned whenever the base pointers are (checked by the launcher)
          for (int s4 = 4 * lane; s4 < S.ns; s4 += 128) {
            int4 k4 = *reinterpret_cast<const int4*>(row + s4);
            if (s4 + 0 >= cnt) k4.x = fillv;
            if (s4 + 1 >= cnt) k4.y = fillv;
            if (s4 + 2 >= cnt) k4.z = fillv;
            if (s4 + 3 >= cnt) k4.w = fillv;
            __stcs(reinterpret_cast<int4*>(grow + s4), k4);
            if (g) {
              __stcs(reinterpret_cast<float4*>(g + s4), make_float4(sx[k4.x], sx[k4.y], sx[k4.z], sx[k4.w]));
              __stcs(reinterpret_cast<float4*>(g + cstride + s4), make_float4(sy[k4.x], sy[k4.y], sy[k4.z], sy[k4.w]));
              __stcs(reinterpret_cast<float4*>(g + 2 * cstride + s4), make_float4(sz[k4.x], sz[k4.y], sz[k4.z], sz[k4.w]));
            }
          }
        } else {
          for (int s = lane; s < S.ns; s += 32) {
            const int k = s < cnt ? row[s] : fillv;
            grow[s] = k;
            if (g) {
              g[s] = sx[k];
              g[cstride + s] = sy[k];
              g[2 * cstride + s] = sz[k];
            }
          }
        }
      }
      __syncwarp();  // the row buffer is reused by this warp's next query
    }
  }
  if (STAGED) return;
  __syncthreads();  // every row's hits are in global memory (same-CTA visibility), counters final
  // ---- completion: pad the rows with the first hit (zero rows when no hit) + grouped xyz
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
          float* o = g + (size_t)j * S.ns + s;
          __stcs(o, __ldg(xyz + k * 3 + 0));
          __stcs(o + cstride, __ldg(xyz + k * 3 + 1));
          __stcs(o + 2 * cstride, __ldg(xyz + k * 3 + 2));
        }
      }
    }
  }
}

constexpr size_t BG_SMEM_BASE = (size_t)3 * BG_TILE * sizeof(float) + (size_t)BG_QPB * BG_MPITCH * sizeof(unsigned);
constexpr size_t BG_SMEM_STAGED = BG_SMEM_BASE + (size_t)BG_WARPS * BG_ROWCAP * sizeof(int);

template <bool TWO, bool STAGED>
static int launch_ball_group(const float* new_xyz, const float* xyz, int b, int n, int m, const BgScale& in,
                             const BgScale& out, cudaStream_t st) {
  auto kern = ball_group_kernel<TWO, STAGED>;
  const size_t smem = STAGED ? BG_SMEM_STAGED : BG_SMEM_BASE;
  if (smem > 47 * 1024)
    UPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(m, BG_QPB), b);
  kern<<<grid, BG_THREADS, smem, st>>>(new_xyz, xyz, n, m, in, out);
  count_launch();
  UPK_RETURN_LAST_ERROR();
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
  const bool two = nsample0 > 0 && nsample1 > 0;
  // shared-memory row staging: one tile, rows fit, and every 128-bit store is aligned
  auto aligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool staged = n <= BG_TILE && nsample0 + nsample1 + 3 <= BG_ROWCAP &&
                      (nsample0 == 0 || (aligned(idx0) && aligned(grouped0))) &&
                      (nsample1 == 0 || (aligned(idx1) && aligned(grouped1)));
  if (two) {
    // the scan runs on the larger radius; the other one is a subset of its hits
    const bool swap = s0.r2 > s1.r2;
    const BgScale& in = swap ? s1 : s0;
    const BgScale& out = swap ? s0 : s1;
    return staged ? launch_ball_group<true, true>(new_xyz, xyz, b, n, m, in, out, st)
                  : launch_ball_group<true, false>(new_xyz, xyz, b, n, m, in, out, st);
  }
  const BgScale& s = nsample0 > 0 ? s0 : s1;
  return staged ? launch_ball_group<false, true>(new_xyz, xyz, b, n, m, s, s, st)
                : launch_ball_group<false, false>(new_xyz, xyz, b, n, m, s, s, st);
}

}  // extern "C"
