// (1b) Dual-softmax assignment, LARGE geometry (fine stage: 2049 x 2049 logits per instance).
// Reference: compute_fine_Rt[_overlap], model_utils.py:493-566 (softmax over rows x softmax over columns
// x scores, background-vs-foreground arg-max tests, masked soft correspondences).
//
// Three streaming passes over `atten`, each reading every logit exactly once from HBM:
//   pass 1  k_fine_stats   sum of exponentials per row and per column
//   pass 2  k_fine_labels  max_j S_ij / max_i S_ij against the background entries  -> w1, w2
//   pass 3  k_fine_rows    sum_j S_ij w2_j {x_j, y_j, z_j, 1}                       -> soft correspondences
// with S_ij = softmax_row * softmax_col * s1_i * s2_j = 2^(2 v log2e - rml_i - cml_j) * rmul_i * cmul_j.
//
// Layout.  The background row 0 and column 0 are peeled off, so the main block (rows 1.., columns 1..)
// of the standard shape is exactly 2048 x 2048: no edge tiles.  A CTA of 8 warps owns 128 rows x one
// 256-column strip; a warp streams whole 256-column row segments from global memory into registers, two
// rows per step (lane l owns columns l, l+32, ..: every load instruction is one coalesced 128-byte line;
// `atten` rows are only 4-byte aligned, pitch 2049 floats, so neither 128-bit loads nor TMA apply), the
// next two rows are prefetched while the current two are processed.  Column constants stay in registers
// for the CTA's 128 rows, row constants are precomputed by the merge kernels (no division in the loops).
// Arithmetic is packed f32x2 where two independent values share an instruction.
//
// Pass 1 uses ONE reference exponent per warp instead of per-row / per-column maxima: any reference
// gives the same softmax as long as nothing overflows or underflows, so e = 2^(v log2e - g) serves the
// row sum AND the column sum (one MUFU per logit).  g starts 8 octaves above the first row pair's
// maximum and is raised (column sums rescaled) if a later partial sum exceeds 2^20.  Partial sums that
// come out below 2^-80 or non-finite (logit range > ~50: never for cosine/temperature logits) raise a
// per-instance flag; the merge kernel then recomputes that instance's statistics exactly (max-subtracting).
//
// Row pitch.  `atten` may be PITCHED (ld >= C floats per row).  compute_feature_similarity hands out a padded layout for
// the fine shape — ld % 4 == 0 and element (i, 1) 16-byte aligned — so that the GEMM stores tiles with TMA and these
// passes read them with 128-bit loads (VEC = true: a lane owns columns 4 lane .. 4 lane + 3 and 128 + 4 lane .. of the
// strip); a plain contiguous tensor (pitch 2049: rows only 4-byte aligned) takes the scalar loads (VEC = false: lane
// owns columns lane + 32 k).  Everything downstream of the loads is agnostic of which columns a lane owns.
//
// TMA-fed variants (round 2; the default for the padded layout on whole 128 x 256 tiles, i.e. the UNOPose fine shape):
// k_fine_labels_tma / k_fine_rows_tma further down are passes 2 and 3 as persistent kernels with a TMA box ring, a
// constants warp and L2 eviction hints — same partials bit for bit, 5.5-5.7 TB/s instead of 4.1-4.6.  The streaming
// kernels below remain for contiguous tensors, ragged shapes and UPK_FINE_TMA=0.
#include <math.h>

#include "common.cuh"
#include "launch_count.h"
#include "pose_internal.h"
#include "tc_ptx.cuh"
#include "../../include/unopose_b200.h"

namespace upk {

constexpr int F2_RT = 128;                   // rows per CTA
constexpr int F2_TC = 256;                   // columns per strip
constexpr int F2_CPT = F2_TC / 32;           // 8 columns per lane
constexpr int F2_WARPS = 8;
constexpr int F2_THREADS = F2_WARPS * 32;
constexpr int F2_PAIRS = F2_RT / (2 * F2_WARPS);  // 8 row pairs per warp
constexpr float kL2E = 1.4426950408889634f;
constexpr float kRefMargin = 8.f;            // octaves between a fresh reference and the observed maximum
constexpr float kSumHigh = 1048576.f;        // 2^20: raise the reference
constexpr float kSumLow = 8.271806e-25f;     // 2^-80: partial sums below this are not trusted

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ld_stream_f1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ bool sum_untrusted(float s) { return !(s >= kSumLow && s < INFINITY); }

// Sum / max of two per-lane values across the warp: lanes 0-15 return row a's total, lanes 16-31 row b's.
template <class Op>
__device__ __forceinline__ float reduce_pair(float a, float b, int lane, Op op) {
  const bool upper = (lane & 16) != 0;
  float v = op(upper ? b : a, __shfl_xor_sync(kFull, upper ? a : b, 16));
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
struct F2Add { __device__ __forceinline__ float operator()(float a, float b) const { return a + b; } };
struct F2Max { __device__ __forceinline__ float operator()(float a, float b) const { return fmaxf(a, b); } };

// Strip-local column of a lane's k-th value
template <bool VEC>
__device__ __forceinline__ int f2_lcol(int lane, int k) {
  return VEC ? ((k >> 2) * 128 + lane * 4 + (k & 3)) : (lane + 32 * k);
}

// Two row segments (rows ra and ra + 8) of the strip: 16 values per lane, all loads independent.
struct RowPair {
  float a[F2_CPT], b[F2_CPT];
};
// pa -> A[ra][first column of the strip]; nv = number of valid columns in the strip (C - c0)
template <bool CHECK, bool VEC>
__device__ __forceinline__ void load_pair(const float* __restrict__ pa, size_t pitch8, int nv, int lane, bool oka,
                                          bool okb, RowPair& v) {
  if (VEC) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int off = 128 * h + 4 * lane;
      if (!CHECK || off + 3 < nv) {
        float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY), y = x;
        if (!CHECK || oka) x = ld_stream_f4(reinterpret_cast<const float4*>(pa + off));
        if (!CHECK || okb) y = ld_stream_f4(reinterpret_cast<const float4*>(pa + pitch8 + off));
        v.a[4 * h] = x.x; v.a[4 * h + 1] = x.y; v.a[4 * h + 2] = x.z; v.a[4 * h + 3] = x.w;
        v.b[4 * h] = y.x; v.b[4 * h + 1] = y.y; v.b[4 * h + 2] = y.z; v.b[4 * h + 3] = y.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const bool c = off + e < nv;
          v.a[4 * h + e] = (oka && c) ? ld_stream_f1(pa + off + e) : -INFINITY;
          v.b[4 * h + e] = (okb && c) ? ld_stream_f1(pa + pitch8 + off + e) : -INFINITY;
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < F2_CPT; ++k) {
      if (CHECK) {
        const bool c = lane + 32 * k < nv;
        v.a[k] = (oka && c) ? ld_stream_f1(pa + lane + 32 * k) : -INFINITY;
        v.b[k] = (okb && c) ? ld_stream_f1(pa + pitch8 + lane + 32 * k) : -INFINITY;
      } else {
        v.a[k] = ld_stream_f1(pa + lane + 32 * k);
        v.b[k] = ld_stream_f1(pa + pitch8 + lane + 32 * k);
      }
    }
  }
}

// CTA coordinates shared by the three passes
struct F2Tile {
  int b, cs, rt, warp, lane;
  int i0;    // first row of this warp (rows i0 + 16 p and i0 + 16 p + 8, p < F2_PAIRS)
  int c0;    // first column of the strip (global index, >= 1)
  int nv;    // C - c0: strip-local column c is valid iff c < nv
  bool full; // no row / column of the CTA's tile is out of range
};
// flip: walk the instances in descending order (the pass then starts on the instances the previous kernel touched
// last, which are the ones still resident in L2)
__device__ __forceinline__ F2Tile f2_tile(int R, int C, bool flip) {
  F2Tile t;
  t.b = flip ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
  t.rt = blockIdx.y; t.cs = blockIdx.x;
  t.warp = threadIdx.x >> 5; t.lane = threadIdx.x & 31;
  t.i0 = 1 + t.rt * F2_RT + t.warp;
  t.c0 = 1 + t.cs * F2_TC;
  t.nv = C - t.c0;
  t.full = (1 + (t.rt + 1) * F2_RT <= R) && (1 + (t.cs + 1) * F2_TC <= C);
  return t;
}

// ------------------------------------------------------------------ pass 1: sums of exponentials
struct StatsState {
  float g;            // warp reference exponent (log2 units); -inf until the first row pair
  float cs[F2_CPT];   // column sums of 2^(v log2e - g) over this warp's rows
  float c0;           // same for the background column (strip 0, lane 0)
  bool bad;
};

// One row pair.  SINGLE: only row `ra` exists (the background row 0, or a ragged last pair).
template <bool CHECK, bool STRIP0>
__device__ __forceinline__ void stats_pair(const RowPair& v, float v0a, float v0b, int lane, StatsState& st,
                                           float2& out_a, float2& out_b) {
  if (st.g == -INFINITY) {  // first pair of this warp: reference = maximum + margin (warp-uniform branch)
    float m = fmaxf(v0a, v0b);
#pragma unroll
    for (int k = 0; k < F2_CPT; ++k) m = fmaxf(m, fmaxf(v.a[k], v.b[k]));
    m = warp_max(m);
    st.g = fmaf(m, kL2E, kRefMargin);
  }
  const float g = st.g;
  const unsigned long long L2 = pack2(kL2E, kL2E), NG = pack2(-g, -g);
  unsigned long long sa2 = pack2(0.f, 0.f), sb2 = sa2;
#pragma unroll
  for (int kk = 0; kk < F2_CPT / 2; ++kk) {
    float xa0, xa1, xb0, xb1;
    unpack2(fma2(pack2(v.a[2 * kk], v.a[2 * kk + 1]), L2, NG), xa0, xa1);
    unpack2(fma2(pack2(v.b[2 * kk], v.b[2 * kk + 1]), L2, NG), xb0, xb1);
    const unsigned long long ea = pack2(ex2_approx(xa0), ex2_approx(xa1));
    const unsigned long long eb = pack2(ex2_approx(xb0), ex2_approx(xb1));
    sa2 = add2(sa2, ea);
    sb2 = add2(sb2, eb);
    float c0, c1;
    unpack2(add2(add2(ea, eb), pack2(st.cs[2 * kk], st.cs[2 * kk + 1])), c0, c1);
    st.cs[2 * kk] = c0;
    st.cs[2 * kk + 1] = c1;
  }
  float sa, sb, t0, t1;
  unpack2(sa2, t0, t1); sa = t0 + t1;
  unpack2(sb2, t0, t1); sb = t0 + t1;
  if (STRIP0) {  // background column: lane 0 carries v0a / v0b, the other lanes -inf (-> 0)
    const float e0a = ex2_approx(fmaf(v0a, kL2E, -g)), e0b = ex2_approx(fmaf(v0b, kL2E, -g));
    sa += e0a; sb += e0b;
    st.c0 += e0a + e0b;
  }
  const float tot = reduce_pair(sa, sb, lane, F2Add());
  out_a = out_b = make_float2(g, tot);
  if (__any_sync(kFull, fmaxf(sa, sb) > kSumHigh)) {
    // some logit sits far above the reference: raise it for the rows to come (this pair's sums stay
    // attached to the old reference) and rescale the running column sums
    const float mx = warp_max(fmaxf(sa, sb));
    const float gn = g + lg2_approx(mx) + kRefMargin;
    const float sc = ex2_approx(g - gn);
#pragma unroll
    for (int k = 0; k < F2_CPT; ++k) st.cs[k] *= sc;
    st.c0 *= sc;
    st.g = gn;
  }
}

template <bool CHECK, bool STRIP0, bool VEC>
__device__ __forceinline__ void stats_body(const float* __restrict__ A, int R, int ld, int nstrip, const F2Tile& t,
                                           float2* __restrict__ rowpart, StatsState& st) {
  const size_t pitch8 = (size_t)8 * ld;
  const float* col0 = A;  // column 0 of row i: A[i * ld]
  auto emit = [&](int ra, bool okb, const float2& oa, const float2& ob) {
    if (t.lane == 0) {
      rowpart[((size_t)t.b * R + ra) * nstrip + t.cs] = oa;
      st.bad |= sum_untrusted(oa.y);
    }
    if (t.lane == 16 && okb) {
      rowpart[((size_t)t.b * R + ra + 8) * nstrip + t.cs] = ob;
      st.bad |= sum_untrusted(ob.y);
    }
  };
  if (t.rt == 0 && t.warp == 0) {  // the background row 0 rides with the first row tile
    RowPair v;
    load_pair<true, VEC>(A + t.c0, pitch8, t.nv, t.lane, true, false, v);
    float v0a = -INFINITY;
    if (STRIP0 && t.lane == 0) v0a = col0[0];
    float2 oa, ob;
    stats_pair<true, STRIP0>(v, v0a, -INFINITY, t.lane, st, oa, ob);
    if (t.lane == 0) {
      rowpart[((size_t)t.b * R) * nstrip + t.cs] = oa;
      st.bad |= sum_untrusted(oa.y);
    }
  }
  // two row pairs in flight: the loads of pair p + 1 are issued before pair p is processed
  auto fetch = [&](RowPair& v, float& z0a, float& z0b, int p) {
    const int ra = t.i0 + 16 * p;
    const bool oka = !CHECK || ra < R, okb = !CHECK || ra + 8 < R;
    load_pair<CHECK, VEC>(A + (size_t)ra * ld + t.c0, pitch8, t.nv, t.lane, oka, okb, v);
    if (STRIP0) {
      z0a = z0b = -INFINITY;
      if (t.lane == 0) {
        if (oka) z0a = col0[(size_t)ra * ld];
        if (okb) z0b = col0[(size_t)(ra + 8) * ld];
      }
    }
  };
  auto comp = [&](const RowPair& v, float z0a, float z0b, int p) {
    const int ra = t.i0 + 16 * p;
    if (CHECK && ra >= R) return;
    float2 oa, ob;
    stats_pair<CHECK, STRIP0>(v, z0a, z0b, t.lane, st, oa, ob);
    emit(ra, !CHECK || ra + 8 < R, oa, ob);
  };
  RowPair va, vb;
  float a0a = -INFINITY, a0b = -INFINITY, b0a = -INFINITY, b0b = -INFINITY;
  fetch(va, a0a, a0b, 0);
#pragma unroll 1
  for (int p = 0; p < F2_PAIRS; p += 2) {
    fetch(vb, b0a, b0b, p + 1);
    comp(va, a0a, a0b, p);
    if (p + 2 < F2_PAIRS) fetch(va, a0a, a0b, p + 2);
    comp(vb, b0a, b0b, p + 1);
  }
}

// (80 registers: 3 CTAs / SM, 54 us vs 59 us at 2)
template <bool VEC>
__global__ void __launch_bounds__(F2_THREADS, 3)
k_fine_stats(const float* __restrict__ atten, int R, int C, int ld, int nstrip, int nrt, float2* __restrict__ rowpart,
             float2* __restrict__ colpart, int* __restrict__ flags) {
  __shared__ float s_cs[F2_WARPS][F2_TC + 1];
  __shared__ float s_g[F2_WARPS];
  const F2Tile t = f2_tile(R, C, false);
  const float* A = atten + (size_t)t.b * R * ld;
  StatsState st;
  st.g = -INFINITY; st.c0 = 0.f; st.bad = false;
#pragma unroll
  for (int k = 0; k < F2_CPT; ++k) st.cs[k] = 0.f;
  if (t.cs == 0) {
    if (t.full) stats_body<false, true, VEC>(A, R, ld, nstrip, t, rowpart, st);
    else stats_body<true, true, VEC>(A, R, ld, nstrip, t, rowpart, st);
  } else {
    if (t.full) stats_body<false, false, VEC>(A, R, ld, nstrip, t, rowpart, st);
    else stats_body<true, false, VEC>(A, R, ld, nstrip, t, rowpart, st);
  }
  // combine the 8 warps' column sums (each attached to its own reference)
#pragma unroll
  for (int k = 0; k < F2_CPT; ++k) s_cs[t.warp][f2_lcol<VEC>(t.lane, k)] = st.cs[k];
  if (t.lane == 0) {
    s_cs[t.warp][F2_TC] = st.c0;
    s_g[t.warp] = st.g;
  }
  __syncthreads();
  float G = -INFINITY;
#pragma unroll
  for (int w = 0; w < F2_WARPS; ++w) G = fmaxf(G, s_g[w]);
  bool bad = st.bad;
  for (int c = threadIdx.x; c < F2_TC + (t.cs == 0 ? 1 : 0); c += F2_THREADS) {
    const int gj = c == F2_TC ? 0 : t.c0 + c;
    if (gj >= C) continue;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < F2_WARPS; ++w) {
      const float gw = s_g[w];
      if (gw != -INFINITY) s += s_cs[w][c] * ex2_approx(gw - G);
    }
    colpart[((size_t)t.b * C + gj) * nrt + t.rt] = make_float2(G, s);
    bad |= sum_untrusted(s);
  }
  if (bad) atomicOr(flags + t.b, 1);
}

// rows and columns: total = sum_p s_p 2^(g_p - G);  outputs rml = G (log2 units), rmul = score / total.
// Instances flagged by pass 1 (logit range beyond what one reference per warp can carry) are recomputed here
// exactly, max-subtracting like torch.softmax, one thread per row / column: slow, but only ever taken for
// pathological logits, and it keeps the common path at one launch.
__global__ void __launch_bounds__(256)
k_fine_stats_merge(const float* __restrict__ atten, const float2* __restrict__ rowpart,
                   const float2* __restrict__ colpart, int R, int C, int ld, int nstrip, int nrt,
                   const float* __restrict__ score1, int ld1, const float* __restrict__ score2, int ld2,
                   const int* __restrict__ flags, float* __restrict__ rml, float* __restrict__ rmul,
                   float* __restrict__ cml, float* __restrict__ cmul) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= R + C) return;
  const bool is_row = i < R;
  const int c = i - R;
  float sc = 1.f;
  if (is_row) { if (i > 0 && score1) sc = score1[(size_t)b * ld1 + i - 1]; }
  else if (c > 0 && score2) sc = score2[(size_t)b * ld2 + c - 1];
  float* const ol = is_row ? rml + (size_t)b * R + i : cml + (size_t)b * C + c;
  float* const om = is_row ? rmul + (size_t)b * R + i : cmul + (size_t)b * C + c;
  if (flags[b]) {
    const float* A = atten + (size_t)b * R * ld + (is_row ? (size_t)i * ld : (size_t)c);
    const size_t step = is_row ? 1 : (size_t)ld;
    const int len = is_row ? C : R;
    float mx = -INFINITY;
    for (int k = 0; k < len; ++k) mx = fmaxf(mx, A[k * step]);
    float sum = 0.f;
    for (int k = 0; k < len; ++k) sum += expf(A[k * step] - mx);
    *ol = mx * kL2E;
    *om = sc / sum;
    return;
  }
  const float2* p = is_row ? rowpart + ((size_t)b * R + i) * nstrip : colpart + ((size_t)b * C + c) * nrt;
  const int n = is_row ? nstrip : nrt;
  float G = -INFINITY;
  for (int k = 0; k < n; ++k) G = fmaxf(G, p[k].x);
  float s = 0.f;
  for (int k = 0; k < n; ++k) s += p[k].y * exp2f(p[k].x - G);
  *ol = G;
  *om = sc / s;
}

// ---- fused-statistics path: pass 1 came out of the similarity GEMM's epilogue (similarity_tc.cu, STATS) as
// per-tile partial sums of 2^(v log2e - gref) over the main block.  One launch merges them (4-way unrolled partial
// reads); the LAST CTA of every instance (blockIdx.x == nmerge) adds up the peeled background row / column itself.
__global__ void __launch_bounds__(256)
k_fine_stats_merge_fused(const float* __restrict__ atten, const float* __restrict__ rowpart,
                         const float* __restrict__ colpart, int npr, int npc, float gref, int nmerge, int R, int C,
                         int ld, const float* __restrict__ score1, int ld1, const float* __restrict__ score2, int ld2,
                         float* __restrict__ rml, float* __restrict__ rmul, float* __restrict__ cml,
                         float* __restrict__ cmul) {
  __shared__ float s_part[2][8];
  const int b = blockIdx.y;
  const float* A = atten + (size_t)b * R * ld;
  if ((int)blockIdx.x == nmerge) {   // background row 0 and column 0 (their own exponent sums; score 1.0)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float r0 = 0.f, c0 = 0.f;
    for (int j = threadIdx.x; j < C; j += 256) r0 += ex2_approx(fmaf(A[j], kL2E, -gref));
    for (int i = threadIdx.x; i < R; i += 256) c0 += ex2_approx(fmaf(A[(size_t)i * ld], kL2E, -gref));
    r0 = warp_sum(r0);
    c0 = warp_sum(c0);
    if (lane == 0) { s_part[0][warp] = r0; s_part[1][warp] = c0; }
    __syncthreads();
    if (threadIdx.x < 2) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += s_part[threadIdx.x][w];
      if (threadIdx.x == 0) { rml[(size_t)b * R] = gref; rmul[(size_t)b * R] = 1.f / t; }
      else { cml[(size_t)b * C] = gref; cmul[(size_t)b * C] = 1.f / t; }
    }
    return;
  }
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= R + C) return;
  const bool is_row = i < R;
  const int k = is_row ? i : i - R;
  if (k == 0) return;
  float sc = 1.f;
  if (is_row) { if (score1) sc = score1[(size_t)b * ld1 + k - 1]; }
  else if (score2) sc = score2[(size_t)b * ld2 + k - 1];
  // partial-major layouts: consecutive threads read consecutive words
  const float* p = is_row ? rowpart + (size_t)b * npr * R + k : colpart + (size_t)b * npc * C + k;
  const int n = is_row ? npr : npc;
  const size_t step = is_row ? (size_t)R : (size_t)C;
  float s0 = ex2_approx(fmaf(is_row ? A[(size_t)k * ld] : A[k], kL2E, -gref));   // the background column / row entry
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int q = 0;
  for (; q + 3 < n; q += 4) {
    s0 += p[q * step]; s1 += p[(q + 1) * step]; s2 += p[(q + 2) * step]; s3 += p[(q + 3) * step];
  }
  for (; q < n; ++q) s0 += p[q * step];
  const float s = (s0 + s1) + (s2 + s3);
  if (is_row) { rml[(size_t)b * R + k] = gref; rmul[(size_t)b * R + k] = sc / s; }
  else { cml[(size_t)b * C + k] = gref; cmul[(size_t)b * C + k] = sc / s; }
}

// ------------------------------------------------------------------ passes 2 and 3: shared pieces
struct ColConst2 {
  unsigned long long ncml[F2_CPT / 2];  // packed -cml_j
  unsigned long long cmul[F2_CPT / 2];  // packed  s2_j / cs_j   (pass 3: times w2_j)
};
template <bool VEC>
__device__ __forceinline__ void load_col_consts2(ColConst2& kc, const F2Tile& t, int C, const float* __restrict__ cml,
                                                 const float* __restrict__ cmul, const float* __restrict__ w2) {
#pragma unroll
  for (int kk = 0; kk < F2_CPT / 2; ++kk) {
    float l[2], m[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int lc = f2_lcol<VEC>(t.lane, 2 * kk + h);
      const bool ok = lc < t.nv;
      const size_t o = (size_t)t.b * C + t.c0 + lc;
      l[h] = ok ? -cml[o] : 0.f;
      m[h] = ok ? cmul[o] : 0.f;
      if (w2 && ok) m[h] *= w2[(size_t)t.b * (C - 1) + t.c0 + lc - 1];
    }
    kc.ncml[kk] = pack2(l[0], l[1]);
    kc.cmul[kk] = pack2(m[0], m[1]);
  }
}

// e_k = 2^(2 v_k log2e - rml - cml_k) for the 8 columns of one row, as 4 packed pairs
__device__ __forceinline__ void row_exps(const float (&v)[F2_CPT], float nrml, const ColConst2& kc,
                                         unsigned long long (&e)[F2_CPT / 2]) {
  const unsigned long long L2 = pack2(2.f * kL2E, 2.f * kL2E), NR = pack2(nrml, nrml);
#pragma unroll
  for (int kk = 0; kk < F2_CPT / 2; ++kk) {
    float x0, x1;
    unpack2(fma2(pack2(v[2 * kk], v[2 * kk + 1]), L2, add2(NR, kc.ncml[kk])), x0, x1);
    e[kk] = pack2(ex2_approx(x0), ex2_approx(x1));
  }
}

// ------------------------------------------------------------------ pass 2: background-vs-foreground tests
//   w1[i-1] = (argmax_j S[i][:] > 0)  <=>  max_{j>=1} S[i][j] > S[i][0]   (torch.max keeps the first maximum)
template <bool CHECK, bool STRIP0, bool VEC>
__device__ __forceinline__ void labels_body(const float* __restrict__ A, int R, int ld, int nstrip, const F2Tile& t,
                                            const ColConst2& kc, float cml0, float cmul0,
                                            const float* __restrict__ rml, const float* __restrict__ rmul,
                                            float* __restrict__ rowpm, float* __restrict__ ai0,
                                            float (&cmx)[F2_CPT]) {
  const size_t pitch8 = (size_t)8 * ld;
  auto fetch = [&](RowPair& v, float& z0a, float& z0b, int p) {
    const int ra = t.i0 + 16 * p;
    const bool oka = !CHECK || ra < R, okb = !CHECK || ra + 8 < R;
    load_pair<CHECK, VEC>(A + (size_t)ra * ld + t.c0, pitch8, t.nv, t.lane, oka, okb, v);
    if (STRIP0 && t.lane == 0) {
      z0a = oka ? A[(size_t)ra * ld] : 0.f;
      z0b = okb ? A[(size_t)(ra + 8) * ld] : 0.f;
    }
  };
  auto comp = [&](const RowPair& cur, float c0a, float c0b, int p) {
    const int ra = t.i0 + 16 * p, rb = ra + 8;
    if (CHECK && ra >= R) return;
    const bool okb = !CHECK || rb < R;
    const size_t oa = (size_t)t.b * R + ra, ob = (size_t)t.b * R + (okb ? rb : ra);
    const float rmla = rml[oa], rmula = rmul[oa], rmlb = rml[ob], rmulb = rmul[ob];
    unsigned long long ea[F2_CPT / 2], eb[F2_CPT / 2];
    row_exps(cur.a, -rmla, kc, ea);
    row_exps(cur.b, -rmlb, kc, eb);
    const unsigned long long MA = pack2(rmula, rmula), MB = pack2(rmulb, rmulb);
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < F2_CPT / 2; ++kk) {
      float a0, a1, b0, b1;
      unpack2(mul2(mul2(ea[kk], MA), kc.cmul[kk]), a0, a1);
      unpack2(mul2(mul2(eb[kk], MB), kc.cmul[kk]), b0, b1);
      if (CHECK) {  // out-of-range entries must not take part in any maximum
        const bool v0 = f2_lcol<VEC>(t.lane, 2 * kk) < t.nv, v1 = f2_lcol<VEC>(t.lane, 2 * kk + 1) < t.nv;
        a0 = v0 ? a0 : -INFINITY; a1 = v1 ? a1 : -INFINITY;
        b0 = (v0 && okb) ? b0 : -INFINITY; b1 = (v1 && okb) ? b1 : -INFINITY;
      }
      ma = fmaxf(ma, fmaxf(a0, a1));
      mb = fmaxf(mb, fmaxf(b0, b1));
      cmx[2 * kk] = fmaxf(cmx[2 * kk], fmaxf(a0, b0));
      cmx[2 * kk + 1] = fmaxf(cmx[2 * kk + 1], fmaxf(a1, b1));
    }
    const float m = reduce_pair(ma, mb, t.lane, F2Max());
    if (t.lane == 0) rowpm[oa * nstrip + t.cs] = m;
    if (t.lane == 16 && okb) rowpm[((size_t)t.b * R + rb) * nstrip + t.cs] = m;
    if (STRIP0 && t.lane == 0) {
      ai0[oa] = (ex2_approx(fmaf(c0a, 2.f * kL2E, -(rmla + cml0))) * rmula) * cmul0;
      if (okb) ai0[(size_t)t.b * R + rb] = (ex2_approx(fmaf(c0b, 2.f * kL2E, -(rmlb + cml0))) * rmulb) * cmul0;
    }
  };
  RowPair va, vb;
  float a0a = 0.f, a0b = 0.f, b0a = 0.f, b0b = 0.f;
  fetch(va, a0a, a0b, 0);
#pragma unroll 1
  for (int p = 0; p < F2_PAIRS; p += 2) {
    fetch(vb, b0a, b0b, p + 1);
    comp(va, a0a, a0b, p);
    if (p + 2 < F2_PAIRS) fetch(va, a0a, a0b, p + 2);
    comp(vb, b0a, b0b, p + 1);
  }
}

// (2 CTAs / SM with the next row pair prefetched in registers beats 3 CTAs without: 59 vs 72 us)
template <bool VEC>
__global__ void __launch_bounds__(F2_THREADS, 2)
k_fine_labels(const float* __restrict__ atten, int R, int C, int ld, int nstrip, int nrt, const float* __restrict__ rml,
              const float* __restrict__ rmul, const float* __restrict__ cml, const float* __restrict__ cmul,
              float* __restrict__ rowpm, float* __restrict__ colpm, float* __restrict__ ai0,
              float* __restrict__ a0j, int flip) {
  __shared__ float s_cm[F2_WARPS][F2_TC];
  const F2Tile t = f2_tile(R, C, flip != 0);
  const float* A = atten + (size_t)t.b * R * ld;
  ColConst2 kc;
  load_col_consts2<VEC>(kc, t, C, cml, cmul, nullptr);
  float cmx[F2_CPT];
#pragma unroll
  for (int k = 0; k < F2_CPT; ++k) cmx[k] = -INFINITY;
  if (t.rt == 0 && t.warp == 0) {  // S[0][j] for this strip's columns (the background row is in no maximum)
    RowPair v;
    load_pair<true, VEC>(A + t.c0, 0, t.nv, t.lane, true, false, v);
    const float rml0 = rml[(size_t)t.b * R], rmul0 = rmul[(size_t)t.b * R];
    unsigned long long e[F2_CPT / 2];
    row_exps(v.a, -rml0, kc, e);
    const unsigned long long M0 = pack2(rmul0, rmul0);
#pragma unroll
    for (int kk = 0; kk < F2_CPT / 2; ++kk) {
      float a0, a1;
      unpack2(mul2(mul2(e[kk], M0), kc.cmul[kk]), a0, a1);
      const int l0 = f2_lcol<VEC>(t.lane, 2 * kk), l1 = f2_lcol<VEC>(t.lane, 2 * kk + 1);
      if (l0 < t.nv) a0j[(size_t)t.b * C + t.c0 + l0] = a0;
      if (l1 < t.nv) a0j[(size_t)t.b * C + t.c0 + l1] = a1;
    }
  }
  const float cml0 = cml[(size_t)t.b * C], cmul0 = cmul[(size_t)t.b * C];
  if (t.cs == 0) {
    if (t.full) labels_body<false, true, VEC>(A, R, ld, nstrip, t, kc, cml0, cmul0, rml, rmul, rowpm, ai0, cmx);
    else labels_body<true, true, VEC>(A, R, ld, nstrip, t, kc, cml0, cmul0, rml, rmul, rowpm, ai0, cmx);
  } else {
    if (t.full) labels_body<false, false, VEC>(A, R, ld, nstrip, t, kc, cml0, cmul0, rml, rmul, rowpm, ai0, cmx);
    else labels_body<true, false, VEC>(A, R, ld, nstrip, t, kc, cml0, cmul0, rml, rmul, rowpm, ai0, cmx);
  }
#pragma unroll
  for (int k = 0; k < F2_CPT; ++k) s_cm[t.warp][f2_lcol<VEC>(t.lane, k)] = cmx[k];
  __syncthreads();
  const int c = threadIdx.x;
  const int gj = t.c0 + c;
  if (gj < C) {
    float m = -INFINITY;
#pragma unroll
    for (int w = 0; w < F2_WARPS; ++w) m = fmaxf(m, s_cm[w][c]);
    colpm[((size_t)t.b * C + gj) * nrt + t.rt] = m;
  }
}

// ------------------------------------------------------------------ pass 3: masked soft correspondences
//   per row i >= 1:  rmul_i * sum_{j>=1} e_ij (cmul_j w2_j) {x_j, y_j, z_j, 1}      (model_utils.py:549-555)
template <bool CHECK, bool VEC>
__device__ __forceinline__ void rows_body(const float* __restrict__ A, int R, int ld, int nstrip, const F2Tile& t,
                                          const ColConst2& kc, const unsigned long long (&px)[F2_CPT / 2],
                                          const unsigned long long (&py)[F2_CPT / 2],
                                          const unsigned long long (&pz)[F2_CPT / 2], const float* __restrict__ rml,
                                          const float* __restrict__ rmul, float4* __restrict__ rowpart4) {
  const size_t pitch8 = (size_t)8 * ld;
  const int N1 = R - 1;
  auto fetch = [&](RowPair& v, int p) {
    const int ra = t.i0 + 16 * p;
    load_pair<CHECK, VEC>(A + (size_t)ra * ld + t.c0, pitch8, t.nv, t.lane, !CHECK || ra < R, !CHECK || ra + 8 < R, v);
  };
  auto comp = [&](const RowPair& cur, int p) {
    const int ra = t.i0 + 16 * p, rb = ra + 8;
    if (CHECK && ra >= R) return;
    const bool okb = !CHECK || rb < R;
    const size_t oa = (size_t)t.b * R + ra, ob = (size_t)t.b * R + (okb ? rb : ra);
    const float rmla = rml[oa], rmula = rmul[oa], rmlb = rml[ob], rmulb = rmul[ob];
    unsigned long long ea[F2_CPT / 2], eb[F2_CPT / 2];
    row_exps(cur.a, -rmla, kc, ea);
    row_exps(cur.b, -rmlb, kc, eb);
    const unsigned long long Z = pack2(0.f, 0.f);
    unsigned long long ax = Z, ay = Z, az = Z, aw = Z, bx = Z, by = Z, bz = Z, bw = Z;
#pragma unroll
    for (int kk = 0; kk < F2_CPT / 2; ++kk) {
      const unsigned long long ta = mul2(ea[kk], kc.cmul[kk]), tb = mul2(eb[kk], kc.cmul[kk]);
      ax = fma2(ta, px[kk], ax); ay = fma2(ta, py[kk], ay); az = fma2(ta, pz[kk], az); aw = add2(aw, ta);
      bx = fma2(tb, px[kk], bx); by = fma2(tb, py[kk], by); bz = fma2(tb, pz[kk], bz); bw = add2(bw, tb);
    }
    float acc[8], u0, u1;
    unpack2(ax, u0, u1); acc[0] = u0 + u1;
    unpack2(ay, u0, u1); acc[1] = u0 + u1;
    unpack2(az, u0, u1); acc[2] = u0 + u1;
    unpack2(aw, u0, u1); acc[3] = u0 + u1;
    unpack2(bx, u0, u1); acc[4] = u0 + u1;
    unpack2(by, u0, u1); acc[5] = u0 + u1;
    unpack2(bz, u0, u1); acc[6] = u0 + u1;
    unpack2(bw, u0, u1); acc[7] = u0 + u1;
    // 8 values across 32 lanes: halving exchanges (4 + 2 + 1 shuffles) then two xor steps;
    // afterwards lane group (lane >> 2) holds component (lane >> 2): 0..3 = row a {x,y,z,w}, 4..7 = row b
    {
      int off = 16;
#pragma unroll
      for (int n = 8; n > 1; n >>= 1, off >>= 1) {
        const bool upper = (t.lane & off) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
          const float send = upper ? acc[i] : acc[i + n / 2];
          const float keep = upper ? acc[i + n / 2] : acc[i];
          acc[i] = keep + __shfl_xor_sync(kFull, send, off);
        }
      }
      acc[0] += __shfl_xor_sync(kFull, acc[0], 2);
      acc[0] += __shfl_xor_sync(kFull, acc[0], 1);
    }
    if ((t.lane & 3) == 0) {
      const int cmp = t.lane >> 2;
      const bool isb = cmp >= 4;
      if (!isb || okb) {
        float* dst = reinterpret_cast<float*>(rowpart4 + ((size_t)t.b * N1 + (isb ? rb : ra) - 1) * nstrip + t.cs) +
                     (cmp & 3);
        *dst = acc[0] * (isb ? rmulb : rmula);
      }
    }
  };
  RowPair va, vb;
  fetch(va, 0);
#pragma unroll 1
  for (int p = 0; p < F2_PAIRS; p += 2) {
    fetch(vb, p + 1);
    comp(va, p);
    if (p + 2 < F2_PAIRS) fetch(va, p + 2);
    comp(vb, p + 1);
  }
}

template <bool VEC>
__global__ void __launch_bounds__(F2_THREADS, 2)
k_fine_rows(const float* __restrict__ atten, int R, int C, int ld, int nstrip, const float* __restrict__ rml,
            const float* __restrict__ rmul, const float* __restrict__ cml, const float* __restrict__ cmul,
            const float* __restrict__ w2, const float* __restrict__ pts2, float4* __restrict__ rowpart4) {
  const F2Tile t = f2_tile(R, C, false);
  const float* A = atten + (size_t)t.b * R * ld;
  ColConst2 kc;
  load_col_consts2<VEC>(kc, t, C, cml, cmul, w2);
  unsigned long long px[F2_CPT / 2], py[F2_CPT / 2], pz[F2_CPT / 2];
#pragma unroll
  for (int kk = 0; kk < F2_CPT / 2; ++kk) {
    float x[2], y[2], z[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int lc = f2_lcol<VEC>(t.lane, 2 * kk + h);
      const bool ok = lc < t.nv;
      const float* p = pts2 + ((size_t)t.b * (C - 1) + (ok ? t.c0 + lc - 1 : 0)) * 3;
      x[h] = p[0]; y[h] = p[1]; z[h] = p[2];
    }
    px[kk] = pack2(x[0], x[1]); py[kk] = pack2(y[0], y[1]); pz[kk] = pack2(z[0], z[1]);
  }
  if (t.full) rows_body<false, VEC>(A, R, ld, nstrip, t, kc, px, py, pz, rml, rmul, rowpart4);
  else rows_body<true, VEC>(A, R, ld, nstrip, t, kc, px, py, pz, rml, rmul, rowpart4);
}

// ------------------------------------------------------------------ passes 2 and 3 fed by TMA (round 2)
// The register-streaming kernels above top out at 50-55 % of the HBM copy bandwidth: a CTA lives for one 128 x 256 tile
// (128 KB, ~6 us), and its prologue (column constants, first row pair) and epilogue are only covered by the ONE other CTA
// of the SM.  With the pitched layout the main block is a TMA-addressable plane, so these variants are persistent:
// 2 CTAs per SM walk the tile list (strips fastest: the CTAs running side by side read whole rows);
//   * a producer thread keeps a ring of 16-row x 256-column boxes (16 KB each) in flight ACROSS tile boundaries
//     (cp.async.bulk.tensor, mbarrier complete_tx);
//   * eight consumer warps take one row pair per stage out of shared memory (LDS.128, the lane -> column mapping of the
//     VEC loads, so every partial is bit-identical to the kernels above) and hand the slot back as soon as the pair
//     sits in registers;
//   * a constants warp runs one tile ahead: it stages the tile's column / row constants in shared memory (double
//     buffered, named-barrier hand-off) and computes the two border vectors of pass 2 (S[i][0], S[0][j]) itself, so the
//     consumers never wait for a global load.
// Full tiles only ((R - 1) % 128 == 0, (C - 1) % 256 == 0: the UNOPose fine shape); everything else takes the kernels above.
constexpr int FT_ROWS = 16;                          // rows per stage: one pair per consumer warp
constexpr int FT_STAGE_FLOATS = FT_ROWS * F2_TC;     // 16 KB
constexpr int FT_SPT = F2_RT / FT_ROWS;              // 8 stages per tile
constexpr int FT_THREADS = F2_THREADS + 64;          // + the producer warp + the constants warp
static_assert(FT_ROWS == 2 * F2_WARPS, "one row pair per consumer warp and stage");

// per-tile constants: NCOL column arrays of F2_TC floats, then rml | rmul of the tile's F2_RT rows
template <int STAGES, int NCOL, bool CM>
struct FtSmem {
  alignas(128) float buf[STAGES][FT_STAGE_FLOATS];
  alignas(16) float cst[2][NCOL * F2_TC + 2 * F2_RT];
  float cm[CM ? 2 : 1][CM ? F2_WARPS : 1][CM ? F2_TC : 1];
  unsigned long long full[STAGES], empty[STAGES];
};

struct FtTile { int b, rt, cs; };
__device__ __forceinline__ FtTile ft_tile(int T, int nb, int nrt, int nstrip, bool flip) {
  FtTile t;
  const int per = nrt * nstrip;
  const int bb = T / per, rem = T - bb * per;
  t.b = flip ? nb - 1 - bb : bb;
  t.rt = rem / nstrip;
  t.cs = rem - t.rt * nstrip;
  return t;
}

// keep_lo: instances b < keep_lo are what the NEXT pass reads first (it walks the instances upwards): their lines get
// evict_last, everything else evict_first (read once by this kernel).  keep_lo < 0: no hints.
template <int STAGES, class Smem>
__device__ __forceinline__ void ft_producer(const CUtensorMap* map, Smem& sm, int ntile, int nb, int nrt, int nstrip,
                                            bool flip, int keep_lo) {
  int s = 0;
  uint32_t ph = 0;
  const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
  for (int T = blockIdx.x; T < ntile; T += gridDim.x) {
    const FtTile t = ft_tile(T, nb, nrt, nstrip, flip);
    for (int q = 0; q < FT_SPT; ++q) {
      mbar_wait(&sm.empty[s], ph ^ 1);
      mbar_expect_tx(&sm.full[s], FT_STAGE_FLOATS * sizeof(float));
      if (keep_lo >= 0)
        tma_load_3d_hint(map, &sm.full[s], sm.buf[s], t.cs * F2_TC, t.rt * F2_RT + q * FT_ROWS, t.b,
                         t.b < keep_lo ? pol_last : pol_first);
      else
        tma_load_3d(map, &sm.full[s], sm.buf[s], t.cs * F2_TC, t.rt * F2_RT + q * FT_ROWS, t.b);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  }
}

// this warp's row pair of stage `s` out of shared memory; the slot is released once every lane holds its values
template <class Smem>
__device__ __forceinline__ void ft_take_pair(Smem& sm, int s, uint32_t ph, int warp, int lane, RowPair& v) {
  mbar_wait(&sm.full[s], ph);
  const float* ra = sm.buf[s] + (2 * warp) * F2_TC + 4 * lane;
  const float4 x0 = *reinterpret_cast<const float4*>(ra), x1 = *reinterpret_cast<const float4*>(ra + 128);
  const float4 y0 = *reinterpret_cast<const float4*>(ra + F2_TC), y1 = *reinterpret_cast<const float4*>(ra + F2_TC + 128);
  v.a[0] = x0.x; v.a[1] = x0.y; v.a[2] = x0.z; v.a[3] = x0.w; v.a[4] = x1.x; v.a[5] = x1.y; v.a[6] = x1.z; v.a[7] = x1.w;
  v.b[0] = y0.x; v.b[1] = y0.y; v.b[2] = y0.z; v.b[3] = y0.w; v.b[4] = y1.x; v.b[5] = y1.y; v.b[6] = y1.z; v.b[7] = y1.w;
  __syncwarp();
  if (lane == 0) mbar_arrive(&sm.empty[s]);
}

template <int STAGES, class Smem>
__device__ __forceinline__ void ft_init(Smem& sm) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], F2_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

// Hand-off of the per-tile constants buffers between the constants warp and the eight consumer warps: named barriers
// (bar.arrive / bar.sync over the 288 participating threads), ids FT_BAR_FULL + buffer and FT_BAR_EMPTY + buffer.  The
// constants warp arrives on `full` after its stores and syncs on `empty` before it reuses a buffer (from the third tile
// on); the consumers sync on `full`, read, and arrive on `empty`.  (An mbarrier pair did the same job; named barriers are
// what compute-sanitizer's racecheck models for thread-issued shared-memory accesses.)
constexpr int FT_BAR_CM = 1, FT_BAR_FULL = 2, FT_BAR_EMPTY = 4;
constexpr int FT_CST_THREADS = F2_THREADS + 32;
template <int BASE>   // immediate barrier ids (a register id makes ptxas reserve all 16 barriers)
__device__ __forceinline__ void ft_bar_sync(int buf) {
  if (buf) asm volatile("bar.sync %0, %1;" ::"n"(BASE + 1), "n"(FT_CST_THREADS) : "memory");
  else asm volatile("bar.sync %0, %1;" ::"n"(BASE), "n"(FT_CST_THREADS) : "memory");
}
template <int BASE>
__device__ __forceinline__ void ft_bar_arrive(int buf) {
  if (buf) asm volatile("bar.arrive %0, %1;" ::"n"(BASE + 1), "n"(FT_CST_THREADS) : "memory");
  else asm volatile("bar.arrive %0, %1;" ::"n"(BASE), "n"(FT_CST_THREADS) : "memory");
}

// a lane's 8 values of column array `a` (the VEC lane -> column mapping) as 4 packed pairs
__device__ __forceinline__ void ft_col8(const float* a, int lane, unsigned long long (&o)[F2_CPT / 2]) {
  const float4 lo = *reinterpret_cast<const float4*>(a + 4 * lane), hi = *reinterpret_cast<const float4*>(a + 128 + 4 * lane);
  o[0] = pack2(lo.x, lo.y); o[1] = pack2(lo.z, lo.w); o[2] = pack2(hi.x, hi.y); o[3] = pack2(hi.z, hi.w);
}
// lane q < 16 holds the constants of the warp's q-th row of a tile (stage q >> 1, half q & 1): its tile-local row
__device__ __forceinline__ int ft_lane_lrow(int warp, int lane) {
  const int q = lane & 15;
  return 2 * warp + FT_ROWS * (q >> 1) + (q & 1);
}

constexpr int FT_LABEL_STAGES = 5;
constexpr int FT_ROWS_STAGES = 6;

__global__ void __launch_bounds__(FT_THREADS, 2)
k_fine_labels_tma(const __grid_constant__ CUtensorMap map, const float* __restrict__ atten, int nb, int R, int C, int ld,
                  int nstrip, int nrt, const float* __restrict__ rml, const float* __restrict__ rmul,
                  const float* __restrict__ cml, const float* __restrict__ cmul, float* __restrict__ rowpm,
                  float* __restrict__ colpm, float* __restrict__ ai0, float* __restrict__ a0j, int flip, int l2_keep) {
  extern __shared__ unsigned char ft_raw[];
  using Smem = FtSmem<FT_LABEL_STAGES, 2, true>;
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(ft_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = nb * nrt * nstrip;
  ft_init<FT_LABEL_STAGES>(sm);
  if (warp == F2_WARPS) {
    if (lane == 0) ft_producer<FT_LABEL_STAGES>(&map, sm, ntile, nb, nrt, nstrip, flip != 0, l2_keep > 0 ? l2_keep : -1);
    return;
  }
  if (warp == F2_WARPS + 1) {
    // ===== constants warp: cst = -cml[256] | cmul[256] | rml[128] | rmul[128]; border vectors straight to global
    int it = 0;
    for (int T = blockIdx.x; T < ntile; T += gridDim.x, ++it) {
      const FtTile t = ft_tile(T, nb, nrt, nstrip, flip != 0);
      const size_t oc = (size_t)t.b * C + 1 + t.cs * F2_TC, orw = (size_t)t.b * R + 1 + t.rt * F2_RT;
      float l[F2_CPT], m[F2_CPT], rl[F2_RT / 32], rm[F2_RT / 32];
#pragma unroll
      for (int k = 0; k < F2_CPT; ++k) { l[k] = -cml[oc + lane + 32 * k]; m[k] = cmul[oc + lane + 32 * k]; }
#pragma unroll
      for (int k = 0; k < F2_RT / 32; ++k) { rl[k] = rml[orw + lane + 32 * k]; rm[k] = rmul[orw + lane + 32 * k]; }
      if (it >= 2) ft_bar_sync<FT_BAR_EMPTY>(it & 1);   // the consumers are done with this buffer's previous tile
      float* c = sm.cst[it & 1];
#pragma unroll
      for (int k = 0; k < F2_CPT; ++k) { c[lane + 32 * k] = l[k]; c[F2_TC + lane + 32 * k] = m[k]; }
#pragma unroll
      for (int k = 0; k < F2_RT / 32; ++k) { c[2 * F2_TC + lane + 32 * k] = rl[k]; c[2 * F2_TC + F2_RT + lane + 32 * k] = rm[k]; }
      ft_bar_arrive<FT_BAR_FULL>(it & 1);
      const float* A = atten + (size_t)t.b * R * ld;
      if (t.cs == 0) {   // S[i][0] of the tile's rows (the background column is in no box)
        const float cml0 = cml[(size_t)t.b * C], cmul0 = cmul[(size_t)t.b * C];
#pragma unroll
        for (int k = 0; k < F2_RT / 32; ++k) {
          const float v0 = A[(size_t)(1 + t.rt * F2_RT + lane + 32 * k) * ld];
          ai0[orw + lane + 32 * k] = (ex2_approx(fmaf(v0, 2.f * kL2E, -(rl[k] + cml0))) * rm[k]) * cmul0;
        }
      }
      if (t.rt == 0) {   // S[0][j] of the strip's columns (the background row is in no maximum)
        const float rml0 = rml[(size_t)t.b * R], rmul0 = rmul[(size_t)t.b * R];
        const unsigned long long L2 = pack2(2.f * kL2E, 2.f * kL2E), NR = pack2(-rml0, -rml0), M0 = pack2(rmul0, rmul0);
#pragma unroll
        for (int k = 0; k < F2_CPT; k += 2) {
          // the arithmetic of row_exps / labels_body on the pair (k, k + 1), whatever columns the pair holds
          const float va = A[1 + t.cs * F2_TC + lane + 32 * k], vb = A[1 + t.cs * F2_TC + lane + 32 * (k + 1)];
          float x0, x1, a0, a1;
          unpack2(fma2(pack2(va, vb), L2, add2(NR, pack2(l[k], l[k + 1]))), x0, x1);
          unpack2(mul2(mul2(pack2(ex2_approx(x0), ex2_approx(x1)), M0), pack2(m[k], m[k + 1])), a0, a1);
          a0j[oc + lane + 32 * k] = a0;
          a0j[oc + lane + 32 * (k + 1)] = a1;
        }
      }
    }
    return;
  }
  int s = 0, it = 0;
  uint32_t ph = 0;
  for (int T = blockIdx.x; T < ntile; T += gridDim.x, ++it) {
    const FtTile t = ft_tile(T, nb, nrt, nstrip, flip != 0);
    ColConst2 kc;
    ft_bar_sync<FT_BAR_FULL>(it & 1);
    const float* c = sm.cst[it & 1];
    ft_col8(c, lane, kc.ncml);
    ft_col8(c + F2_TC, lane, kc.cmul);
    const int lrow = ft_lane_lrow(warp, lane);
    const float my_rml = c[2 * F2_TC + lrow], my_rmul = c[2 * F2_TC + F2_RT + lrow];
    ft_bar_arrive<FT_BAR_EMPTY>(it & 1);
    float cmx[F2_CPT];
#pragma unroll
    for (int k = 0; k < F2_CPT; ++k) cmx[k] = -INFINITY;
#pragma unroll 1
    for (int st = 0; st < FT_SPT; ++st) {
      RowPair cur;
      ft_take_pair(sm, s, ph, warp, lane, cur);
      if (++s == FT_LABEL_STAGES) { s = 0; ph ^= 1; }
      const float rmla = __shfl_sync(kFull, my_rml, 2 * st), rmula = __shfl_sync(kFull, my_rmul, 2 * st);
      const float rmlb = __shfl_sync(kFull, my_rml, 2 * st + 1), rmulb = __shfl_sync(kFull, my_rmul, 2 * st + 1);
      unsigned long long ea[F2_CPT / 2], eb[F2_CPT / 2];
      row_exps(cur.a, -rmla, kc, ea);
      row_exps(cur.b, -rmlb, kc, eb);
      const unsigned long long MA = pack2(rmula, rmula), MB = pack2(rmulb, rmulb);
      float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
      for (int kk = 0; kk < F2_CPT / 2; ++kk) {
        float a0, a1, b0, b1;
        unpack2(mul2(mul2(ea[kk], MA), kc.cmul[kk]), a0, a1);
        unpack2(mul2(mul2(eb[kk], MB), kc.cmul[kk]), b0, b1);
        ma = fmaxf(ma, fmaxf(a0, a1));
        mb = fmaxf(mb, fmaxf(b0, b1));
        cmx[2 * kk] = fmaxf(cmx[2 * kk], fmaxf(a0, b0));
        cmx[2 * kk + 1] = fmaxf(cmx[2 * kk + 1], fmaxf(a1, b1));
      }
      const float m = reduce_pair(ma, mb, lane, F2Max());
      const size_t oa = (size_t)t.b * R + 1 + t.rt * F2_RT + 2 * warp + FT_ROWS * st;
      if (lane == 0) rowpm[oa * nstrip + t.cs] = m;
      if (lane == 16) rowpm[(oa + 1) * nstrip + t.cs] = m;
    }
    // column maxima of the tile: combine the 8 warps (double-buffered: one named barrier per tile)
    float (*cm)[F2_TC] = sm.cm[it & 1];
#pragma unroll
    for (int k = 0; k < F2_CPT; ++k) cm[warp][f2_lcol<true>(lane, k)] = cmx[k];
    asm volatile("bar.sync %0, %1;" ::"n"(FT_BAR_CM), "n"(F2_THREADS) : "memory");
    {
      const int cc = threadIdx.x;
      float m = -INFINITY;
#pragma unroll
      for (int w = 0; w < F2_WARPS; ++w) m = fmaxf(m, cm[w][cc]);
      colpm[((size_t)t.b * C + 1 + t.cs * F2_TC + cc) * nrt + t.rt] = m;
    }
  }
}

__global__ void __launch_bounds__(FT_THREADS, 2)
k_fine_rows_tma(const __grid_constant__ CUtensorMap map, int nb, int R, int C, int nstrip, int nrt,
                const float* __restrict__ rml, const float* __restrict__ rmul, const float* __restrict__ cml,
                const float* __restrict__ cmul, const float* __restrict__ w2, const float* __restrict__ pts2,
                float4* __restrict__ rowpart4, int l2_keep) {
  extern __shared__ unsigned char ft_raw[];
  using Smem = FtSmem<FT_ROWS_STAGES, 5, false>;
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(ft_raw) + 127) & ~(uintptr_t)127);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntile = nb * nrt * nstrip;
  const int N1 = R - 1;
  ft_init<FT_ROWS_STAGES>(sm);
  if (warp == F2_WARPS) {
    if (lane == 0) ft_producer<FT_ROWS_STAGES>(&map, sm, ntile, nb, nrt, nstrip, false, l2_keep > 0 ? 0 : -1);   // read once
    return;
  }
  if (warp == F2_WARPS + 1) {
    // ===== constants warp: cst = -cml | cmul * w2 | x | y | z (256 each) | rml[128] | rmul[128]
    int it = 0;
    for (int T = blockIdx.x; T < ntile; T += gridDim.x, ++it) {
      const FtTile t = ft_tile(T, nb, nrt, nstrip, false);
      const size_t oc = (size_t)t.b * C + 1 + t.cs * F2_TC, orw = (size_t)t.b * R + 1 + t.rt * F2_RT;
      const size_t ow = (size_t)t.b * (C - 1) + t.cs * F2_TC;
      float l[F2_CPT], m[F2_CPT], rl[F2_RT / 32], rm[F2_RT / 32], xyz[3 * F2_CPT];
#pragma unroll
      for (int k = 0; k < F2_CPT; ++k) {
        l[k] = -cml[oc + lane + 32 * k];
        m[k] = cmul[oc + lane + 32 * k];
        m[k] *= w2[ow + lane + 32 * k];
      }
#pragma unroll
      for (int k = 0; k < 3 * F2_CPT; ++k) xyz[k] = pts2[ow * 3 + lane + 32 * k];   // 768 consecutive floats, AoS
#pragma unroll
      for (int k = 0; k < F2_RT / 32; ++k) { rl[k] = rml[orw + lane + 32 * k]; rm[k] = rmul[orw + lane + 32 * k]; }
      if (it >= 2) ft_bar_sync<FT_BAR_EMPTY>(it & 1);   // the consumers are done with this buffer's previous tile
      float* c = sm.cst[it & 1];
#pragma unroll
      for (int k = 0; k < F2_CPT; ++k) { c[lane + 32 * k] = l[k]; c[F2_TC + lane + 32 * k] = m[k]; }
#pragma unroll
      for (int k = 0; k < 3 * F2_CPT; ++k) {
        const int e = lane + 32 * k, j = e / 3, a = e - 3 * j;   // AoS element e -> column j, component a
        c[(2 + a) * F2_TC + j] = xyz[k];
      }
#pragma unroll
      for (int k = 0; k < F2_RT / 32; ++k) { c[5 * F2_TC + lane + 32 * k] = rl[k]; c[5 * F2_TC + F2_RT + lane + 32 * k] = rm[k]; }
      ft_bar_arrive<FT_BAR_FULL>(it & 1);
    }
    return;
  }
  int s = 0, it = 0;
  uint32_t ph = 0;
  for (int T = blockIdx.x; T < ntile; T += gridDim.x, ++it) {
    const FtTile t = ft_tile(T, nb, nrt, nstrip, false);
    ColConst2 kc;
    unsigned long long px[F2_CPT / 2], py[F2_CPT / 2], pz[F2_CPT / 2];
    ft_bar_sync<FT_BAR_FULL>(it & 1);
    const float* c = sm.cst[it & 1];
    ft_col8(c, lane, kc.ncml);
    ft_col8(c + F2_TC, lane, kc.cmul);
    ft_col8(c + 2 * F2_TC, lane, px);
    ft_col8(c + 3 * F2_TC, lane, py);
    ft_col8(c + 4 * F2_TC, lane, pz);
    const int lrow = ft_lane_lrow(warp, lane);
    const float my_rml = c[5 * F2_TC + lrow], my_rmul = c[5 * F2_TC + F2_RT + lrow];
    ft_bar_arrive<FT_BAR_EMPTY>(it & 1);
#pragma unroll 1
    for (int st = 0; st < FT_SPT; ++st) {
      RowPair cur;
      ft_take_pair(sm, s, ph, warp, lane, cur);
      if (++s == FT_ROWS_STAGES) { s = 0; ph ^= 1; }
      const float rmla = __shfl_sync(kFull, my_rml, 2 * st), rmula = __shfl_sync(kFull, my_rmul, 2 * st);
      const float rmlb = __shfl_sync(kFull, my_rml, 2 * st + 1), rmulb = __shfl_sync(kFull, my_rmul, 2 * st + 1);
      unsigned long long ea[F2_CPT / 2], eb[F2_CPT / 2];
      row_exps(cur.a, -rmla, kc, ea);
      row_exps(cur.b, -rmlb, kc, eb);
      const unsigned long long Z = pack2(0.f, 0.f);
      unsigned long long ax = Z, ay = Z, az = Z, aw = Z, bx = Z, by = Z, bz = Z, bw = Z;
#pragma unroll
      for (int kk = 0; kk < F2_CPT / 2; ++kk) {
        const unsigned long long ta = mul2(ea[kk], kc.cmul[kk]), tb = mul2(eb[kk], kc.cmul[kk]);
        ax = fma2(ta, px[kk], ax); ay = fma2(ta, py[kk], ay); az = fma2(ta, pz[kk], az); aw = add2(aw, ta);
        bx = fma2(tb, px[kk], bx); by = fma2(tb, py[kk], by); bz = fma2(tb, pz[kk], bz); bw = add2(bw, tb);
      }
      float acc[8], u0, u1;
      unpack2(ax, u0, u1); acc[0] = u0 + u1;
      unpack2(ay, u0, u1); acc[1] = u0 + u1;
      unpack2(az, u0, u1); acc[2] = u0 + u1;
      unpack2(aw, u0, u1); acc[3] = u0 + u1;
      unpack2(bx, u0, u1); acc[4] = u0 + u1;
      unpack2(by, u0, u1); acc[5] = u0 + u1;
      unpack2(bz, u0, u1); acc[6] = u0 + u1;
      unpack2(bw, u0, u1); acc[7] = u0 + u1;
      {   // same exchange tree as rows_body: afterwards lane group (lane >> 2) holds component (lane >> 2)
        int off = 16;
#pragma unroll
        for (int n = 8; n > 1; n >>= 1, off >>= 1) {
          const bool upper = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const float send = upper ? acc[i] : acc[i + n / 2];
            const float keep = upper ? acc[i + n / 2] : acc[i];
            acc[i] = keep + __shfl_xor_sync(kFull, send, off);
          }
        }
        acc[0] += __shfl_xor_sync(kFull, acc[0], 2);
        acc[0] += __shfl_xor_sync(kFull, acc[0], 1);
      }
      if ((lane & 3) == 0) {
        const int cmp = lane >> 2;
        const bool isb = cmp >= 4;
        const int row = 1 + t.rt * F2_RT + 2 * warp + FT_ROWS * st + (isb ? 1 : 0);
        float* dst = reinterpret_cast<float*>(rowpart4 + ((size_t)t.b * N1 + row - 1) * nstrip + t.cs) + (cmp & 3);
        *dst = acc[0] * (isb ? rmulb : rmula);
      }
    }
  }
}

// TMA-fed passes: pitched 128-bit-aligned layout, whole tiles only.  UPK_FINE_TMA=0 keeps the register-streaming kernels.
static bool fine_tma_ok(const float* atten, int ld, int R, int C) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("UPK_FINE_TMA"); enabled = e ? atoi(e) : 1; }
  return enabled && tc_tma_available() && R > 1 && C > 1 && (R - 1) % F2_RT == 0 && (C - 1) % F2_TC == 0 && ld % 4 == 0 &&
         ((reinterpret_cast<uintptr_t>(atten + ld + 1)) & 15) == 0;
}

static int fine_tma_grid(int ntile) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return ntile < 2 * sms ? ntile : 2 * sms;
}

static int launch_fine_labels_tma(const float* atten, int ld, int b, int R, int C, const FineGeom2& f, const AssignWs& ws,
                                  int flip, cudaStream_t st) {
  CUtensorMap map;
  int rc = tc_make_map_plane(&map, atten + ld + 1, b, R - 1, C - 1, ld, (size_t)R * ld, F2_TC, FT_ROWS);
  if (rc) return rc;
  const size_t smem = sizeof(FtSmem<FT_LABEL_STAGES, 2, true>) + 128;
  UPK_CUDA_TRY(cudaFuncSetAttribute(k_fine_labels_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_fine_labels_tma<<<fine_tma_grid(b * f.nrt * f.nstrip), FT_THREADS, smem, st>>>(
      map, atten, b, R, C, ld, f.nstrip, f.nrt, ws.rmax, ws.rsum, ws.cmax, ws.csum, ws.rowpm, ws.colpm, ws.ai0, ws.a0j, flip,
      flip ? fine_l2_keep(b, R, C) : 0);   // only behind the statistics-fused GEMM (descending walk)
  UPK_RETURN_LAST_ERROR();
}

// ------------------------------------------------------------------ host side
FineGeom2 fine_geom2(int R, int C) {
  FineGeom2 g;
  g.nstrip = ceil_div(C - 1, F2_TC);
  g.nrt = ceil_div(R - 1, F2_RT);
  return g;
}

// 128-bit loads need: pitch a multiple of 4 floats and element (i, 1) of every row 16-byte aligned
static bool fine_vec_ok(const float* atten, int ld) {
  return ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(atten + 1)) & 15) == 0;
}

int run_fine_labels2(const float* atten, int ld, const float* score1, int ld1, const float* score2, int ld2, int b,
                     const AssignGeom& g, const AssignWs& ws, float* w1, float* w2, cudaStream_t st) {
  const FineGeom2 f = fine_geom2(g.R, g.C);
  const dim3 grid(f.nstrip, f.nrt, b);
  const dim3 mg(ceil_div(g.R + g.C, 256), b);
  const bool vec = fine_vec_ok(atten, ld);
  UPK_CUDA_TRY(cudaMemsetAsync(ws.flags, 0, sizeof(int) * (size_t)b, st));
  if (vec) k_fine_stats<true><<<grid, F2_THREADS, 0, st>>>(atten, g.R, g.C, ld, f.nstrip, f.nrt, ws.rowpart, ws.colpart, ws.flags);
  else k_fine_stats<false><<<grid, F2_THREADS, 0, st>>>(atten, g.R, g.C, ld, f.nstrip, f.nrt, ws.rowpart, ws.colpart, ws.flags);
  k_fine_stats_merge<<<mg, 256, 0, st>>>(atten, ws.rowpart, ws.colpart, g.R, g.C, ld, f.nstrip, f.nrt, score1, ld1,
                                         score2, ld2, ws.flags, ws.rmax, ws.rsum, ws.cmax, ws.csum);
  if (fine_tma_ok(atten, ld, g.R, g.C)) {
    int rc = launch_fine_labels_tma(atten, ld, b, g.R, g.C, f, ws, 0, st);
    if (rc) return rc;
  } else if (vec) k_fine_labels<true><<<grid, F2_THREADS, 0, st>>>(atten, g.R, g.C, ld, f.nstrip, f.nrt, ws.rmax, ws.rsum, ws.cmax, ws.csum,
                                                            ws.rowpm, ws.colpm, ws.ai0, ws.a0j, 0);
  else k_fine_labels<false><<<grid, F2_THREADS, 0, st>>>(atten, g.R, g.C, ld, f.nstrip, f.nrt, ws.rmax, ws.rsum, ws.cmax, ws.csum,
                                                             ws.rowpm, ws.colpm, ws.ai0, ws.a0j, 0);
  count_launch(3);
  return launch_labels_merge(ws.rowpm, ws.colpm, ws.ai0, ws.a0j, b, g.R, g.C, f.nrt, f.nstrip, w1, w2, st);
}

int run_fine_labels2_fused(const float* atten, int ld, const float* stats, float temp, const float* score1, int ld1,
                           const float* score2, int ld2, int b, const AssignGeom& g, const AssignWs& ws, float* w1,
                           float* w2, cudaStream_t st) {
  const FineGeom2 f = fine_geom2(g.R, g.C);
  const SimStatsGeom sg = sim_stats_geom(b, g.R, g.C);
  const float gref = sim_stats_gref(temp);
  const dim3 grid(f.nstrip, f.nrt, b);
  const int nmerge = ceil_div(g.R + g.C, 256);
  const dim3 mg(nmerge + 1, b);
  k_fine_stats_merge_fused<<<mg, 256, 0, st>>>(atten, stats, stats + sg.col_off_floats, sg.npr, sg.npc, gref, nmerge,
                                               g.R, g.C, ld, score1, ld1, score2, ld2, ws.rmax, ws.rsum, ws.cmax, ws.csum);
  // the GEMM wrote the instances in ascending order: start the labels pass on the last ones (still in L2)
  if (fine_tma_ok(atten, ld, g.R, g.C)) {
    int rc = launch_fine_labels_tma(atten, ld, b, g.R, g.C, f, ws, 1, st);
    if (rc) return rc;
  } else if (fine_vec_ok(atten, ld))
    k_fine_labels<true><<<grid, F2_THREADS, 0, st>>>(atten, g.R, g.C, ld, f.nstrip, f.nrt, ws.rmax, ws.rsum, ws.cmax, ws.csum,
                                                     ws.rowpm, ws.colpm, ws.ai0, ws.a0j, 1);
  else
    k_fine_labels<false><<<grid, F2_THREADS, 0, st>>>(atten, g.R, g.C, ld, f.nstrip, f.nrt, ws.rmax, ws.rsum, ws.cmax, ws.csum,
                                                      ws.rowpm, ws.colpm, ws.ai0, ws.a0j, 1);
  count_launch(2);
  return launch_labels_merge(ws.rowpm, ws.colpm, ws.ai0, ws.a0j, b, g.R, g.C, f.nrt, f.nstrip, w1, w2, st);
}

int run_fine_rows2(const float* atten, int ld, int b, const AssignGeom& g, const AssignWs& ws, const float* w1,
                   const float* w2, const float* pts2, float4* rowpart4, float* soft, float* asum, cudaStream_t st) {
  const FineGeom2 f = fine_geom2(g.R, g.C);
  const dim3 grid(f.nstrip, f.nrt, b);
  if (fine_tma_ok(atten, ld, g.R, g.C)) {
    CUtensorMap map;
    int rc = tc_make_map_plane(&map, atten + ld + 1, b, g.R - 1, g.C - 1, ld, (size_t)g.R * ld, F2_TC, FT_ROWS);
    if (rc) return rc;
    const size_t smem = sizeof(FtSmem<FT_ROWS_STAGES, 5, false>) + 128;
    UPK_CUDA_TRY(cudaFuncSetAttribute(k_fine_rows_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fine_rows_tma<<<fine_tma_grid(b * f.nrt * f.nstrip), FT_THREADS, smem, st>>>(
        map, b, g.R, g.C, f.nstrip, f.nrt, ws.rmax, ws.rsum, ws.cmax, ws.csum, w2, pts2, rowpart4, fine_l2_keep(b, g.R, g.C));
  } else if (fine_vec_ok(atten, ld))
    k_fine_rows<true><<<grid, F2_THREADS, 0, st>>>(atten, g.R, g.C, ld, f.nstrip, ws.rmax, ws.rsum, ws.cmax, ws.csum, w2, pts2,
                                                   rowpart4);
  else
    k_fine_rows<false><<<grid, F2_THREADS, 0, st>>>(atten, g.R, g.C, ld, f.nstrip, ws.rmax, ws.rsum, ws.cmax, ws.csum, w2, pts2,
                                                    rowpart4);
  count_launch();
  return launch_fine_rows_merge(rowpart4, w1, b, g.R - 1, f.nstrip, soft, asum, st);
}

}  // namespace upk
