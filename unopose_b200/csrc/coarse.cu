// Coarse pose solve: kernel families (2) correspondence sampling + triplet hypotheses,
// (3) batched Kabsch on 3-point sets, (4) fused hypothesis scoring + arg-max.
// Reference: compute_coarse_Rt[_overlap], core/unopose/utils/model_utils.py:336-490
// (~60 ATen/cuBLAS/cuSOLVER launches, a (B*K,196,196) distance tensor of 46 MB per
// instance).  Here: 4 kernels after the assignment passes, nothing larger than the
// (B,H) hypothesis pool is ever materialised.
#include <math.h>

#include "common.cuh"
#include "launch_count.h"
#include "peer.cuh"
#include "pose_internal.h"
#include "solver3.cuh"
#include "../../include/unopose_b200.h"

namespace upk {

// ---------------------------------------------------------------- (2)+(3)
// torch.searchsorted(cdf, u)  (right=False): first index with cdf[idx] >= u
__device__ __forceinline__ int lower_bound_f(const float* __restrict__ a, int n, float u) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (__ldg(a + mid) < u) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Kabsch on one 3-point hypothesis, float arithmetic laid out like the reference's
// weighted_procrustes (model_utils.py:704-730) with weights == ones: every weight is
// 1/(3 + eps) (eps = 1e-5 quirk), src = reference triplet p2,
// ref = query triplet p1 (call site :469).  The 3x3 solve runs in fp64 registers.
__device__ __forceinline__ void kabsch_triplet(const float (&p1)[3][3], const float (&p2)[3][3], float* R,
                                               float* t, float& resid) {
  const float wn = 1.0f / (3.0f + 1e-5f);
  float cs[3], cr[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    cs[a] = (p2[0][a] * wn + p2[1][a] * wn) + p2[2][a] * wn;
    cr[a] = (p1[0][a] * wn + p1[1][a] * wn) + p1[2][a] * wn;
  }
  double H[9];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float h = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) h = fmaf(p2[k][a] - cs[a], wn * (p1[k][c] - cr[c]), h);
      H[a * 3 + c] = (double)h;
    }
  double Rd[9];
  procrustes_rotation(H, Rd);
#pragma unroll
  for (int i = 0; i < 9; ++i) R[i] = (float)Rd[i];
  // t = c_ref - R c_src   (:729)
#pragma unroll
  for (int a = 0; a < 3; ++a)
    t[a] = cr[a] - fmaf(R[a * 3 + 2], cs[2], fmaf(R[a * 3 + 1], cs[1], R[a * 3 + 0] * cs[0]));
  // residual: mean_k || (p1_k - t) @ R - p2_k ||   (:475)
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float d0 = p1[k][0] - t[0], d1 = p1[k][1] - t[1], d2 = p1[k][2] - t[2];
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float x = fmaf(d2, R[6 + c], fmaf(d1, R[3 + c], d0 * R[c])) - p2[k][c];
      s = fmaf(x, x, s);
    }
    acc += sqrtf(s);
  }
  resid = acc / 3.0f;
}

constexpr int HY_THREADS = 128;

// one thread = one hypothesis: 3 searchsorted draws -> triplet gather -> Kabsch -> residual
__global__ void __launch_bounds__(HY_THREADS)
k_hypotheses(const float* __restrict__ cdf, const float* __restrict__ u, const float* __restrict__ pts1,
             const float* __restrict__ pts2, int n1, int n2, int H, int h0, int h1,
             int* __restrict__ idx1_out, int* __restrict__ idx2_out, float* __restrict__ Rs,
             float* __restrict__ ts, float* __restrict__ resid) {
  const int b = blockIdx.y;
  const int h = h0 + blockIdx.x * HY_THREADS + threadIdx.x;
  if (h >= h1) return;
  const float* cdf_b = cdf + (size_t)b * n1 * n2;
  const float* p1b = pts1 + (size_t)b * n1 * 3;
  const float* p2b = pts2 + (size_t)b * n2 * 3;
  float p1[3][3], p2[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float uu = u[((size_t)b * H + h) * 3 + k];
    int idx = lower_bound_f(cdf_b, n1 * n2, uu);
    int i1 = min(idx / n2, n1 - 1);  // :463-465 (clamp of the overflow index n1*n2)
    int i2 = min(idx % n2, n2 - 1);
    if (idx1_out) {
      idx1_out[((size_t)b * H + h) * 3 + k] = i1;
      idx2_out[((size_t)b * H + h) * 3 + k] = i2;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p1[k][a] = __ldg(p1b + i1 * 3 + a);
      p2[k][a] = __ldg(p2b + i2 * 3 + a);
    }
  }
  float R[9], t[3], r;
  kabsch_triplet(p1, p2, R, t, r);
  float* Ro = Rs + ((size_t)b * H + h) * 9;
#pragma unroll
  for (int i = 0; i < 9; ++i) Ro[i] = R[i];
  float* to = ts + ((size_t)b * H + h) * 3;
  to[0] = t[0]; to[1] = t[1]; to[2] = t[2];
  resid[(size_t)b * H + h] = r;
}

// Kabsch only (stage-wise entry: caller supplies the triplets)
__global__ void __launch_bounds__(HY_THREADS)
k_kabsch_triplets(const float* __restrict__ p1s, const float* __restrict__ p2s, int n,
                  float* __restrict__ Rs, float* __restrict__ ts, float* __restrict__ resid) {
  const int h = blockIdx.x * HY_THREADS + threadIdx.x;
  if (h >= n) return;
  float p1[3][3], p2[3][3];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      p1[k][a] = p1s[(size_t)h * 9 + k * 3 + a];
      p2[k][a] = p2s[(size_t)h * 9 + k * 3 + a];
    }
  float R[9], t[3], r;
  kabsch_triplet(p1, p2, R, t, r);
#pragma unroll
  for (int i = 0; i < 9; ++i) Rs[(size_t)h * 9 + i] = R[i];
  ts[(size_t)h * 3 + 0] = t[0]; ts[(size_t)h * 3 + 1] = t[1]; ts[(size_t)h * 3 + 2] = t[2];
  if (resid) resid[h] = r;
}

// ---------------------------------------------------------------- top-K smallest residuals
// torch.topk(dis, K, largest=False) (:476).  One CTA per instance: 4-pass 8-bit radix select on
// the float bit patterns (residuals are >= 0, NaN sorts last like torch), then an ORDERED
// compaction: the selected set is emitted in ascending pool index; ties at the K-th value are
// resolved towards the lower pool index (torch leaves that order unspecified).
constexpr int TK_THREADS = 1024;

__device__ __forceinline__ unsigned block_excl_scan_u32(unsigned v, unsigned* s_warp, unsigned& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    unsigned t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0u;
    unsigned winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      unsigned t = __shfl_up_sync(kFull, winc, o);
      if (lane >= o) winc += t;
    }
    s_warp[lane] = winc - w;  // exclusive per-warp offsets
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  unsigned res = s_warp[warp] + inc - v;
  total = s_warp[32];
  __syncthreads();
  return res;
}

// CACHED (H <= 8 * TK_THREADS: every pool size of the reference): a thread keeps its contiguous chunk of the row in
// registers for all four passes and the compaction (one read of the row instead of six); the digit search is a warp
// scan over 8 bins per lane instead of one thread walking 256 bins (round 1: 17 us per launch, now 10).
constexpr int TK_PER = 8;

template <bool CACHED>
__global__ void __launch_bounds__(TK_THREADS)
k_topk_smallest(const float* __restrict__ vals, int H, int K, int* __restrict__ top, int ld, int* __restrict__ zero_me) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_warp[33];
  __shared__ unsigned s_prefix, s_kth_rank;
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31;
  const unsigned* v = reinterpret_cast<const unsigned*>(vals) + (size_t)b * ld;   // ld >= H: row pitch (a slice of a pool)
  top += (size_t)b * K;
  if (zero_me && threadIdx.x == 0) zero_me[b] = 0;   // the ticket counter of the scoring launch that follows
  // each thread owns a contiguous chunk (the compaction below emits in ascending pool index)
  const int per = (H + TK_THREADS - 1) / TK_THREADS;
  const int beg = min(H, (int)threadIdx.x * per), end = min(H, beg + per);
  unsigned x[TK_PER];
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < TK_PER; ++k) x[k] = beg + k < end ? v[beg + k] : 0u;
  }
  unsigned prefix = 0, mask = 0;
  unsigned want = (unsigned)K;  // rank (1-based) of the K-th smallest within the current bucket
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    // plain shared atomics: aggregating per warp with __match_any_sync was measured slower at every pool size (the first
    // pass sees few distinct exponent bytes, but same-address shared atomics of a warp are combined by the hardware)
    if (CACHED) {
#pragma unroll
      for (int k = 0; k < TK_PER; ++k)
        if (beg + k < end && (x[k] & mask) == prefix) atomicAdd(&hist[(x[k] >> shift) & 255u], 1u);
    } else {
      for (int i = threadIdx.x; i < H; i += TK_THREADS) {
        const unsigned xv = v[i];
        if ((xv & mask) == prefix) atomicAdd(&hist[(xv >> shift) & 255u], 1u);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {   // digit of the K-th smallest: first bin whose inclusive count reaches `want`
      unsigned c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { c[j] = hist[8 * lane + j]; sum += c[j]; }
      unsigned inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
      }
      const unsigned reach = __ballot_sync(kFull, inc >= want);   // never empty: the bucket holds >= want elements
      if (lane == __ffs(reach) - 1) {
        unsigned acc = inc - sum;
        int j = 0;
#pragma unroll
        for (; j < 7; ++j) {
          if (acc + c[j] >= want) break;
          acc += c[j];
        }
        s_prefix = prefix | ((unsigned)(8 * lane + j) << shift);
        s_kth_rank = want - acc;
      }
    }
    __syncthreads();
    prefix = s_prefix;
    want = s_kth_rank;
    mask |= 255u << shift;
  }
  const unsigned kth = prefix;    // bit pattern of the K-th smallest value
  const unsigned n_equal = want;  // how many elements == kth belong to the selection
  // ordered compaction: counts of (< kth, == kth) packed in one word (H < 2^16 when CACHED), one block scan
  unsigned nl = 0, ne = 0;
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < TK_PER; ++k)
      if (beg + k < end) { nl += x[k] < kth; ne += x[k] == kth; }
  } else {
    for (int i = beg; i < end; ++i) {
      const unsigned xv = v[i];
      nl += xv < kth;
      ne += xv == kth;
    }
  }
  unsigned tot, l0, e0;
  if (CACHED) {
    const unsigned both = block_excl_scan_u32(nl | (ne << 16), s_warp, tot);
    l0 = both & 0xffffu; e0 = both >> 16;
  } else {
    l0 = block_excl_scan_u32(nl, s_warp, tot);
    e0 = block_excl_scan_u32(ne, s_warp, tot);
  }
  auto emit = [&](int i, unsigned xv) {
    if (xv < kth) {
      top[l0 + min(e0, n_equal)] = i;
      ++l0;
    } else if (xv == kth) {
      if (e0 < n_equal) top[l0 + e0] = i;
      ++e0;
    }
  };
  if (CACHED) {
#pragma unroll
    for (int k = 0; k < TK_PER; ++k)
      if (beg + k < end) emit(beg + k, x[k]);
  } else {
    for (int i = beg; i < end; ++i) emit(i, v[i]);
  }
}

static void launch_topk(const float* vals, int b, int n, int ld, int k, int* idx_out, cudaStream_t st,
                        int* zero_me = nullptr) {
  if (n <= TK_PER * TK_THREADS) k_topk_smallest<true><<<b, TK_THREADS, 0, st>>>(vals, n, k, idx_out, ld, zero_me);
  else k_topk_smallest<false><<<b, TK_THREADS, 0, st>>>(vals, n, k, idx_out, ld, zero_me);
  count_launch();
}

// arg-max over the K kept hypotheses (first maximum), gather R, t, score, pool index (:486-488); called by every
// thread of a CTA of at most 8 warps.  VOLATILE_SCORES: the scores were written by other CTAs of the same launch or by
// peers (read past L1).
template <bool VOLATILE_SCORES>
__device__ __forceinline__ void select_best(const float* __restrict__ scores, const int* __restrict__ top,
                                            const float* __restrict__ Rs, const float* __restrict__ ts, int H, int K, int b,
                                            float* __restrict__ R_out, float* __restrict__ t_out,
                                            float* __restrict__ score_out, int* __restrict__ pool_out,
                                            const int* __restrict__ pool_map, float* s_v, int* s_i) {
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  bool seen_nan = false;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float v = VOLATILE_SCORES ? __ldcg(scores + (size_t)b * K + k) : scores[(size_t)b * K + k];
    if (v != v) { if (!seen_nan) { seen_nan = true; bv = v; bi = k; } continue; }
    if (!seen_nan && (v > bv || (v == bv && k < bi))) { bv = v; bi = k; }
  }
  // warp reduce (NaN wins, then larger value, then smaller index) — torch.max propagates NaN
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(kFull, bv, o);
    int oi = __shfl_xor_sync(kFull, bi, o);
    bool a_nan = bv != bv, b_nan = ov != ov;
    bool take = b_nan ? (!a_nan || oi < bi) : (!a_nan && (ov > bv || (ov == bv && oi < bi)));
    if (take) { bv = ov; bi = oi; }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { s_v[warp] = bv; s_i[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nw; ++w) {
      float ov = s_v[w];
      int oi = s_i[w];
      bool a_nan = bv != bv, b_nan = ov != ov;
      bool take = b_nan ? (!a_nan || oi < bi) : (!a_nan && (ov > bv || (ov == bv && oi < bi)));
      if (take) { bv = ov; bi = oi; }
    }
    if (bi == 0x7fffffff) bi = 0;
    int h = top ? top[(size_t)b * K + bi] : bi;
    for (int i = 0; i < 9; ++i) R_out[(size_t)b * 9 + i] = Rs[((size_t)b * H + h) * 9 + i];
    for (int i = 0; i < 3; ++i) t_out[(size_t)b * 3 + i] = ts[((size_t)b * H + h) * 3 + i];
    score_out[b] = bv;
    if (pool_out) pool_out[b] = pool_map ? pool_map[(size_t)b * H + h] : h;   // compact candidate list -> pool index
  }
}

// ---------------------------------------------------------------- (4) scoring
// For hypothesis (R,t):  X = (pts1 - t) @ R ; d_i = sqrt(clamp(min_j (|X_i|^2 - 2 X_i.Y_j + |Y_j|^2), 0))
// score = sum_i w1_i / (sum_i d_i w1_i + 1e-8)        (model_utils.py:481-485, pairwise_distance :246-256)
// One CTA per (instance, kept hypothesis); model points staged in smem (SoA x|y|z||y|^2), four model
// points per step with packed f32x2 arithmetic (4.5 issue slots per point pair); the (B*K,N1,N2) distance tensor never exists.
constexpr int SC_THREADS = 224;  // 7 warps: one query point per thread at n1 = 196

// SC_HPC kept hypotheses per CTA: the staged model and every LDS.128 of the scan serve both (four per CTA, measured in
// round 2 with a generic Q-query scan: 80 registers with spills at 3 CTAs / SM, coarse solve 134 -> 138 us; a persistent
// grid pulling (instance, pair) items from a device-side queue instead of 4.05 waves of CTAs: 132 -> 143 us — the
// per-item barriers and the lost overlap of one CTA's prologue with its neighbours' scans cost more than the fifth wave)
constexpr int SC_HPC = 2;

// PEER: the scores go into EVERY rank's score table (peer stores) and the last CTA publishes the channel (peer.cuh)
template <bool PEER>
__global__ void __launch_bounds__(SC_THREADS, 4)
k_score(const float* __restrict__ pts1, const float* __restrict__ model, const float* __restrict__ w1,
        const float* __restrict__ Rs, const float* __restrict__ ts, const int* __restrict__ top,
        int n1, int nm, int H, int K, int k0, int k1, float* __restrict__ scores,
        const PeerCtx pc, size_t peer_off, size_t slab_bytes, int channel,
        int* __restrict__ tickets = nullptr, float* __restrict__ R_out = nullptr, float* __restrict__ t_out = nullptr,
        float* __restrict__ score_out = nullptr, int* __restrict__ pool_out = nullptr) {
  extern __shared__ __align__(16) float sm_model[];  // 4 x nm_pad (SoA x | y | z | |y|^2)
  __shared__ double s_red[2 * SC_HPC][SC_THREADS / 32];
  const int b = blockIdx.y;
  const int ka = k0 + blockIdx.x * SC_HPC;
  const int kb = min(ka + 1, k1 - 1);               // odd tail: the second slot repeats the first hypothesis
  const int nm_pad = (nm + 3) & ~3;
  float* mx = sm_model; float* my = mx + nm_pad; float* mz = my + nm_pad; float* mn = mz + nm_pad;
  stage_model_soa(model + (size_t)b * nm * 3, nm, nm_pad, mx, my, mz, mn);
  const int ha = top ? top[(size_t)b * K + ka] : ka, hb = top ? top[(size_t)b * K + kb] : kb;
  const float* Ra = Rs + ((size_t)b * H + ha) * 9;
  const float* ta = ts + ((size_t)b * H + ha) * 3;
  const float* Rb = Rs + ((size_t)b * H + hb) * 9;
  const float* tb = ts + ((size_t)b * H + hb) * 3;
  float ra[9], rb[9], tta[3], ttb[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) { ra[i] = Ra[i]; rb[i] = Rb[i]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) { tta[i] = ta[i]; ttb[i] = tb[i]; }
  __syncthreads();
  double num = 0.0, den_a = 0.0, den_b = 0.0;
  for (int i = threadIdx.x; i < n1; i += SC_THREADS) {
    const float* p = pts1 + ((size_t)b * n1 + i) * 3;
    const float p0 = p[0], p1 = p[1], p2 = p[2];
    float xa[4], xb[4];
    {
      const float d0 = p0 - tta[0], d1 = p1 - tta[1], d2 = p2 - tta[2];
      xa[0] = fmaf(d2, ra[6], fmaf(d1, ra[3], d0 * ra[0]));
      xa[1] = fmaf(d2, ra[7], fmaf(d1, ra[4], d0 * ra[1]));
      xa[2] = fmaf(d2, ra[8], fmaf(d1, ra[5], d0 * ra[2]));
      xa[3] = sumsq3_torch(xa[0], xa[1], xa[2]);
    }
    {
      const float d0 = p0 - ttb[0], d1 = p1 - ttb[1], d2 = p2 - ttb[2];
      xb[0] = fmaf(d2, rb[6], fmaf(d1, rb[3], d0 * rb[0]));
      xb[1] = fmaf(d2, rb[7], fmaf(d1, rb[4], d0 * rb[1]));
      xb[2] = fmaf(d2, rb[8], fmaf(d1, rb[5], d0 * rb[2]));
      xb[3] = sumsq3_torch(xb[0], xb[1], xb[2]);
    }
    float best_a, best_b;
    nn_min_expansion2(mx, my, mz, mn, nm_pad, xa, xb, best_a, best_b);
    const float w = w1[(size_t)b * n1 + i];
    num += (double)w;
    den_a += (double)(sqrtf(fmaxf(best_a, 0.f)) * w);
    den_b += (double)(sqrtf(fmaxf(best_b, 0.f)) * w);
  }
  num = warp_sum(num);
  den_a = warp_sum(den_a);
  den_b = warp_sum(den_b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_red[0][warp] = num; s_red[1][warp] = den_a; s_red[2][warp] = den_b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double n = 0.0, da = 0.0, db = 0.0;
    for (int w = 0; w < SC_THREADS / 32; ++w) { n += s_red[0][w]; da += s_red[1][w]; db += s_red[2][w]; }
    const float sa = (float)n / ((float)da + 1e-8f), sb = (float)n / ((float)db + 1e-8f);
    if (!PEER) {
      scores[(size_t)b * K + ka] = sa;
      if (ka + 1 < k1) scores[(size_t)b * K + ka + 1] = sb;
    } else {
      const size_t o = peer_off + ((pc.epoch[channel] + 1) & 1) * slab_bytes;
      for (int r = 0; r < pc.world; ++r) {
        float* d = reinterpret_cast<float*>(pc.data[r] + o);
        d[(size_t)b * K + ka] = sa;
        if (ka + 1 < k1) d[(size_t)b * K + ka + 1] = sb;
      }
    }
  }
  if (PEER) peer_publish(pc, channel, pc.epoch[channel] + 1, gridDim.x * gridDim.y);
  if (!PEER && tickets) {
    // the last CTA of an instance (ticket counter, zeroed by the top-K launch) selects the best hypothesis: the
    // separate k_select launch (5 us incl. its boundary) is gone
    __shared__ int s_last;
    __shared__ float s_v[8];
    __shared__ int s_i[8];
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(tickets + b, 1) == (int)gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      select_best<true>(scores, top, Rs, ts, H, K, b, R_out, t_out, score_out, pool_out, nullptr, s_v, s_i);
    }
  }
}

// PEER: `scores` is the local copy of the exchanged score table; wait until every rank has stored its slice
template <bool PEER>
__global__ void __launch_bounds__(256)
k_select(const float* __restrict__ scores, const int* __restrict__ top, const float* __restrict__ Rs,
         const float* __restrict__ ts, int H, int K, float* __restrict__ R_out, float* __restrict__ t_out,
         float* __restrict__ score_out, int* __restrict__ pool_out,
         const PeerCtx pc, size_t peer_off, size_t slab_bytes, int channel, const int* __restrict__ pool_map = nullptr) {
  __shared__ float s_v[8];
  __shared__ int s_i[8];
  const int b = blockIdx.x;
  if (PEER) {
    const unsigned long long e = pc.epoch[channel];
    peer_wait(pc, channel, e);
    scores = reinterpret_cast<const float*>(pc.data[pc.rank] + peer_off + (e & 1) * slab_bytes);
  }
  select_best<PEER>(scores, top, Rs, ts, H, K, b, R_out, t_out, score_out, pool_out, pool_map, s_v, s_i);
}

// ---------------------------------------------------------------- hypothesis sharding
// Candidate exchange between the ranks that share one instance batch: every rank packs the kl smallest hypotheses of
// ITS slice of the pool into a fixed-size record list (kc records per instance, padded with +inf / -1), the lists are
// all-gathered, and every rank scatters the other ranks' records into its dense pool arrays, so that the SAME top-K
// kernel a single GPU runs re-selects the global top-K (identical tie rule).  Record: {residual, pool index (int
// bits), R (9), t (3)} = 14 floats.
constexpr int CAND_F = 14;

// PEER: the record goes into slot [rank] of EVERY rank's gathered[world][b][kc][14] array, then the channel is published
template <bool PEER>
__global__ void __launch_bounds__(128)
k_pack_candidates(const float* __restrict__ resid, const float* __restrict__ Rs, const float* __restrict__ ts,
                  const int* __restrict__ top_local, int h0, int H, int kl, int kc, float* __restrict__ cand,
                  const PeerCtx pc, size_t peer_off, size_t slab_bytes, int channel) {
  const int b = blockIdx.y;
  const int j = blockIdx.x * 128 + threadIdx.x;
  if (j < kc) {
    float rec[CAND_F];
    if (j < kl) {
      const int h = h0 + top_local[(size_t)b * kl + j];
      rec[0] = resid[(size_t)b * H + h];
      rec[1] = __int_as_float(h);
#pragma unroll
      for (int i = 0; i < 9; ++i) rec[2 + i] = Rs[((size_t)b * H + h) * 9 + i];
#pragma unroll
      for (int i = 0; i < 3; ++i) rec[11 + i] = ts[((size_t)b * H + h) * 3 + i];
    } else {
      rec[0] = INFINITY;
      rec[1] = __int_as_float(-1);
#pragma unroll
      for (int i = 2; i < CAND_F; ++i) rec[i] = 0.f;
    }
    if (!PEER) {
      float* o = cand + ((size_t)b * kc + j) * CAND_F;
#pragma unroll
      for (int i = 0; i < CAND_F; ++i) o[i] = rec[i];
    } else {
      const size_t o = peer_off + ((pc.epoch[channel] + 1) & 1) * slab_bytes +
                       ((((size_t)pc.rank * gridDim.y + b) * kc + j) * CAND_F) * sizeof(float);
      for (int r = 0; r < pc.world; ++r) {
        float2* d = reinterpret_cast<float2*>(pc.data[r] + o);     // records are 56 bytes: 8-byte aligned
#pragma unroll
        for (int i = 0; i < CAND_F / 2; ++i) d[i] = make_float2(rec[2 * i], rec[2 * i + 1]);
      }
    }
  }
  if (PEER) peer_publish(pc, channel, pc.epoch[channel] + 1, gridDim.x * gridDim.y);
}

template <bool PEER>
__global__ void __launch_bounds__(128)
k_unpack_candidates(const float* __restrict__ allc /* [world][b][kc][14] */, int world, int skip_rank, int nb, int H,
                    int kc, float* __restrict__ resid, float* __restrict__ Rs, float* __restrict__ ts,
                    const PeerCtx pc, size_t peer_off, size_t slab_bytes, int channel) {
  if (PEER) {   // the local copy of the exchanged array, once every rank has stored its list
    const unsigned long long e = pc.epoch[channel];
    peer_wait(pc, channel, e);
    allc = reinterpret_cast<const float*>(pc.data[pc.rank] + peer_off + (e & 1) * slab_bytes);
  }
  const int b = blockIdx.y;
  const int q = blockIdx.x * 128 + threadIdx.x;   // (rank, j)
  if (q >= world * kc) return;
  const int r = q / kc, j = q - r * kc;
  if (r == skip_rank) return;                     // my own slice is already in place
  const float* c = allc + (((size_t)r * nb + b) * kc + j) * CAND_F;
  float rec[CAND_F];
#pragma unroll
  for (int i = 0; i < CAND_F; ++i) rec[i] = PEER ? __ldcg(c + i) : c[i];
  const int h = __float_as_int(rec[1]);
  if (h < 0 || h >= H) return;
  resid[(size_t)b * H + h] = rec[0];
#pragma unroll
  for (int i = 0; i < 9; ++i) Rs[((size_t)b * H + h) * 9 + i] = rec[2 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) ts[((size_t)b * H + h) * 3 + i] = rec[11 + i];
}

// Compact form of the merge (round 2): the gathered lists ARE the only candidates for the global top-K, so instead of
// scattering them into the dense pool arrays (and running the top-K over H mostly-infinite entries) they are laid out
// as a compact pool of world * kc entries per instance.  Every rank's list is ascending in pool index and the ranks
// own ascending slices, so compact order == pool-index order and the SAME top-K kernel applies the SAME tie rule
// (lower index first); padding records sort after everything (key 0xFFFFFFFF).
template <bool PEER>
__global__ void __launch_bounds__(128)
k_unpack_candidates_compact(const float* __restrict__ allc /* [world][b][kc][14] */, int world, int nb, int kc,
                            float* __restrict__ resid_c, float* __restrict__ Rs_c, float* __restrict__ ts_c,
                            int* __restrict__ pool_c, const PeerCtx pc, size_t peer_off, size_t slab_bytes, int channel) {
  if (PEER) {
    const unsigned long long e = pc.epoch[channel];
    peer_wait(pc, channel, e);
    allc = reinterpret_cast<const float*>(pc.data[pc.rank] + peer_off + (e & 1) * slab_bytes);
  }
  const int b = blockIdx.y;
  const int q = blockIdx.x * 128 + threadIdx.x;   // (rank, j) == compact index
  const int n = world * kc;
  if (q >= n) return;
  const int r = q / kc, j = q - r * kc;
  const float* c = allc + (((size_t)r * nb + b) * kc + j) * CAND_F;
  float rec[CAND_F];
#pragma unroll
  for (int i = 0; i < CAND_F; ++i) rec[i] = PEER ? __ldcg(c + i) : c[i];
  const int h = __float_as_int(rec[1]);
  const size_t o = (size_t)b * n + q;
  resid_c[o] = h < 0 ? __uint_as_float(0xFFFFFFFFu) : rec[0];
  pool_c[o] = h;
#pragma unroll
  for (int i = 0; i < 9; ++i) Rs_c[o * 9 + i] = rec[2 + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) ts_c[o * 3 + i] = rec[11 + i];
}

__global__ void __launch_bounds__(256)
k_fill_f32(float* __restrict__ p, size_t n, float v) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) p[i] = v;
}

struct CoarseWs {
  AssignWs a;
  float* w1; float* w2;
  float* pmat; double* prow; float* cdf;
  float* Rs; float* ts; float* resid;
  int* top; float* scores;
  int* tickets;
};

static void carve_coarse(Carver& cv, int b, int n1, int n2, int H, int K, const AssignGeom& g, CoarseWs& w) {
  carve_assign(cv, b, g, w.a);
  w.w1 = cv.take<float>((size_t)b * n1);
  w.w2 = cv.take<float>((size_t)b * n2);
  w.pmat = cv.take<float>((size_t)b * n1 * n2);
  w.prow = cv.take<double>((size_t)b * n1 * g.ntc);
  w.cdf = cv.take<float>((size_t)b * n1 * n2);
  w.Rs = cv.take<float>((size_t)b * H * 9);
  w.ts = cv.take<float>((size_t)b * H * 3);
  w.resid = cv.take<float>((size_t)b * H);
  w.top = cv.take<int>((size_t)b * K);
  w.scores = cv.take<float>((size_t)b * K);
  w.tickets = cv.take<int>((size_t)b);
}

static int launch_score(const float* pts1, const float* model, const float* w1, const float* Rs,
                        const float* ts, const int* top, int b, int n1, int nm, int H, int K, int k0, int k1,
                        float* scores, cudaStream_t st, const upk_peer_t* peer = nullptr, size_t peer_off = 0,
                        size_t slab_bytes = 0, int channel = 0, int* tickets = nullptr, float* R_out = nullptr,
                        float* t_out = nullptr, float* score_out = nullptr, int* pool_out = nullptr) {
  if (k1 <= k0 && !peer) return UPK_OK;
  size_t smem = (size_t)((nm + 3) & ~3) * 4 * sizeof(float);
  if (smem > 200 * 1024) return UPK_ERR_UNSUPPORTED;
  dim3 grid(ceil_div(k1 - k0, SC_HPC), b);
  if (peer) {
    if (k1 <= k0) return UPK_ERR_INVALID_ARG;   // every rank must publish: the host gives each rank a non-empty slice
    if (smem > 40 * 1024)
      UPK_CUDA_TRY(cudaFuncSetAttribute(k_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_score<true><<<grid, SC_THREADS, smem, st>>>(pts1, model, w1, Rs, ts, top, n1, nm, H, K, k0, k1, nullptr,
                                                  make_peer_ctx(peer), peer_off, slab_bytes, channel);
  } else {
    if (smem > 40 * 1024)
      UPK_CUDA_TRY(cudaFuncSetAttribute(k_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_score<false><<<grid, SC_THREADS, smem, st>>>(pts1, model, w1, Rs, ts, top, n1, nm, H, K, k0, k1, scores, PeerCtx(), 0, 0, 0,
                                                   tickets, R_out, t_out, score_out, pool_out);
  }
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

}  // namespace upk

using namespace upk;

extern "C" {

size_t upk_coarse_pose_workspace_bytes(int b, int n1, int n2, int n_hyp, int n_keep) {
  if (b <= 0 || n1 <= 0 || n2 <= 0 || n_hyp <= 0 || n_keep <= 0) return 0;
  Carver cv(nullptr);
  CoarseWs w;
  carve_coarse(cv, b, n1, n2, n_hyp, n_keep, assign_geom(n1 + 1, n2 + 1), w);
  return cv.bytes();
}

int upk_coarse_pose(const float* atten, const float* score1, int score1_ld, const float* score2,
                    int score2_ld, const float* pts1, const float* pts2, const float* model_pts,
                    int n_model, const float* u, int b, int n1, int n2, int n_hyp, int n_keep,
                    void* workspace, size_t workspace_bytes, float* R_out, float* t_out, float* score_out,
                    int* pool_idx_out, const upk_coarse_debug* dbg, upk_stream_t stream) {
  if (b < 0 || n1 <= 0 || n2 <= 0 || n_hyp <= 0 || n_keep <= 0 || n_keep > n_hyp) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  if (!atten || !pts1 || !pts2 || !u || !workspace || !R_out || !t_out || !score_out) return UPK_ERR_INVALID_ARG;
  if ((score1 == nullptr) != (score2 == nullptr)) return UPK_ERR_INVALID_ARG;
  if ((long long)n1 * n2 >= (1LL << 31)) return UPK_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  AssignGeom g = assign_geom(n1 + 1, n2 + 1);
  Carver cv(workspace);
  CoarseWs w;
  carve_coarse(cv, b, n1, n2, n_hyp, n_keep, g, w);
  if (cv.bytes() > workspace_bytes) return UPK_ERR_INVALID_ARG;
  if (!model_pts) { model_pts = pts2; n_model = n2; }
  int rc;
  // masks + sampling CDF: one cluster kernel (8 CTAs per instance, DSMEM), bit-exact with torch's CUDA kernels.
  // (a single-CTA-per-instance fused variant was measured at 70 us against ~50 us for the six launches of the tile
  //  pipeline at B = 16: too little parallelism)
  rc = run_coarse_assign_exact(atten, score1, score1_ld, score2, score2_ld, b, n1 + 1, n2 + 1, w.w1, w.w2, w.cdf, st);
  if (rc == UPK_ERR_UNSUPPORTED) {   // geometry outside the cluster kernel: the tile pipeline (close to, not bit-identical with torch)
    if ((rc = run_assignment_labels(atten, score1, score1_ld, score2, score2_ld, b, g, w.a, w.w1, w.w2, st))) return rc;
    if ((rc = run_coarse_P(atten, score1, score1_ld, score2, score2_ld, b, g, w.a, w.w1, w.w2, w.pmat, w.prow, st))) return rc;
    rc = run_cdf(w.pmat, w.prow, b, n1, n2, g.ntc, w.cdf, st);
  }
  if (rc) return rc;
  {
    dim3 grid(ceil_div(n_hyp, HY_THREADS), b);
    k_hypotheses<<<grid, HY_THREADS, 0, st>>>(w.cdf, u, pts1, pts2, n1, n2, n_hyp, 0, n_hyp,
                                              dbg ? dbg->idx1 : nullptr, dbg ? dbg->idx2 : nullptr, w.Rs,
                                              w.ts, w.resid);
    count_launch();
  }
  // the top-K launch zeroes the per-instance tickets; the LAST scoring CTA of an instance selects the best hypothesis
  launch_topk(w.resid, b, n_hyp, n_hyp, n_keep, w.top, st, w.tickets);
  if ((rc = launch_score(pts1, model_pts, w.w1, w.Rs, w.ts, w.top, b, n1, n_model, n_hyp, n_keep, 0, n_keep,
                         w.scores, st, nullptr, 0, 0, 0, w.tickets, R_out, t_out, score_out, pool_idx_out)))
    return rc;
  if (dbg) {
    // optional copies of the intermediates for stage-wise parity tests
    if (dbg->w1) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->w1, w.w1, sizeof(float) * (size_t)b * n1, cudaMemcpyDeviceToDevice, st));
    if (dbg->w2) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->w2, w.w2, sizeof(float) * (size_t)b * n2, cudaMemcpyDeviceToDevice, st));
    if (dbg->cdf) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->cdf, w.cdf, sizeof(float) * (size_t)b * n1 * n2, cudaMemcpyDeviceToDevice, st));
    if (dbg->Rs) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->Rs, w.Rs, sizeof(float) * (size_t)b * n_hyp * 9, cudaMemcpyDeviceToDevice, st));
    if (dbg->ts) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->ts, w.ts, sizeof(float) * (size_t)b * n_hyp * 3, cudaMemcpyDeviceToDevice, st));
    if (dbg->resid) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->resid, w.resid, sizeof(float) * (size_t)b * n_hyp, cudaMemcpyDeviceToDevice, st));
    if (dbg->top) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->top, w.top, sizeof(int) * (size_t)b * n_keep, cudaMemcpyDeviceToDevice, st));
    if (dbg->scores) UPK_CUDA_TRY(cudaMemcpyAsync(dbg->scores, w.scores, sizeof(float) * (size_t)b * n_keep, cudaMemcpyDeviceToDevice, st));
  }
  UPK_RETURN_LAST_ERROR();
}

// ---- stage-wise entry points (identical-input parity tests, hypothesis sharding) ----

size_t upk_coarse_assignment_workspace_bytes(int b, int n1, int n2) {
  if (b <= 0 || n1 <= 0 || n2 <= 0) return 0;
  Carver cv(nullptr);
  AssignGeom g = assign_geom(n1 + 1, n2 + 1);
  AssignWs a;
  carve_assign(cv, b, g, a);
  cv.take<float>((size_t)b * n1 * n2);
  cv.take<double>((size_t)b * n1 * g.ntc);
  return cv.bytes();
}

int upk_coarse_assignment(const float* atten, const float* score1, int score1_ld, const float* score2,
                          int score2_ld, int b, int n1, int n2, void* workspace, size_t workspace_bytes,
                          float* w1_out, float* w2_out, float* cdf_out, upk_stream_t stream) {
  if (b < 0 || n1 <= 0 || n2 <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  if (!atten || !workspace || !w1_out || !w2_out || !cdf_out) return UPK_ERR_INVALID_ARG;
  if ((score1 == nullptr) != (score2 == nullptr)) return UPK_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  AssignGeom g = assign_geom(n1 + 1, n2 + 1);
  Carver cv(workspace);
  AssignWs a;
  carve_assign(cv, b, g, a);
  float* pmat = cv.take<float>((size_t)b * n1 * n2);
  double* prow = cv.take<double>((size_t)b * n1 * g.ntc);
  if (cv.bytes() > workspace_bytes) return UPK_ERR_INVALID_ARG;
  int rc = run_coarse_assign_exact(atten, score1, score1_ld, score2, score2_ld, b, n1 + 1, n2 + 1, w1_out, w2_out, cdf_out, st);
  if (rc != UPK_ERR_UNSUPPORTED) return rc;
  if ((rc = run_assignment_labels(atten, score1, score1_ld, score2, score2_ld, b, g, a, w1_out, w2_out, st))) return rc;
  if ((rc = run_coarse_P(atten, score1, score1_ld, score2, score2_ld, b, g, a, w1_out, w2_out, pmat, prow, st))) return rc;
  return run_cdf(pmat, prow, b, n1, n2, g.ntc, cdf_out, st);
}

int upk_sample_hypotheses(const float* cdf, const float* u, const float* pts1, const float* pts2, int b,
                          int n1, int n2, int n_hyp, int h_begin, int h_end, int* idx1_out, int* idx2_out,
                          float* Rs, float* ts, float* resid, upk_stream_t stream) {
  if (b < 0 || n1 <= 0 || n2 <= 0 || n_hyp <= 0 || h_begin < 0 || h_end > n_hyp || h_begin > h_end)
    return UPK_ERR_INVALID_ARG;
  if (b == 0 || h_begin == h_end) return UPK_OK;
  if ((idx1_out == nullptr) != (idx2_out == nullptr)) return UPK_ERR_INVALID_ARG;
  dim3 grid(ceil_div(h_end - h_begin, HY_THREADS), b);
  k_hypotheses<<<grid, HY_THREADS, 0, (cudaStream_t)stream>>>(cdf, u, pts1, pts2, n1, n2, n_hyp, h_begin, h_end,
                                                              idx1_out, idx2_out, Rs, ts, resid);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_kabsch_triplets(const float* p1, const float* p2, int n, float* Rs, float* ts, float* resid,
                        upk_stream_t stream) {
  if (n < 0) return UPK_ERR_INVALID_ARG;
  if (n == 0) return UPK_OK;
  k_kabsch_triplets<<<ceil_div(n, HY_THREADS), HY_THREADS, 0, (cudaStream_t)stream>>>(p1, p2, n, Rs, ts, resid);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_topk_smallest(const float* vals, int b, int n, int k, int* idx_out, upk_stream_t stream) {
  if (b < 0 || n <= 0 || k <= 0 || k > n) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  launch_topk(vals, b, n, n, k, idx_out, (cudaStream_t)stream);
  UPK_RETURN_LAST_ERROR();
}

int upk_topk_smallest_ld(const float* vals, int b, int n, int ld, int k, int* idx_out, upk_stream_t stream) {
  if (b < 0 || n <= 0 || k <= 0 || k > n || ld < n) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  launch_topk(vals, b, n, ld, k, idx_out, (cudaStream_t)stream);
  UPK_RETURN_LAST_ERROR();
}

int upk_score_hypotheses(const float* pts1, const float* model_pts, const float* w1, const float* Rs,
                         const float* ts, const int* top, int b, int n1, int n_model, int n_hyp, int n_keep,
                         int k_begin, int k_end, float* scores, upk_stream_t stream) {
  if (b < 0 || n1 <= 0 || n_model <= 0 || n_hyp <= 0 || n_keep <= 0 || k_begin < 0 || k_end > n_keep ||
      k_begin > k_end)
    return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  return launch_score(pts1, model_pts, w1, Rs, ts, top, b, n1, n_model, n_hyp, n_keep, k_begin, k_end, scores,
                      (cudaStream_t)stream);
}

int upk_fill_f32(float* p, size_t n, float value, upk_stream_t stream) {
  if (n == 0) return UPK_OK;
  if (!p) return UPK_ERR_INVALID_ARG;
  const size_t blocks = (n + 255) / 256;
  k_fill_f32<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, (cudaStream_t)stream>>>(p, n, value);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_pack_candidates(const float* resid, const float* Rs, const float* ts, const int* top_local, int b, int n_hyp,
                        int h_begin, int n_local, int n_slots, float* cand_out, upk_stream_t stream) {
  if (b < 0 || n_hyp <= 0 || h_begin < 0 || n_local < 0 || n_slots < n_local || n_slots <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  if (!resid || !Rs || !ts || !cand_out || (n_local > 0 && !top_local)) return UPK_ERR_INVALID_ARG;
  k_pack_candidates<false><<<dim3(ceil_div(n_slots, 128), b), 128, 0, (cudaStream_t)stream>>>(
      resid, Rs, ts, top_local, h_begin, n_hyp, n_local, n_slots, cand_out, PeerCtx(), 0, 0, 0);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_unpack_candidates(const float* gathered, int world, int my_rank, int b, int n_hyp, int n_slots, float* resid,
                          float* Rs, float* ts, upk_stream_t stream) {
  if (world <= 0 || b < 0 || n_hyp <= 0 || n_slots <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  if (!gathered || !resid || !Rs || !ts) return UPK_ERR_INVALID_ARG;
  k_unpack_candidates<false><<<dim3(ceil_div(world * n_slots, 128), b), 128, 0, (cudaStream_t)stream>>>(
      gathered, world, my_rank, b, n_hyp, n_slots, resid, Rs, ts, PeerCtx(), 0, 0, 0);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_select_best(const float* scores, const int* top, const float* Rs, const float* ts, int b, int n_hyp,
                    int n_keep, float* R_out, float* t_out, float* score_out, int* pool_idx_out,
                    upk_stream_t stream) {
  if (b < 0 || n_hyp <= 0 || n_keep <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  k_select<false><<<b, 256, 0, (cudaStream_t)stream>>>(scores, top, Rs, ts, n_hyp, n_keep, R_out, t_out, score_out,
                                                       pool_idx_out, PeerCtx(), 0, 0, 0);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_unpack_candidates_compact(const float* gathered, int world, int b, int n_slots, float* resid_c, float* Rs_c,
                                  float* ts_c, int* pool_c, upk_stream_t stream) {
  if (world <= 0 || b < 0 || n_slots <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  if (!gathered || !resid_c || !Rs_c || !ts_c || !pool_c) return UPK_ERR_INVALID_ARG;
  k_unpack_candidates_compact<false><<<dim3(ceil_div(world * n_slots, 128), b), 128, 0, (cudaStream_t)stream>>>(
      gathered, world, b, n_slots, resid_c, Rs_c, ts_c, pool_c, PeerCtx(), 0, 0, 0);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_select_best_map(const float* scores, const int* top, const float* Rs, const float* ts, const int* pool_map, int b,
                        int n_hyp, int n_keep, float* R_out, float* t_out, float* score_out, int* pool_idx_out,
                        upk_stream_t stream) {
  if (b < 0 || n_hyp <= 0 || n_keep <= 0) return UPK_ERR_INVALID_ARG;
  if (b == 0) return UPK_OK;
  k_select<false><<<b, 256, 0, (cudaStream_t)stream>>>(scores, top, Rs, ts, n_hyp, n_keep, R_out, t_out, score_out,
                                                       pool_idx_out, PeerCtx(), 0, 0, 0, pool_map);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

// ---- peer-exchange forms (fused compute + all-gather over NVLink peer memory; csrc/peer.cuh) ----

int upk_unpack_candidates_compact_peer(const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel, int b,
                                       int n_slots, float* resid_c, float* Rs_c, float* ts_c, int* pool_c,
                                       upk_stream_t stream) {
  if (b <= 0 || n_slots <= 0 || !resid_c || !Rs_c || !ts_c || !pool_c || !peer_ctx_ok(peer, channel)) return UPK_ERR_INVALID_ARG;
  k_unpack_candidates_compact<true><<<dim3(ceil_div(peer->world * n_slots, 128), b), 128, 0, (cudaStream_t)stream>>>(
      nullptr, peer->world, b, n_slots, resid_c, Rs_c, ts_c, pool_c, make_peer_ctx(peer), data_offset, slab_bytes, channel);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_pack_candidates_peer(const float* resid, const float* Rs, const float* ts, const int* top_local, int b, int n_hyp,
                             int h_begin, int n_local, int n_slots, const upk_peer_t* peer, size_t data_offset,
                             size_t slab_bytes, int channel, upk_stream_t stream) {
  if (b <= 0 || n_hyp <= 0 || h_begin < 0 || n_local < 0 || n_slots < n_local || n_slots <= 0) return UPK_ERR_INVALID_ARG;
  if (!resid || !Rs || !ts || (n_local > 0 && !top_local) || !peer_ctx_ok(peer, channel)) return UPK_ERR_INVALID_ARG;
  if ((data_offset | slab_bytes) & 15) return UPK_ERR_INVALID_ARG;
  if ((size_t)peer->world * b * n_slots * CAND_F * sizeof(float) > slab_bytes) return UPK_ERR_INVALID_ARG;
  k_pack_candidates<true><<<dim3(ceil_div(n_slots, 128), b), 128, 0, (cudaStream_t)stream>>>(
      resid, Rs, ts, top_local, h_begin, n_hyp, n_local, n_slots, nullptr, make_peer_ctx(peer), data_offset, slab_bytes,
      channel);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_unpack_candidates_peer(const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel, int b, int n_hyp,
                               int n_slots, float* resid, float* Rs, float* ts, upk_stream_t stream) {
  if (b <= 0 || n_hyp <= 0 || n_slots <= 0 || !resid || !Rs || !ts || !peer_ctx_ok(peer, channel)) return UPK_ERR_INVALID_ARG;
  k_unpack_candidates<true><<<dim3(ceil_div(peer->world * n_slots, 128), b), 128, 0, (cudaStream_t)stream>>>(
      nullptr, peer->world, peer->rank, b, n_hyp, n_slots, resid, Rs, ts, make_peer_ctx(peer), data_offset, slab_bytes,
      channel);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

int upk_score_hypotheses_peer(const float* pts1, const float* model_pts, const float* w1, const float* Rs, const float* ts,
                              const int* top, int b, int n1, int n_model, int n_hyp, int n_keep, int k_begin, int k_end,
                              const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel,
                              upk_stream_t stream) {
  if (b <= 0 || n1 <= 0 || n_model <= 0 || n_hyp <= 0 || n_keep <= 0 || k_begin < 0 || k_end > n_keep ||
      k_begin >= k_end || !peer_ctx_ok(peer, channel))
    return UPK_ERR_INVALID_ARG;
  if (((data_offset | slab_bytes) & 15) || (size_t)b * n_keep * sizeof(float) > slab_bytes) return UPK_ERR_INVALID_ARG;
  return launch_score(pts1, model_pts, w1, Rs, ts, top, b, n1, n_model, n_hyp, n_keep, k_begin, k_end, nullptr,
                      (cudaStream_t)stream, peer, data_offset, slab_bytes, channel);
}

int upk_select_best_peer(const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel, const int* top,
                         const float* Rs, const float* ts, const int* pool_map, int b, int n_hyp, int n_keep, float* R_out,
                         float* t_out, float* score_out, int* pool_idx_out, upk_stream_t stream) {
  if (b <= 0 || n_hyp <= 0 || n_keep <= 0 || !peer_ctx_ok(peer, channel)) return UPK_ERR_INVALID_ARG;
  k_select<true><<<b, 256, 0, (cudaStream_t)stream>>>(nullptr, top, Rs, ts, n_hyp, n_keep, R_out, t_out, score_out,
                                                      pool_idx_out, make_peer_ctx(peer), data_offset, slab_bytes, channel,
                                                      pool_map);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

}  // extern "C"
