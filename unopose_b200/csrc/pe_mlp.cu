// (f1) The MLP half of PositionalEncoding, fused: SharedMLP([cin, C1, C2, C3]) -> max over the ball
//      (core/unopose/model/oneref_predator_fine_point_matching.py:138-178: `self.mlp1(...).max(dim=3)[0]`;
//       SharedMLP = three 1x1 Conv2d + BatchNorm2d + ReLU, pointnet2/pytorch_utils.py:25-48).
//
// The reference runs, per layer, a cuDNN/cuBLAS GEMM, a batch-norm kernel and a ReLU kernel over activations of up to
// (B,128,2048,256) fp32 = 4.3 GB at B = 16, then a max reduction: ~10 full passes over GB-sized tensors.  Here a CTA
// keeps a 128-sample tile ON CHIP through all three layers: the host folds the eval-mode batch norm into the conv
// (W' = W g/sqrt(v+eps), b' = beta - mean g/sqrt(v+eps)); the three GEMMs run on tcgen05 (3xTF32, accumulators in
// TMEM); each layer's epilogue (bias, ReLU, hi/lo split) writes the NEXT layer's A operand straight into shared memory
// in the UMMA SWIZZLE_64B layout; the last epilogue reduces the max over the ball.  HBM sees the (B,cin,m,ns) input
// once and the (B,C3,m) output.
//
// CTA anatomy (320 threads, persistent): warp 1 issues every MMA; two groups of 4 warps (2..5, 6..9) each own a tile
// slot (64 KB of operand smem, 224 TMEM columns) and walk their tile through
//   load+split A1 -> [L1] -> epilogue 1 -> [L2] -> epilogue 2 -> [L3] -> epilogue 3 (max)
// handing over with two mbarriers per group (a_ready: 4 warp arrivals, d_ready: tcgen05.commit); the groups run half a
// tile apart so one group's epilogues overlap the other's MMAs.  The folded weights (84 KB, hi + lo) stay resident.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "launch_count.h"
#include "pose_internal.h"
#include "tc_ptx.cuh"
#include "../../include/unopose_b200.h"

namespace upk {

constexpr int PM_C1 = 32, PM_C2 = 64, PM_C3 = 128;      // maxima (= the UNOPose SharedMLP [cin, 32, 64, 128])
constexpr int PM_THREADS = 320;
constexpr int PM_TILE = TC_BM * TC_BK;                   // floats of one 128-row x 16-wide operand chunk (8 KB)

struct __align__(1024) PmSmem {
  float w1_hi[PM_C1 * TC_BK], w1_lo[PM_C1 * TC_BK];                      // [row][16], cols >= cin zero
  float w2_hi[PM_C1 / 16][PM_C2 * TC_BK], w2_lo[PM_C1 / 16][PM_C2 * TC_BK];   // chunk-major K
  float w3_hi[PM_C2 / 16][PM_C3 * TC_BK], w3_lo[PM_C2 / 16][PM_C3 * TC_BK];
  float a_hi[2][4][PM_TILE], a_lo[2][4][PM_TILE];                        // per group: A1 = chunk 0, A2 = 0..1, A3 = 0..3
  float red[2][4][PM_C3];
  float b1[PM_C1], b2[PM_C2], b3[PM_C3];
  unsigned long long a_ready[2], d_ready[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float pm_tf32(float v) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v));
  return __uint_as_float(u);
}

// element (row, k) of a K-major [rows][16]-chunked SWIZZLE_64B operand: chunk k / 16, float offset inside the chunk
__device__ __forceinline__ int pm_off(int row, int k16) {
  return row * 16 + ((((k16 >> 2) ^ ((row >> 1) & 3)) << 2) | (k16 & 3));
}

__device__ __forceinline__ void pm_stage_weights(const float* __restrict__ w, int cout, int cin, int chunks,
                                                 float* hi, float* lo, int chunk_floats) {
  const int kpad = chunks * 16;
  for (int i = threadIdx.x; i < cout * kpad; i += PM_THREADS) {
    const int row = i / kpad, k = i - row * kpad;
    const float v = k < cin ? w[row * cin + k] : 0.f;
    const float h = pm_tf32(v);
    const int o = (k >> 4) * chunk_floats + pm_off(row, k & 15);
    hi[o] = h;
    lo[o] = v - h;
  }
}

__device__ __forceinline__ void pm_group_handoff(void* bar, int lane) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// relu(acc + bias) of 16 consecutive channels -> one 64-byte operand row segment (4 swizzled 16-byte pieces), hi and lo
__device__ __forceinline__ void pm_store16(const uint32_t* r, const float* bias, float* hi_chunk, float* lo_chunk,
                                           int row, int sw) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    float v[4], h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = __uint_as_float(r[4 * p + e]) + bias[4 * p + e];
      v[e] = a > 0.f ? a : 0.f;
      h[e] = pm_tf32(v[e]);
    }
    const int o = row * 16 + ((p ^ sw) << 2);
    *reinterpret_cast<float4*>(hi_chunk + o) = make_float4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<float4*>(lo_chunk + o) = make_float4(v[0] - h[0], v[1] - h[1], v[2] - h[2], v[3] - h[3]);
  }
}

// x [b][cin][m][ns] -> out [b][c3][m] = max_s relu(W3' relu(W2' relu(W1' x + b1') + b2') + b3')
// rows_per_batch = m * ns (multiple of 128); ns >= 32 with ns % 128 == 0 or 128 % ns == 0.
// `out` must be zero-filled when ns > 128 (several tiles per ball merge with atomicMax on the non-negative bit patterns).
__global__ void __launch_bounds__(PM_THREADS, 1)
k_shared_mlp_max(const float* __restrict__ x, int cin, int c1, int c2, int c3, int m, int ns, long long total_tiles,
                 const float* __restrict__ w1, const float* __restrict__ bb1, const float* __restrict__ w2,
                 const float* __restrict__ bb2, const float* __restrict__ w3, const float* __restrict__ bb3,
                 float* out) {
  extern __shared__ unsigned char smem_raw[];
  PmSmem& sm = *reinterpret_cast<PmSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rpb = (long long)m * ns;

  if (threadIdx.x == 0) {
    for (int g = 0; g < 2; ++g) { mbar_init(&sm.a_ready[g], 4); mbar_init(&sm.d_ready[g], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pm_stage_weights(w1, c1, cin, 1, sm.w1_hi, sm.w1_lo, PM_C1 * TC_BK);
  pm_stage_weights(w2, c2, c1, c1 / 16, &sm.w2_hi[0][0], &sm.w2_lo[0][0], PM_C2 * TC_BK);
  pm_stage_weights(w3, c3, c2, c2 / 16, &sm.w3_hi[0][0], &sm.w3_lo[0][0], PM_C3 * TC_BK);
  for (int t = threadIdx.x; t < c1; t += PM_THREADS) sm.b1[t] = bb1[t];
  for (int t = threadIdx.x; t < c2; t += PM_THREADS) sm.b2[t] = bb2[t];
  for (int t = threadIdx.x; t < c3; t += PM_THREADS) sm.b3[t] = bb3[t];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the staged weights are read by the UMMA (async proxy)
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 1) {
    // ===== MMA issuer: layer by layer, alternating between the two tile slots =====
    uint32_t par_a[2] = {0u, 0u};
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    for (long long j = 0;; ++j) {
      bool any = false;
#pragma unroll 1
      for (int L = 0; L < 3; ++L) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const long long t = blockIdx.x + (2 * j + g) * (long long)gridDim.x;
          if (t >= total_tiles) continue;
          any = true;
          mbar_wait(&sm.a_ready[g], par_a[g]);
          par_a[g] ^= 1u;
          tc_fence_after();
          if (lane == 0) {
            const int n_out = L == 0 ? c1 : (L == 1 ? c2 : c3);
            const int chunks = L == 0 ? 1 : (L == 1 ? c1 / 16 : c2 / 16);
            const int ksteps = L == 0 ? (cin > 8 ? 2 : 1) : 2;
            const uint32_t idesc = idesc_base | ((uint32_t)(n_out >> 3) << 17);
            const uint32_t d_tmem = tmem_base + g * 256 + (L == 0 ? 0 : (L == 1 ? PM_C1 : PM_C1 + PM_C2));
            for (int kc = 0; kc < chunks; ++kc) {
              const float* bh = L == 0 ? sm.w1_hi : (L == 1 ? sm.w2_hi[kc] : sm.w3_hi[kc]);
              const float* bl = L == 0 ? sm.w1_lo : (L == 1 ? sm.w2_lo[kc] : sm.w3_lo[kc]);
              const uint64_t ahi = make_desc_sw128(sm.a_hi[g][kc]), alo = make_desc_sw128(sm.a_lo[g][kc]);
              const uint64_t bhi = make_desc_sw128(bh), blo = make_desc_sw128(bl);
              for (int kk = 0; kk < ksteps; ++kk) {
                const uint64_t adv = (uint64_t)(kk * 8 * 4 >> 4);
                tc_mma_tf32(d_tmem, alo + adv, bhi + adv, idesc, (kc | kk) ? 1u : 0u);
                tc_mma_tf32(d_tmem, ahi + adv, blo + adv, idesc, 1u);
                tc_mma_tf32(d_tmem, ahi + adv, bhi + adv, idesc, 1u);
              }
            }
            tc_commit(&sm.d_ready[g]);
          }
          __syncwarp();
        }
      }
      if (!any) break;
    }
  } else if (warp >= 2) {
    // ===== tile-slot groups: thread = one sample (row of the tile = TMEM lane) =====
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int row = q * 32 + lane;
    const int sw = (row >> 1) & 3;
    const int gt = threadIdx.x - 64 - 128 * g;    // 0..127 inside the group (column owner in the final reduction)
    float* tr = &sm.a_hi[g][2][0] + (q & 1) * 1024 + (q >> 1) * PM_TILE;   // 32x32 transpose buffer (XOR-swizzled), free after L3
    const uint32_t tq = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    const int cpt = ns >= 128 ? 1 : 128 / ns;     // balls per tile
    const int qpc = 4 / cpt;                      // TMEM quarters per ball
    uint32_t par_d = 0;
    // the cin input channels of this thread's sample, loaded one tile AHEAD (the global-load latency would otherwise
    // sit at the head of every tile's serial load -> L1 -> ... -> L3 chain)
    float v[16];
    auto load_tile = [&](long long t) {
      const long long R0 = t * 128;
      const long long b = R0 / rpb, rib = R0 - b * rpb;
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = c < cin ? __ldg(x + ((size_t)b * cin + c) * rpb + rib + row) : 0.f;
    };
    {
      const long long t0 = blockIdx.x + (long long)g * gridDim.x;
      if (t0 < total_tiles) load_tile(t0);
    }
    for (long long j = 0;; ++j) {
      const long long t = blockIdx.x + (2 * j + g) * (long long)gridDim.x;
      if (t >= total_tiles) break;
      const long long R0 = t * 128;
      const long long b = R0 / rpb, rib = R0 - b * rpb;
      // ---- A1 (coalesced loads: consecutive rows are consecutive floats per channel)
      {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = pm_tf32(v[4 * p + e]);
          const int o = row * 16 + ((p ^ sw) << 2);
          *reinterpret_cast<float4*>(&sm.a_hi[g][0][o]) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(&sm.a_lo[g][0][o]) =
              make_float4(v[4 * p] - h[0], v[4 * p + 1] - h[1], v[4 * p + 2] - h[2], v[4 * p + 3] - h[3]);
        }
        const long long tn = t + 2 * (long long)gridDim.x;
        if (tn < total_tiles) load_tile(tn);
      }
      pm_group_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 1: D1 (c1 columns) -> A2
      mbar_wait(&sm.d_ready[g], par_d);
      par_d ^= 1u;
      tc_fence_after();
      {
        uint32_t r[32];
        tmem_ld_32x32(tq, r);
#pragma unroll
        for (int kc = 0; kc < PM_C1 / 16; ++kc)
          if (kc * 16 < c1) pm_store16(r + 16 * kc, sm.b1 + 16 * kc, sm.a_hi[g][kc], sm.a_lo[g][kc], row, sw);
      }
      pm_group_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 2: D2 (c2 columns) -> A3
      mbar_wait(&sm.d_ready[g], par_d);
      par_d ^= 1u;
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb * 32 < c2; ++cb) {
        uint32_t r[32];
        tmem_ld_32x32(tq + PM_C1 + cb * 32, r);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int kc = 2 * cb + h2;
          if (kc * 16 < c2) pm_store16(r + 16 * h2, sm.b2 + 16 * kc, sm.a_hi[g][kc], sm.a_lo[g][kc], row, sw);
        }
      }
      pm_group_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 3: D3 (c3 columns) -> relu -> max over the 32 samples of this quarter -> red[q][col]
      mbar_wait(&sm.d_ready[g], par_d);
      par_d ^= 1u;
      tc_fence_after();
#pragma unroll 1
      for (int cb = 0; cb * 32 < c3; ++cb) {
        uint32_t r[32];
        tmem_ld_32x32(tq + PM_C1 + PM_C2 + cb * 32, r);
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float a = __uint_as_float(r[jj]) + sm.b3[cb * 32 + jj];
          tr[lane * 32 + (jj ^ lane)] = a > 0.f ? a : 0.f;
        }
        __syncwarp();
        float mx = 0.f;
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) mx = fmaxf(mx, tr[rr * 32 + (lane ^ rr)]);
        sm.red[g][q][cb * 32 + lane] = mx;
        __syncwarp();
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");     // the four quarters of this group's tile
      if (gt < c3) {
        for (int s = 0; s < cpt; ++s) {
          float mx = sm.red[g][s * qpc][gt];
          for (int qq = 1; qq < qpc; ++qq) mx = fmaxf(mx, sm.red[g][s * qpc + qq][gt]);
          const long long ball = (rib + (long long)s * (128 / cpt)) / ns;
          float* dst = out + ((size_t)b * c3 + gt) * m + ball;
          if (ns <= 128) *dst = mx;
          else atomicMax(reinterpret_cast<int*>(dst), __float_as_int(mx));
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");     // red[] is rewritten by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Four tile slots (round 2).  With two slots the tensor pipe was 20-23 % active: a tile's chain
//   load -> [L1] -> epilogue -> [L2] -> epilogue -> [L3] -> epilogue
// is serial (each arrow is a commit -> mbarrier -> TMEM-load round trip) and two chains hide only half of it.  Operand
// space, not threads, was the limit: A3 (128 x 64, hi + lo) is 64 KB per slot.  Here L3 runs in two K halves, so a slot
// only ever holds TWO 16-wide operand chunks (32 KB): epilogue 2 writes A3 chunks 0-1, keeps the other 32 activations of
// its row in registers, and writes chunks 2-3 into the same ring once L3a has retired.  The accumulators of a slot
// share 128 TMEM columns (D1 [0,32), D2 [32,96), D3 [0,128) — D3 is first written after D2 has been read completely).
// 4 slots x 32 KB + 84 KB of weights = 212 KB of shared memory, 4 x 128 = 512 TMEM columns, 18 warps; the MMA warp
// polls the four slots' barriers and issues whatever is ready.  The max over the ball uses REDUX on the non-negative
// bit patterns instead of a shared-memory transpose.
constexpr int PM_SLOTS = 4;
constexpr int PM4_THREADS = 64 + PM_SLOTS * 128;

struct __align__(1024) Pm4Smem {
  float w1_hi[PM_C1 * TC_BK], w1_lo[PM_C1 * TC_BK];
  float w2_hi[PM_C1 / 16][PM_C2 * TC_BK], w2_lo[PM_C1 / 16][PM_C2 * TC_BK];
  float w3_hi[PM_C2 / 16][PM_C3 * TC_BK], w3_lo[PM_C2 / 16][PM_C3 * TC_BK];
  float a_hi[PM_SLOTS][2][PM_TILE], a_lo[PM_SLOTS][2][PM_TILE];          // per slot: a ring of two operand chunks
  float red[PM_SLOTS][4][PM_C3];
  float b1[PM_C1], b2[PM_C2], b3[PM_C3];
  unsigned long long a_ready[PM_SLOTS], d_ready[PM_SLOTS];
  uint32_t tmem_base;
};
static_assert(sizeof(Pm4Smem) + 1024 <= 227 * 1024, "Pm4Smem must fit the 227 KB of a CTA");

__device__ __forceinline__ bool mbar_test(void* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

template <bool CIN16>
__global__ void __launch_bounds__(PM4_THREADS, 1)
k_shared_mlp_max4(const float* __restrict__ x, int cin, int c1, int c2, int c3, int m, int ns, long long total_tiles,
                  const float* __restrict__ w1, const float* __restrict__ bb1, const float* __restrict__ w2,
                  const float* __restrict__ bb2, const float* __restrict__ w3, const float* __restrict__ bb3,
                  float* out) {
  extern __shared__ unsigned char smem_raw[];
  Pm4Smem& sm = *reinterpret_cast<Pm4Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rpb = (long long)m * ns;
  const int k3 = c2 / 16;                 // A3 chunks: 2 or 4
  const bool two_halves = k3 > 2;         // L3 in two K halves (both sides of the hand-off derive this identically)

  if (threadIdx.x == 0) {
    for (int g = 0; g < PM_SLOTS; ++g) { mbar_init(&sm.a_ready[g], 4); mbar_init(&sm.d_ready[g], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    // stage the folded weights (hi + lo, UMMA SWIZZLE_64B layout); same helper as the two-slot kernel, other stride
    auto stage = [&](const float* w, int cout, int kin, int chunks, float* hi, float* lo, int chunk_floats) {
      const int kpad = chunks * 16;
      for (int i = threadIdx.x; i < cout * kpad; i += PM4_THREADS) {
        const int row = i / kpad, k = i - row * kpad;
        const float v = k < kin ? w[row * kin + k] : 0.f;
        const float h = pm_tf32(v);
        const int o = (k >> 4) * chunk_floats + pm_off(row, k & 15);
        hi[o] = h;
        lo[o] = v - h;
      }
    };
    stage(w1, c1, cin, 1, sm.w1_hi, sm.w1_lo, PM_C1 * TC_BK);
    stage(w2, c2, c1, c1 / 16, &sm.w2_hi[0][0], &sm.w2_lo[0][0], PM_C2 * TC_BK);
    stage(w3, c3, c2, c2 / 16, &sm.w3_hi[0][0], &sm.w3_lo[0][0], PM_C3 * TC_BK);
  }
  for (int t = threadIdx.x; t < c1; t += PM4_THREADS) sm.b1[t] = bb1[t];
  for (int t = threadIdx.x; t < c2; t += PM4_THREADS) sm.b2[t] = bb2[t];
  for (int t = threadIdx.x; t < c3; t += PM4_THREADS) sm.b3[t] = bb3[t];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 1) {
    // ===== MMA issuer: polls the slots, issues the next phase of whichever slot has handed its operand over =====
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const int nph = two_halves ? 4 : 3;
    uint32_t par_a[PM_SLOTS];
    int phase[PM_SLOTS];
    long long tile[PM_SLOTS];
    int live = 0;
#pragma unroll
    for (int g = 0; g < PM_SLOTS; ++g) {
      par_a[g] = 0u;
      phase[g] = 0;
      tile[g] = blockIdx.x + (long long)g * gridDim.x;
      live += tile[g] < total_tiles;
    }
    while (live > 0) {
#pragma unroll
      for (int g = 0; g < PM_SLOTS; ++g) {
        if (tile[g] >= total_tiles) continue;
        if (!mbar_test(&sm.a_ready[g], par_a[g])) continue;
        par_a[g] ^= 1u;
        tc_fence_after();
        const int ph = phase[g];
        if (lane == 0) {
          const int n_out = ph == 0 ? c1 : (ph == 1 ? c2 : c3);
          const uint32_t idesc = idesc_base | ((uint32_t)(n_out >> 3) << 17);
          const uint32_t d_tmem = tmem_base + g * 128 + (ph == 1 ? PM_C1 : 0);
          const int kc0 = ph == 3 ? 2 : 0;
          const int kc1 = ph == 0 ? 1 : (ph == 1 ? c1 / 16 : (ph == 2 ? (k3 < 2 ? k3 : 2) : k3));
          const int ksteps = ph == 0 ? (cin > 8 ? 2 : 1) : 2;
          for (int kc = kc0; kc < kc1; ++kc) {
            const float* bh = ph == 0 ? sm.w1_hi : (ph == 1 ? sm.w2_hi[kc] : sm.w3_hi[kc]);
            const float* bl = ph == 0 ? sm.w1_lo : (ph == 1 ? sm.w2_lo[kc] : sm.w3_lo[kc]);
            const uint64_t ahi = make_desc_sw128(sm.a_hi[g][kc - kc0]), alo = make_desc_sw128(sm.a_lo[g][kc - kc0]);
            const uint64_t bhi = make_desc_sw128(bh), blo = make_desc_sw128(bl);
            for (int kk = 0; kk < ksteps; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 8 * 4 >> 4);
              tc_mma_tf32(d_tmem, alo + adv, bhi + adv, idesc, (ph == 3 || kc > kc0 || kk) ? 1u : 0u);
              tc_mma_tf32(d_tmem, ahi + adv, blo + adv, idesc, 1u);
              tc_mma_tf32(d_tmem, ahi + adv, bhi + adv, idesc, 1u);
            }
          }
          tc_commit(&sm.d_ready[g]);
        }
        __syncwarp();
        if (++phase[g] == nph) {
          phase[g] = 0;
          tile[g] += (long long)PM_SLOTS * gridDim.x;
          live -= tile[g] >= total_tiles;
        }
      }
    }
  } else if (warp >= 2) {
    // ===== tile-slot groups: thread = one sample (row of the tile = TMEM lane) =====
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int sw = (row >> 1) & 3;
    const int gt = threadIdx.x - 64 - 128 * g;
    const uint32_t tq = tmem_base + g * 128 + ((uint32_t)(q * 32) << 16);
    const int cpt = ns >= 128 ? 1 : 128 / ns;
    const int qpc = 4 / cpt;
    uint32_t par_d = 0;
    constexpr int NV = CIN16 ? 16 : 8;
    float v[NV];
    auto load_tile = [&](long long t) {
      const long long R0 = t * 128;
      const long long b = R0 / rpb, rib = R0 - b * rpb;
#pragma unroll
      for (int c = 0; c < NV; ++c) v[c] = c < cin ? __ldg(x + ((size_t)b * cin + c) * rpb + rib + row) : 0.f;
    };
    auto wait_d = [&]() {
      mbar_wait(&sm.d_ready[g], par_d);
      par_d ^= 1u;
      tc_fence_after();
    };
    {
      const long long t0 = blockIdx.x + (long long)g * gridDim.x;
      if (t0 < total_tiles) load_tile(t0);
    }
    for (long long t = blockIdx.x + (long long)g * gridDim.x; t < total_tiles; t += (long long)PM_SLOTS * gridDim.x) {
      const long long R0 = t * 128;
      const long long b = R0 / rpb, rib = R0 - b * rpb;
      // ---- A1 -> ring chunk 0
      {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float h[4], a[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            a[e] = (4 * p + e < NV) ? v[(4 * p + e) % NV] : 0.f;
            h[e] = pm_tf32(a[e]);
          }
          const int o = row * 16 + ((p ^ sw) << 2);
          *reinterpret_cast<float4*>(&sm.a_hi[g][0][o]) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(&sm.a_lo[g][0][o]) = make_float4(a[0] - h[0], a[1] - h[1], a[2] - h[2], a[3] - h[3]);
        }
        const long long tn = t + (long long)PM_SLOTS * gridDim.x;
        if (tn < total_tiles) load_tile(tn);
      }
      pm_group_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 1: D1 -> A2 (ring chunks 0..c1/16-1)
      wait_d();
      {
        uint32_t r[32];
        tmem_ld_32x32(tq, r);
#pragma unroll
        for (int kc = 0; kc < PM_C1 / 16; ++kc)
          if (kc * 16 < c1) pm_store16(r + 16 * kc, sm.b1 + 16 * kc, sm.a_hi[g][kc], sm.a_lo[g][kc], row, sw);
      }
      pm_group_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 2: D2 -> A3 chunks 0-1 now, chunks 2-3 (kept in registers) once L3a has retired
      wait_d();
      {
        uint32_t r[32], r2[32];
        tmem_ld_32x32(tq + PM_C1, r);
        if (two_halves) tmem_ld_32x32(tq + PM_C1 + 32, r2);
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) pm_store16(r + 16 * h2, sm.b2 + 16 * h2, sm.a_hi[g][h2], sm.a_lo[g][h2], row, sw);
        pm_group_handoff(&sm.a_ready[g], lane);
        if (two_halves) {
          wait_d();                                   // L3a has read ring chunks 0-1 (and may have overwritten D2)
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2)
            pm_store16(r2 + 16 * h2, sm.b2 + 32 + 16 * h2, sm.a_hi[g][h2], sm.a_lo[g][h2], row, sw);
          pm_group_handoff(&sm.a_ready[g], lane);
        }
      }
      // ---- epilogue 3: D3 -> relu -> max over the 32 samples of this quarter (REDUX on the bit patterns) -> red[q][col]
      wait_d();
#pragma unroll 1
      for (int cb = 0; cb * 32 < c3; ++cb) {
        uint32_t r[32];
        tmem_ld_32x32(tq + cb * 32, r);
        unsigned mine = 0u;
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) {
          const float a = __uint_as_float(r[jj]) + sm.b3[cb * 32 + jj];
          const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(a > 0.f ? a : 0.f));
          if (jj == lane) mine = mx;
        }
        sm.red[g][q][cb * 32 + lane] = __uint_as_float(mine);
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (gt < c3) {
        for (int s = 0; s < cpt; ++s) {
          float mx = sm.red[g][s * qpc][gt];
          for (int qq = 1; qq < qpc; ++qq) mx = fmaxf(mx, sm.red[g][s * qpc + qq][gt]);
          const long long ball = (rib + (long long)s * (128 / cpt)) / ns;
          float* dst = out + ((size_t)b * c3 + gt) * m + ball;
          if (ns <= 128) *dst = mx;
          else atomicMax(reinterpret_cast<int*>(dst), __float_as_int(mx));
        }
      }
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Activations in TENSOR MEMORY (round 2, `k_shared_mlp_max_ts`).  The four-slot kernel above is bound by shared-memory
// bandwidth: every MMA re-reads its A operand (the activations, hi + lo) from shared memory and every epilogue writes the
// next layer's activations there — 391 KB per 128-sample tile against 1968 cycles of tensor work.  tcgen05.mma can take
// its A operand from tensor memory (the "TS" form: D[tmem] = A[tmem] * B[smem desc]; A is laid out like an
// accumulator, row = lane, K along the columns, one 32-bit column per tf32 element).  So the activations never leave
// TMEM: an epilogue reads the accumulator (tcgen05.ld), applies bias / ReLU / the hi-lo split in registers and writes
// the next layer's A operand back with tcgen05.st (256 B/clk).  Shared memory only serves the weights (123 KB of
// reads per tile instead of 391 KB).
//
// TMEM columns of a tile slot (256 per slot, two slots):
//   [0,128)   accumulators: D1 [0,32), D2 [32,96), D3 [0,128) (D3 is first written after D2 has been read completely)
//   [128,192) operand ring, two 16-wide K chunks: hi0 [128,144) hi1 [144,160) lo0 [160,176) lo1 [176,192)
// L3 runs in two K halves through the ring exactly like in k_shared_mlp_max4.  EIGHT warps per slot: two per TMEM lane
// quarter, each owning half of the columns of every epilogue, so a slot's serial chain is half as long.
constexpr int PT_SLOTS = 2;
constexpr int PT_THREADS = 64 + PT_SLOTS * 256;
constexpr int PT_A_HI = 128, PT_A_LO = 160;

struct __align__(1024) PtSmem {
  float w1_hi[PM_C1 * TC_BK], w1_lo[PM_C1 * TC_BK];
  float w2_hi[PM_C1 / 16][PM_C2 * TC_BK], w2_lo[PM_C1 / 16][PM_C2 * TC_BK];
  float w3_hi[PM_C2 / 16][PM_C3 * TC_BK], w3_lo[PM_C2 / 16][PM_C3 * TC_BK];
  float red[PT_SLOTS][4][PM_C3];
  float b1[PM_C1], b2[PM_C2], b3[PM_C3];
  unsigned long long a_ready[PT_SLOTS], d_ready[PT_SLOTS];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// relu(acc + bias) of 16 consecutive channels -> the hi and lo halves of one operand chunk in tensor memory
__device__ __forceinline__ void pt_store16(const uint32_t (&r)[16], const float* bias, uint32_t t_hi, uint32_t t_lo) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    const float a = __uint_as_float(r[e]) + bias[e];
    const float v = a > 0.f ? a : 0.f;
    const float hv = pm_tf32(v);
    h[e] = __float_as_uint(hv);
    l[e] = __float_as_uint(v - hv);
  }
  tmem_st_32x16(t_hi, h);
  tmem_st_32x16(t_lo, l);
}

__device__ __forceinline__ void pt_handoff(void* bar, int lane) {
  tmem_st_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__global__ void __launch_bounds__(PT_THREADS, 1)
k_shared_mlp_max_ts(const float* __restrict__ x, int cin, int c1, int c2, int c3, int m, int ns, long long total_tiles,
                    const float* __restrict__ w1, const float* __restrict__ bb1, const float* __restrict__ w2,
                    const float* __restrict__ bb2, const float* __restrict__ w3, const float* __restrict__ bb3,
                    float* out) {
  extern __shared__ unsigned char smem_raw[];
  PtSmem& sm = *reinterpret_cast<PtSmem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rpb = (long long)m * ns;
  const int k3 = c2 / 16;
  const bool two_halves = k3 > 2;

  if (threadIdx.x == 0) {
    for (int g = 0; g < PT_SLOTS; ++g) { mbar_init(&sm.a_ready[g], 8); mbar_init(&sm.d_ready[g], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    auto stage = [&](const float* w, int cout, int kin, int chunks, float* hi, float* lo, int chunk_floats) {
      const int kpad = chunks * 16;
      for (int i = threadIdx.x; i < cout * kpad; i += PT_THREADS) {
        const int row = i / kpad, k = i - row * kpad;
        const float v = k < kin ? w[row * kin + k] : 0.f;
        const float h = pm_tf32(v);
        const int o = (k >> 4) * chunk_floats + pm_off(row, k & 15);
        hi[o] = h;
        lo[o] = v - h;
      }
    };
    stage(w1, c1, cin, 1, sm.w1_hi, sm.w1_lo, PM_C1 * TC_BK);
    stage(w2, c2, c1, c1 / 16, &sm.w2_hi[0][0], &sm.w2_lo[0][0], PM_C2 * TC_BK);
    stage(w3, c3, c2, c2 / 16, &sm.w3_hi[0][0], &sm.w3_lo[0][0], PM_C3 * TC_BK);
  }
  for (int t = threadIdx.x; t < c1; t += PT_THREADS) sm.b1[t] = bb1[t];
  for (int t = threadIdx.x; t < c2; t += PT_THREADS) sm.b2[t] = bb2[t];
  for (int t = threadIdx.x; t < c3; t += PT_THREADS) sm.b3[t] = bb3[t];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 1) {
    // ===== MMA issuer (TS form: A from tensor memory, B = the resident weights) =====
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    const int nph = two_halves ? 4 : 3;
    uint32_t par_a[PT_SLOTS];
    int phase[PT_SLOTS];
    long long tile[PT_SLOTS];
    int live = 0;
#pragma unroll
    for (int g = 0; g < PT_SLOTS; ++g) {
      par_a[g] = 0u;
      phase[g] = 0;
      tile[g] = blockIdx.x + (long long)g * gridDim.x;
      live += tile[g] < total_tiles;
    }
    while (live > 0) {
#pragma unroll
      for (int g = 0; g < PT_SLOTS; ++g) {
        if (tile[g] >= total_tiles) continue;
        if (!mbar_test(&sm.a_ready[g], par_a[g])) continue;
        par_a[g] ^= 1u;
        tc_fence_after();
        const int ph = phase[g];
        if (lane == 0) {
          const int n_out = ph == 0 ? c1 : (ph == 1 ? c2 : c3);
          const uint32_t idesc = idesc_base | ((uint32_t)(n_out >> 3) << 17);
          const uint32_t slot = tmem_base + g * 256;
          const uint32_t d_tmem = slot + (ph == 1 ? PM_C1 : 0);
          const int kc0 = ph == 3 ? 2 : 0;
          const int kc1 = ph == 0 ? 1 : (ph == 1 ? c1 / 16 : (ph == 2 ? (k3 < 2 ? k3 : 2) : k3));
          const int ksteps = ph == 0 ? (cin > 8 ? 2 : 1) : 2;
          for (int kc = kc0; kc < kc1; ++kc) {
            const float* bh = ph == 0 ? sm.w1_hi : (ph == 1 ? sm.w2_hi[kc] : sm.w3_hi[kc]);
            const float* bl = ph == 0 ? sm.w1_lo : (ph == 1 ? sm.w2_lo[kc] : sm.w3_lo[kc]);
            const uint64_t bhi = make_desc_sw128(bh), blo = make_desc_sw128(bl);
            const uint32_t ahi = slot + PT_A_HI + (kc - kc0) * 16, alo = slot + PT_A_LO + (kc - kc0) * 16;
            for (int kk = 0; kk < ksteps; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 8 * 4 >> 4);
              tc_mma_tf32_ts(d_tmem, alo + kk * 8, bhi + adv, idesc, (ph == 3 || kc > kc0 || kk) ? 1u : 0u);
              tc_mma_tf32_ts(d_tmem, ahi + kk * 8, blo + adv, idesc, 1u);
              tc_mma_tf32_ts(d_tmem, ahi + kk * 8, bhi + adv, idesc, 1u);
            }
          }
          tc_commit(&sm.d_ready[g]);
        }
        __syncwarp();
        if (++phase[g] == nph) {
          phase[g] = 0;
          tile[g] += (long long)PT_SLOTS * gridDim.x;
          live -= tile[g] >= total_tiles;
        }
      }
    }
  } else if (warp >= 2) {
    // ===== tile-slot groups: 8 warps per slot; thread = one sample (TMEM lane), warp pair (q, half) splits the columns
    const int g = (warp - 2) >> 3;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int half = ((warp - 2) >> 2) & 1;
    const int row = q * 32 + lane;
    const int gt = threadIdx.x - 64 - 256 * g;    // 0..255 inside the slot's group
    const uint32_t slot = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    const int cpt = ns >= 128 ? 1 : 128 / ns;
    const int qpc = 4 / cpt;
    uint32_t par_d = 0;
    float v[16];
    auto load_tile = [&](long long t) {
      const long long R0 = t * 128;
      const long long b = R0 / rpb, rib = R0 - b * rpb;
#pragma unroll
      for (int c = 0; c < 16; ++c) v[c] = c < cin ? __ldg(x + ((size_t)b * cin + c) * rpb + rib + row) : 0.f;
    };
    auto wait_d = [&]() {
      mbar_wait(&sm.d_ready[g], par_d);
      par_d ^= 1u;
      tc_fence_after();
    };
    {
      const long long t0 = blockIdx.x + (long long)g * gridDim.x;
      if (t0 < total_tiles) load_tile(t0);
    }
    for (long long t = blockIdx.x + (long long)g * gridDim.x; t < total_tiles; t += (long long)PT_SLOTS * gridDim.x) {
      const long long R0 = t * 128;
      const long long b = R0 / rpb, rib = R0 - b * rpb;
      // ---- A1 -> ring chunk 0: half 0 writes the hi part, half 1 the lo part
      {
        uint32_t o[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float hv = pm_tf32(v[e]);
          o[e] = __float_as_uint(half == 0 ? hv : v[e] - hv);
        }
        tmem_st_32x16(slot + (half == 0 ? PT_A_HI : PT_A_LO), o);
        const long long tn = t + (long long)PT_SLOTS * gridDim.x;
        if (tn < total_tiles) load_tile(tn);
      }
      pt_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 1: D1 -> A2; warp `half` owns chunk `half`
      wait_d();
      if (half * 16 < c1) {
        uint32_t r[16];
        tmem_ld_32x16(slot + half * 16, r);
        pt_store16(r, sm.b1 + half * 16, slot + PT_A_HI + half * 16, slot + PT_A_LO + half * 16);
      }
      pt_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 2: D2 -> A3; warp `half` owns chunks `half` (now) and `half + 2` (after L3a)
      wait_d();
      {
        uint32_t r[16], r2[16];
        tmem_ld_32x16(slot + PM_C1 + half * 16, r);
        if (two_halves) tmem_ld_32x16(slot + PM_C1 + 32 + half * 16, r2);
        pt_store16(r, sm.b2 + half * 16, slot + PT_A_HI + half * 16, slot + PT_A_LO + half * 16);
        pt_handoff(&sm.a_ready[g], lane);
        if (two_halves) {
          wait_d();
          pt_store16(r2, sm.b2 + 32 + half * 16, slot + PT_A_HI + half * 16, slot + PT_A_LO + half * 16);
          pt_handoff(&sm.a_ready[g], lane);
        }
      }
      // ---- epilogue 3: D3 -> relu -> max over the 32 samples of this quarter; warp `half` owns columns [64 half, +64)
      wait_d();
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        const int col0 = half * 64 + cb * 16;
        if (col0 >= c3) break;
        uint32_t r[16];
        tmem_ld_32x16(slot + col0, r);
        unsigned mine = 0u;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const float a = __uint_as_float(r[jj]) + sm.b3[col0 + jj];
          const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(a > 0.f ? a : 0.f));
          if (jj == (lane & 15)) mine = mx;
        }
        if (lane < 16) sm.red[g][q][col0 + lane] = __uint_as_float(mine);
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
      if (gt < c3) {
        for (int s = 0; s < cpt; ++s) {
          float mx = sm.red[g][s * qpc][gt];
          for (int qq = 1; qq < qpc; ++qq) mx = fmaxf(mx, sm.red[g][s * qpc + qq][gt]);
          const long long ball = (rib + (long long)s * (128 / cpt)) / ns;
          float* dst = out + ((size_t)b * c3 + gt) * m + ball;
          if (ns <= 128) *dst = mx;
          else atomicMax(reinterpret_cast<int*>(dst), __float_as_int(mx));
        }
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}


// ------------------------------------------------------------------------------------------------------------------
// `k_shared_mlp_max_ts2`: two round trips per tile instead of five.  Measured on the kernel above: with the activations
// in tensor memory the tile time did NOT move (4.0 us) — neither shared-memory bandwidth nor the tensor pipe (26 %
// active) was the limit, the serial chain of a tile was: five MMA <-> epilogue round trips of ~1 us each (hand-off,
// issue, execute, commit, wake, TMEM load, math, TMEM store).  So the chain is shortened instead:
//   * layer 1 (cin <= 16 inputs, 32 outputs: 192 MACs per sample at cin = 6) runs on the CUDA cores in exact fp32, in the
//     thread that already holds the sample's inputs — no L1 MMA, no epilogue 1;
//   * with the operands in TMEM there is room for the WHOLE A3 (64 + 64 columns) next to the accumulators
//     (128 columns): L3 is one phase again.  Slot = 256 columns: D2 [0,64) / D3 [0,128) | A hi [128,192) | A lo [192,256).
// A tile is now: inputs -> A2 (registers -> TMEM) -> [L2] -> epilogue 2 (A3 -> TMEM) -> [L3] -> epilogue 3 (max).
struct __align__(1024) Pt2Smem {
  float w2_hi[PM_C1 / 16][PM_C2 * TC_BK], w2_lo[PM_C1 / 16][PM_C2 * TC_BK];
  float w3_hi[PM_C2 / 16][PM_C3 * TC_BK], w3_lo[PM_C2 / 16][PM_C3 * TC_BK];
  float w1f[PM_C1][16];                       // layer-1 weights, plain fp32, inputs padded to 16
  float red[PT_SLOTS][4][PM_C3];
  float b1[PM_C1], b2[PM_C2], b3[PM_C3];
  unsigned long long a_ready[PT_SLOTS], d_ready[PT_SLOTS];
  uint32_t tmem_base;
};
constexpr int PT2_A_HI = 128, PT2_A_LO = 192;

template <int CINP>   // inputs padded to CINP (8 or 16)
__global__ void __launch_bounds__(PT_THREADS, 1)
k_shared_mlp_max_ts2(const float* __restrict__ x, int cin, int c1, int c2, int c3, int m, int ns, long long total_tiles,
                     const float* __restrict__ w1, const float* __restrict__ bb1, const float* __restrict__ w2,
                     const float* __restrict__ bb2, const float* __restrict__ w3, const float* __restrict__ bb3,
                     float* out) {
  extern __shared__ unsigned char smem_raw[];
  Pt2Smem& sm = *reinterpret_cast<Pt2Smem*>(smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long rpb = (long long)m * ns;
  const int k2 = c1 / 16, k3 = c2 / 16;

  if (threadIdx.x == 0) {
    for (int g = 0; g < PT_SLOTS; ++g) { mbar_init(&sm.a_ready[g], 8); mbar_init(&sm.d_ready[g], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  {
    auto stage = [&](const float* w, int cout, int kin, int chunks, float* hi, float* lo, int chunk_floats) {
      const int kpad = chunks * 16;
      for (int i = threadIdx.x; i < cout * kpad; i += PT_THREADS) {
        const int row = i / kpad, k = i - row * kpad;
        const float v = k < kin ? w[row * kin + k] : 0.f;
        const float h = pm_tf32(v);
        const int o = (k >> 4) * chunk_floats + pm_off(row, k & 15);
        hi[o] = h;
        lo[o] = v - h;
      }
    };
    stage(w2, c2, c1, k2, &sm.w2_hi[0][0], &sm.w2_lo[0][0], PM_C2 * TC_BK);
    stage(w3, c3, c2, k3, &sm.w3_hi[0][0], &sm.w3_lo[0][0], PM_C3 * TC_BK);
    for (int i = threadIdx.x; i < PM_C1 * 16; i += PT_THREADS) {
      const int o = i >> 4, c = i & 15;
      sm.w1f[o][c] = (o < c1 && c < cin) ? w1[o * cin + c] : 0.f;
    }
  }
  for (int t = threadIdx.x; t < PM_C1; t += PT_THREADS) sm.b1[t] = t < c1 ? bb1[t] : 0.f;
  for (int t = threadIdx.x; t < c2; t += PT_THREADS) sm.b2[t] = bb2[t];
  for (int t = threadIdx.x; t < c3; t += PT_THREADS) sm.b3[t] = bb3[t];
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 1) {
    // ===== MMA issuer: two phases per tile (L2, L3), TS form =====
    const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BM >> 4) << 24);
    uint32_t par_a[PT_SLOTS];
    int phase[PT_SLOTS];
    long long tile[PT_SLOTS];
    int live = 0;
#pragma unroll
    for (int g = 0; g < PT_SLOTS; ++g) {
      par_a[g] = 0u;
      phase[g] = 0;
      tile[g] = blockIdx.x + (long long)g * gridDim.x;
      live += tile[g] < total_tiles;
    }
    while (live > 0) {
#pragma unroll
      for (int g = 0; g < PT_SLOTS; ++g) {
        if (tile[g] >= total_tiles) continue;
        if (!mbar_test(&sm.a_ready[g], par_a[g])) continue;
        par_a[g] ^= 1u;
        tc_fence_after();
        const int ph = phase[g];
        if (lane == 0) {
          const int n_out = ph == 0 ? c2 : c3;
          const uint32_t idesc = idesc_base | ((uint32_t)(n_out >> 3) << 17);
          const uint32_t slot = tmem_base + g * 256;
          const int kc1 = ph == 0 ? k2 : k3;
          for (int kc = 0; kc < kc1; ++kc) {
            const float* bh = ph == 0 ? sm.w2_hi[kc] : sm.w3_hi[kc];
            const float* bl = ph == 0 ? sm.w2_lo[kc] : sm.w3_lo[kc];
            const uint64_t bhi = make_desc_sw128(bh), blo = make_desc_sw128(bl);
            const uint32_t ahi = slot + PT2_A_HI + kc * 16, alo = slot + PT2_A_LO + kc * 16;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 8 * 4 >> 4);
              tc_mma_tf32_ts(slot, alo + kk * 8, bhi + adv, idesc, (kc | kk) ? 1u : 0u);
              tc_mma_tf32_ts(slot, ahi + kk * 8, blo + adv, idesc, 1u);
              tc_mma_tf32_ts(slot, ahi + kk * 8, bhi + adv, idesc, 1u);
            }
          }
          tc_commit(&sm.d_ready[g]);
        }
        __syncwarp();
        if (++phase[g] == 2) {
          phase[g] = 0;
          tile[g] += (long long)PT_SLOTS * gridDim.x;
          live -= tile[g] >= total_tiles;
        }
      }
    }
  } else if (warp >= 2) {
    const int g = (warp - 2) >> 3;
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row = q * 32 + lane;
    const int gt = threadIdx.x - 64 - 256 * g;
    const uint32_t slot = tmem_base + g * 256 + ((uint32_t)(q * 32) << 16);
    const int cpt = ns >= 128 ? 1 : 128 / ns;
    const int qpc = 4 / cpt;
    uint32_t par_d = 0;
    float v[CINP];
    auto load_tile = [&](long long t) {
      const long long R0 = t * 128;
      const long long b = R0 / rpb, rib = R0 - b * rpb;
#pragma unroll
      for (int c = 0; c < CINP; ++c) v[c] = c < cin ? __ldg(x + ((size_t)b * cin + c) * rpb + rib + row) : 0.f;
    };
    auto wait_d = [&]() {
      mbar_wait(&sm.d_ready[g], par_d);
      par_d ^= 1u;
      tc_fence_after();
    };
    {
      const long long t0 = blockIdx.x + (long long)g * gridDim.x;
      if (t0 < total_tiles) load_tile(t0);
    }
    for (long long t = blockIdx.x + (long long)g * gridDim.x; t < total_tiles; t += (long long)PT_SLOTS * gridDim.x) {
      const long long R0 = t * 128;
      const long long b = R0 / rpb, rib = R0 - b * rpb;
      // ---- layer 1 on the CUDA cores (exact fp32): this warp's 16 of the c1 channels -> A2 chunk `half`
      if (half * 16 < c1) {
        uint32_t h[16], l[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) {
          const float4* wr = reinterpret_cast<const float4*>(&sm.w1f[half * 16 + o][0]);
          float acc = sm.b1[half * 16 + o];
#pragma unroll
          for (int c4 = 0; c4 < CINP / 4; ++c4) {
            const float4 w = wr[c4];
            acc = fmaf(w.x, v[4 * c4], acc);
            acc = fmaf(w.y, v[4 * c4 + 1], acc);
            acc = fmaf(w.z, v[4 * c4 + 2], acc);
            acc = fmaf(w.w, v[4 * c4 + 3], acc);
          }
          const float a = acc > 0.f ? acc : 0.f;
          const float hv = pm_tf32(a);
          h[o] = __float_as_uint(hv);
          l[o] = __float_as_uint(a - hv);
        }
        tmem_st_32x16(slot + PT2_A_HI + half * 16, h);
        tmem_st_32x16(slot + PT2_A_LO + half * 16, l);
      }
      {
        const long long tn = t + (long long)PT_SLOTS * gridDim.x;
        if (tn < total_tiles) load_tile(tn);
      }
      pt_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 2: D2 -> A3; warp `half` owns chunks `half` and `half + 2`
      wait_d();
#pragma unroll 1
      for (int kc = half; kc < k3; kc += 2) {
        uint32_t r[16];
        tmem_ld_32x16(slot + kc * 16, r);
        pt_store16(r, sm.b2 + kc * 16, slot + PT2_A_HI + kc * 16, slot + PT2_A_LO + kc * 16);
      }
      pt_handoff(&sm.a_ready[g], lane);
      // ---- epilogue 3: D3 -> relu -> max over the 32 samples of this quarter; warp `half` owns columns [64 half, +64)
      wait_d();
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        const int col0 = half * 64 + cb * 16;
        if (col0 >= c3) break;
        uint32_t r[16];
        tmem_ld_32x16(slot + col0, r);
        unsigned mine = 0u;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
          const float a = __uint_as_float(r[jj]) + sm.b3[col0 + jj];
          const unsigned mx = __reduce_max_sync(0xffffffffu, __float_as_uint(a > 0.f ? a : 0.f));
          if (jj == (lane & 15)) mine = mx;
        }
        if (lane < 16) sm.red[g][q][col0 + lane] = __uint_as_float(mine);
      }
      tc_fence_before();
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
      if (gt < c3) {
        for (int s = 0; s < cpt; ++s) {
          float mx = sm.red[g][s * qpc][gt];
          for (int qq = 1; qq < qpc; ++qq) mx = fmaxf(mx, sm.red[g][s * qpc + qq][gt]);
          const long long ball = (rib + (long long)s * (128 / cpt)) / ns;
          float* dst = out + ((size_t)b * c3 + gt) * m + ball;
          if (ns <= 128) *dst = mx;
          else atomicMax(reinterpret_cast<int*>(dst), __float_as_int(mx));
        }
      }
      asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace upk

using namespace upk;

extern "C" int upk_shared_mlp_max_supported(int cin, int c1, int c2, int c3, int m, int ns) {
  if (cin < 1 || cin > 16) return 0;
  if (c1 < 16 || c1 > PM_C1 || c1 % 16) return 0;
  if (c2 < 32 || c2 > PM_C2 || c2 % 32) return 0;
  if (c3 < 32 || c3 > PM_C3 || c3 % 32) return 0;
  if (ns < 32 || !((ns % 128 == 0) || (128 % ns == 0))) return 0;
  if (m < 1 || ((long long)m * ns) % 128) return 0;
  return 1;
}

extern "C" int upk_shared_mlp_max(const float* x, int b, int cin, int m, int ns, int c1, int c2, int c3,
                                  const float* w1, const float* b1, const float* w2, const float* b2,
                                  const float* w3, const float* b3, float* out, upk_stream_t stream) {
  if (b < 0) return UPK_ERR_INVALID_ARG;
  if (!upk_shared_mlp_max_supported(cin, c1, c2, c3, m, ns)) return UPK_ERR_UNSUPPORTED;
  if (b == 0) return UPK_OK;
  if (!x || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !out) return UPK_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  if (ns > 128) UPK_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)b * c3 * m * sizeof(float), st));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long tiles = (long long)b * m * ns / 128;
  // UPK_PE_MLP_SLOTS (A/B measurements): 1 (default) = activations in tensor memory, layer 1 on the CUDA cores, two round
  // trips per tile (k_shared_mlp_max_ts2); 0 = activations in tensor memory, five round trips; 4 = four smem tile slots;
  // 2 = the round-1 two-slot kernel.  All four produce outputs within 1e-6 of each other.
  static int slots = -1;
  if (slots < 0) { const char* e = getenv("UPK_PE_MLP_SLOTS"); slots = e ? atoi(e) : 1; }
  const int grid = (int)(tiles < sms ? tiles : sms);
  if (slots == 1) {          // activations in tensor memory, layer 1 on the CUDA cores, two round trips per tile
    const size_t smem = sizeof(Pt2Smem) + 1024;
    if (cin > 8) {
      UPK_CUDA_TRY(cudaFuncSetAttribute(k_shared_mlp_max_ts2<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_shared_mlp_max_ts2<16><<<grid, PT_THREADS, smem, st>>>(x, cin, c1, c2, c3, m, ns, tiles, w1, b1, w2, b2, w3, b3, out);
    } else {
      UPK_CUDA_TRY(cudaFuncSetAttribute(k_shared_mlp_max_ts2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_shared_mlp_max_ts2<8><<<grid, PT_THREADS, smem, st>>>(x, cin, c1, c2, c3, m, ns, tiles, w1, b1, w2, b2, w3, b3, out);
    }
  } else if (slots == 0) {   // activations in tensor memory (TS MMAs), five round trips per tile
    const size_t smem = sizeof(PtSmem) + 1024;
    UPK_CUDA_TRY(cudaFuncSetAttribute(k_shared_mlp_max_ts, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_shared_mlp_max_ts<<<grid, PT_THREADS, smem, st>>>(x, cin, c1, c2, c3, m, ns, tiles, w1, b1, w2, b2, w3, b3, out);
  } else if (slots == 2) {
    const size_t smem = sizeof(PmSmem) + 1024;
    UPK_CUDA_TRY(cudaFuncSetAttribute(k_shared_mlp_max, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_shared_mlp_max<<<grid, PM_THREADS, smem, st>>>(x, cin, c1, c2, c3, m, ns, tiles, w1, b1, w2, b2, w3, b3, out);
  } else {
    const size_t smem = sizeof(Pm4Smem) + 1024;
    if (cin > 8) {
      UPK_CUDA_TRY(cudaFuncSetAttribute(k_shared_mlp_max4<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_shared_mlp_max4<true><<<grid, PM4_THREADS, smem, st>>>(x, cin, c1, c2, c3, m, ns, tiles, w1, b1, w2, b2, w3, b3, out);
    } else {
      UPK_CUDA_TRY(cudaFuncSetAttribute(k_shared_mlp_max4<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_shared_mlp_max4<false><<<grid, PM4_THREADS, smem, st>>>(x, cin, c1, c2, c3, m, ns, tiles, w1, b1, w2, b2, w3, b3, out);
    }
  }
  count_launch();
  UPK_RETURN_LAST_ERROR();
}
