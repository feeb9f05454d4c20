// (1a) Feature-similarity GEMM on the 5th-gen tensor cores: tcgen05.mma (kind::tf32) with TMEM
// accumulators, operands staged by TMA (cp.async.bulk.tensor, SWIZZLE_64B) — the fine-stage
// 2049 x 2049 x 256 NT GEMM of compute_feature_similarity (model_utils.py:260-282).
//
// Precision: the reference runs this GEMM in true fp32 (cuBLAS SGEMM, TF32 off,
// main_unopose.py:139-141) and the logits are divided by temp = 0.1, so a plain TF32 product
// (2^-11 per operand) is not enough for the R/t tolerance.  We use the 3xTF32 split
//     a = a_hi + a_lo,  a_hi = rna_tf32(a),  a_lo = a - a_hi   (exact in fp32)
//     a.b ~= a_hi.b_hi + a_hi.b_lo + a_lo.b_hi                 (dropped terms ~2^-22)
// accumulated in the fp32 TMEM accumulator: ~fp32 accuracy at 3 MMAs per product.
// The split (and the F.normalize) is done once per operand by k_normalize_split.
//
// Kernel anatomy (persistent, one CTA per SM, 192 threads):
//   warp 0      TMA producer : 4 boxes / K-chunk (A_hi, A_lo: 128x16 fp32; B_hi, B_lo: 256x16 fp32)
//   warp 1      MMA issuer   : 6 x tcgen05.mma M128 N256 K8 per K-chunk, tcgen05.commit -> mbarriers
//   warps 2..9  epilogue     : tcgen05.ld 32x32b.x32 -> * 1/temp -> smem transpose -> coalesced stores
// Pipelines: 3 smem stages (48 KB each) between TMA and MMA; 2 TMEM accumulators (2 x 256 columns)
// between MMA and epilogue, so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "launch_count.h"
#include "pose_internal.h"
#include "tc_ptx.cuh"
#include "../../include/unopose_b200.h"

#ifndef UPK_EPI_XCHG
#define UPK_EPI_XCHG 0   // see the epilogue of k_similarity_tc2
#endif

namespace upk {

constexpr int TC_STAGES = 3;  // 48 KB / stage
constexpr int TC_EPI_WARPS = 8;  // two warps per TMEM lane quarter, each takes half of the 256 columns
constexpr int TC_THREADS = 64 + TC_EPI_WARPS * 32;
constexpr uint32_t TC_STAGE_BYTES = (2 * TC_BM * TC_BK + 2 * TC_BN * TC_BK) * 4;  // 49152

struct __align__(1024) TcSmem {
  float a_hi[TC_STAGES][TC_BM * TC_BK];
  float a_lo[TC_STAGES][TC_BM * TC_BK];
  float b_hi[TC_STAGES][TC_BN * TC_BK];
  float b_lo[TC_STAGES][TC_BN * TC_BK];
  float epi[TC_EPI_WARPS][32][33];
  unsigned long long full[TC_STAGES], empty[TC_STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
};


// ---------------------------------------------------------------- operand preparation
// out_hi = rna_tf32(x / max(||x||, 1e-12)) ; out_lo = x_n - out_hi.   One warp per row.
// BORDER (the GEMM tiles start at (1,1)): the same warp also produces this row's entry of the peeled
// background row / column of the output, i.e. its dot product with row 0 of the OTHER operand
// (`other`, normalised on the fly with the same operations, so the values equal the split ones):
//   first operand  (is_a = 1): C[b][row][0]      second operand (is_a = 0): C[b][0][row]
constexpr int NS_RPW = 4;   // rows per warp -> 32 rows per CTA

// F16 (3xFP16 split, the default for normalised cosine logits on the CTA-pair kernel; UPK_SIMILARITY_MODE=3 restores
// 3xTF32): hi = fp16(x_n * 2^12), lo = fp16(x_n * 2^12 - hi), stored as halves (the outputs are then __half arrays of the
// same element count).  |x_n| <= 1, so the scaled operands stay far below fp16's maximum; the 2^12 scale keeps the
// residuals out of fp16's subnormal range; fp16 and tf32 both carry 11 significand bits, so the three-product sum has
// the accuracy of 3xTF32 (scripts/dev/split_precision.py; measured on B200: max |logit error| vs fp64 < 2e-5, same as
// 3xTF32) at twice the MMA rate and half the operand bytes.
constexpr float kF16Scale = 4096.0f;

// Both operands are prepared by ONE launch (blockIdx.z selects the operand): the two preparations are independent and
// each alone is a ~10 us launch whose tail and launch latency were serialised on the stream.
struct NsOperand {
  const float* x;       // (batch, rows, c)
  int rows;
  float* hi;
  float* lo;
  const float* other;   // BORDER: the other operand (its row 0 is the background token), or null
  int other_rows;
  int is_a;
};

template <int MODE, bool F16 = false, int RIF = 4>   // RIF: rows in flight per warp (4 for c <= 256, else 2)
__global__ void __launch_bounds__(256)
k_normalize_split(const NsOperand opa, const NsOperand opb, int c, int normalize,
                  float temp, float* __restrict__ C, int M, int N, int ldc) {
  extern __shared__ float s_q[];   // BORDER: the normalised row 0 of the other operand (c floats)
  const NsOperand& op = blockIdx.z ? opb : opa;
  if ((int)blockIdx.x * 8 * NS_RPW >= op.rows) return;   // the grid covers the longer operand
  const float* __restrict__ x = op.x;
  const int rows_per_batch = op.rows;
  float* __restrict__ hi = op.hi;
  float* __restrict__ lo = op.lo;
  const float* __restrict__ other = op.other;
  const int other_rows_per_batch = op.other_rows, is_a = op.is_a;
  const int bidx = blockIdx.y;
  const int lane = threadIdx.x & 31;
  if (other) {
    const float* q = other + (size_t)bidx * other_rows_per_batch * c;
    float qn = 1.f;
    if (normalize) {   // every warp computes the same norm (c loads from L1/L2), no extra barrier needed
      float s = 0.f;
      for (int k = lane; k < c; k += 32) {
        float v = q[k];
        s = fmaf(v, v, s);
      }
      s = warp_sum(s);
      qn = fmaxf(sqrtf(s), 1e-12f);
    }
    const float qinv = 1.0f / qn;
    // F16: the row values below carry the 2^12 scale, so the staged row carries 2^-12 (exact: the products are the same)
    for (int k = threadIdx.x; k < c; k += 256) s_q[k] = F16 ? (q[k] * qinv) * (1.0f / kF16Scale) : q[k] * qinv;
    __syncthreads();
  }
  // NS_RPW rows per warp: the border prologue above (norm + staging of the other operand's row 0) is per CTA.
  // All of a warp's rows are requested before the first one is reduced (c <= 256: 8 independent 128-bit loads per
  // lane in flight instead of 2 - the kernel was latency-bound at 2.9 TB/s); wider rows go two at a time.
  const int nv = c >> 2;   // 128-bit accesses (c % 4 == 0 is guaranteed by the tensor-core path, rows are 16-byte aligned)
  constexpr int rif = RIF;                        // rows in flight
  constexpr int kv = 8 / rif;                     // float4 kept per lane and row (covers c <= 256 / c <= 512)
  static_assert(NS_RPW % RIF == 0 && (RIF == 2 || RIF == 4), "");
  const float4* q4 = reinterpret_cast<const float4*>(s_q);
  const int r_first = (blockIdx.x * 8 + (threadIdx.x >> 5)) * NS_RPW;
#pragma unroll 1
  for (int rw0 = 0; rw0 < NS_RPW; rw0 += rif) {
    float4 keep[8];
#pragma unroll
    for (int j = 0; j < rif; ++j) {
      const int r = r_first + rw0 + j;
      const float4* p4 = reinterpret_cast<const float4*>(x + ((size_t)bidx * rows_per_batch + r) * c);
#pragma unroll
      for (int u = 0; u < kv; ++u) {
        const int k = lane + 32 * u;
        keep[j * kv + u] = (r < rows_per_batch && k < nv) ? __ldg(p4 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int j = 0; j < rif; ++j) {
      const int r = r_first + rw0 + j;
      if (r >= rows_per_batch) break;
      const size_t row = (size_t)bidx * rows_per_batch + r;
      const float4* p4 = reinterpret_cast<const float4*>(x + row * c);
      float s = 0.f;
#pragma unroll
      for (int u = 0; u < kv; ++u) {
        const float4 v = keep[j * kv + u];   // zero beyond the row: adds nothing
        s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
      }
      for (int k = lane + 32 * kv; k < nv; k += 32) {
        const float4 v = __ldg(p4 + k);
        s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
      }
      float inv = 1.f;
      if (normalize) {
        s = warp_sum(s);
        inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);   // x * (1 / ||x||): within 1 ulp of F.normalize's division
      }
      float dot = 0.f;
      float4* hi4 = reinterpret_cast<float4*>(hi + row * c);
      float4* lo4 = reinterpret_cast<float4*>(lo + row * c);
      auto split = [&](float4 v, int k) {
        v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
        if (F16) {
          // packed conversions (F2FP / HADD2.F32 on the FMA pipes; the scalar F2F.F16.F32 runs at 16 lanes per clock)
          v.x *= kF16Scale; v.y *= kF16Scale; v.z *= kF16Scale; v.w *= kF16Scale;
          const __half2 hxy = __floats2half2_rn(v.x, v.y), hzw = __floats2half2_rn(v.z, v.w);
          const float2 fxy = __half22float2(hxy), fzw = __half22float2(hzw);
          const __half2 lxy = __floats2half2_rn(v.x - fxy.x, v.y - fxy.y), lzw = __floats2half2_rn(v.z - fzw.x, v.w - fzw.y);
          uint2 ph, pl;
          ph.x = *reinterpret_cast<const uint32_t*>(&hxy); ph.y = *reinterpret_cast<const uint32_t*>(&hzw);
          pl.x = *reinterpret_cast<const uint32_t*>(&lxy); pl.y = *reinterpret_cast<const uint32_t*>(&lzw);
          reinterpret_cast<uint2*>(reinterpret_cast<__half*>(hi) + row * c)[k] = ph;
          reinterpret_cast<uint2*>(reinterpret_cast<__half*>(lo) + row * c)[k] = pl;
        } else {
          float4 h;
          uint32_t t;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t);
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t);
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t);
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t);
          hi4[k] = h;
          lo4[k] = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
        }
        if (other) {
          const float4 q = q4[k];
          dot = fmaf(v.x, q.x, fmaf(v.y, q.y, fmaf(v.z, q.z, fmaf(v.w, q.w, dot))));
        }
      };
#pragma unroll
      for (int u = 0; u < kv; ++u) {
        const int k = lane + 32 * u;
        if (k < nv) split(keep[j * kv + u], k);
      }
      for (int k = lane + 32 * kv; k < nv; k += 32) split(__ldg(p4 + k), k);
      if (other) {
        dot = warp_sum(dot);
        if (MODE == 1) dot = sqrtf(fmaxf(2.0f - 2.0f * dot, 0.f));
        if (lane == 0) {
          float* o = is_a ? C + ((size_t)bidx * M + r) * ldc : C + (size_t)bidx * M * ldc + r;
          if (is_a || r > 0) *o = dot * (1.0f / temp);   // C[0][0] is written once, by the first operand's row 0
        }
      }
    }
  }
}

// ---------------------------------------------------------------- the GEMM
// STATS (cosine logits only): the epilogue also produces pass 1 of the dual-softmax assignment, the sums of
// exponentials per row and per column.  |logit| <= 1/temp, so ONE fixed reference exponent gref = log2e/temp (+ a
// small margin) serves every row and column: e = 2^(v log2e - gref) <= 1, >= 2^(-2 log2e/temp).  Row sums are
// thread-local at TMEM-load time (a thread holds 32 columns of its row), column sums are thread-local at store time
// (a lane then walks the 32 rows of its column through the transpose buffer); no shuffles, no atomics:
//   rowpart[(b * 4 nt + 4 ni + column slice) * M + row]   colpart[(b * 4 mt + 4 mi + q) * N + col]   (partial-major: coalesced)
// LIN: the same GEMM as a fully-connected layer, C = A B^T + bias (per column), optional ReLU (upk_linear): the bias
// and the activation are applied at store time, where a lane owns one output column.
template <int MODE, int NTERMS, bool STATS = false, bool LIN = false>  // MODE 0: dot/temp, 1: sqrt(clamp(2-2dot,0))/temp ; NTERMS 3 = 3xTF32, 1 = TF32
__global__ void __launch_bounds__(TC_THREADS, 1)
k_similarity_tc(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                int batch, int M, int N, int K, float temp, int off, float* __restrict__ C,
                float* __restrict__ rowpart, float* __restrict__ colpart, float gref, int ldc,
                const float* __restrict__ lin_bias, int lin_relu) {
  // `off` (0 or 1): the tiles cover rows/columns [off, M) x [off, N); with off = 1 the background row 0 and
  // column 0 are produced by k_normalize_split (BORDER), so the 2049 x 2049 fine shape is exactly 16 x 8 tiles
  extern __shared__ unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = (M - off + TC_BM - 1) / TC_BM, nt = (N - off + TC_BN - 1) / TC_BN;
  const int mts = 2 * ((M - off + 2 * TC_BM - 1) / (2 * TC_BM));   // statistics layout: 128-row tiles rounded to pairs
  const int total = batch * mt * nt;
  const int kchunks = K / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tfull[a], 1); mbar_init(&sm.tempty[a], TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===== TMA producer =====
    int s = 0;
    uint32_t ph = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x) {
      const int b = t / (mt * nt), rem = t - b * mt * nt;
      const int mi = rem / nt, ni = rem - mi * nt;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&sm.empty[s], ph ^ 1);
        if (lane == 0) {
          mbar_expect_tx(&sm.full[s], NTERMS == 3 ? TC_STAGE_BYTES : TC_STAGE_BYTES / 2);
          tma_load_3d(&map_a_hi, &sm.full[s], sm.a_hi[s], kc * TC_BK, off + mi * TC_BM, b);
          tma_load_3d(&map_b_hi, &sm.full[s], sm.b_hi[s], kc * TC_BK, off + ni * TC_BN, b);
          if (NTERMS == 3) {
            tma_load_3d(&map_a_lo, &sm.full[s], sm.a_lo[s], kc * TC_BK, off + mi * TC_BM, b);
            tma_load_3d(&map_b_lo, &sm.full[s], sm.b_lo[s], kc * TC_BK, off + ni * TC_BN, b);
          }
        }
        __syncwarp();
        if (++s == TC_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&sm.tempty[acc], acc_ph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * TC_BN;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&sm.full[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t ahi = make_desc_sw128(sm.a_hi[s]), alo = make_desc_sw128(sm.a_lo[s]);
          const uint64_t bhi = make_desc_sw128(sm.b_hi[s]), blo = make_desc_sw128(sm.b_lo[s]);
#pragma unroll
          for (int kk = 0; kk < TC_BK / 8; ++kk) {
            const uint64_t adv = (uint64_t)(kk * 8 * 4 >> 4);  // 32 B along K inside the swizzled row
            if (NTERMS == 3) {
              tc_mma_tf32(d_tmem, alo + adv, bhi + adv, kIdescTf32, (kc | kk) ? 1u : 0u);
              tc_mma_tf32(d_tmem, ahi + adv, blo + adv, kIdescTf32, 1u);
              tc_mma_tf32(d_tmem, ahi + adv, bhi + adv, kIdescTf32, 1u);
            } else {
              tc_mma_tf32(d_tmem, ahi + adv, bhi + adv, kIdescTf32, (kc | kk) ? 1u : 0u);
            }
          }
          tc_commit(&sm.empty[s]);                       // smem stage reusable once these MMAs retire
          if (kc == kchunks - 1) tc_commit(&sm.tfull[acc]);  // accumulator complete
        }
        __syncwarp();
        if (++s == TC_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps (TMEM lane quarter = warp % 4; warps 2..5 take columns [0,128),
    //       warps 6..9 columns [128,256)).  A single warp issues roughly one dependent instruction
    //       every ~6 cycles, so the epilogue is kept short: multiply by 1/temp, pointer-increment
    //       addressing, LDS/STS through a padded 32x33 transpose so every store is a 128-byte row
    //       segment. =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    float* tr = &sm.epi[warp - 2][0][0];
    const float inv_temp = 1.0f / temp;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      const int b = t / (mt * nt), rem = t - b * mt * nt;
      const int mi = rem / nt, ni = rem - mi * nt;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&sm.tfull[acc], acc_ph);
      tc_fence_after();
      const int row0 = off + mi * TC_BM + q * 32;
      const int nrows = min(32, M - row0);
      float* Cb = C + ((size_t)b * M + row0) * ldc;
      float rsum = 0.f;   // STATS: this thread's row (row0 + lane), its 128 columns of the tile
#pragma unroll 1
      for (int cb = half * 4; cb < half * 4 + 4; ++cb) {
        const int col0 = off + ni * TC_BN + cb * 32;
        if (col0 >= N) break;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + acc * TC_BN + cb * 32 + ((uint32_t)(q * 32) << 16), r);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float v = __uint_as_float(r[j]);
          if (MODE == 1) v = sqrtf(fmaxf(2.0f - 2.0f * v, 0.f));
          v *= inv_temp;
          tr[lane * 33 + j] = v;
          if (STATS) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v, 1.4426950408889634f, -gref)));
            rsum += (col0 + j < N) ? e : 0.f;
          }
        }
        __syncwarp();
        if (col0 + lane < N && nrows > 0) {
          float* dst = Cb + col0 + lane;
          const float* src = tr + lane;
          float csum = 0.f;   // STATS: this lane's column (col0 + lane), the 32 rows of this warp
          if (LIN) {
            const float bl = lin_bias ? __ldg(lin_bias + col0 + lane) : 0.f;
            const float lo = lin_relu ? 0.f : -INFINITY;
            for (int rr = 0; rr < nrows; ++rr) dst[(size_t)rr * ldc] = fmaxf(src[rr * 33] + bl, lo);
          } else if (nrows == 32) {
#pragma unroll
            for (int rr = 0; rr < 32; ++rr) {
              const float v = src[rr * 33];
              dst[(size_t)rr * ldc] = v;
              if (STATS) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v, 1.4426950408889634f, -gref)));
                csum += e;
              }
            }
          } else {
            for (int rr = 0; rr < nrows; ++rr) {
              const float v = src[rr * 33];
              dst[(size_t)rr * ldc] = v;
              if (STATS) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v, 1.4426950408889634f, -gref)));
                csum += e;
              }
            }
          }
          if (STATS) colpart[((size_t)b * (4 * mts) + 4 * mi + q) * N + col0 + lane] = csum;   // lanes: consecutive columns
        }
        __syncwarp();
      }
      if (STATS && lane < nrows) rowpart[((size_t)b * (4 * nt) + 4 * ni + 2 * half) * M + row0 + lane] = rsum;  // consecutive rows
      tc_fence_before();
      if (lane == 0) mbar_arrive(&sm.tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------- the GEMM, CTA-pair version
// tcgen05.mma.cta_group::2: the two CTAs of a cluster (one TPC) compute ONE 256 x 256 output tile.  Each CTA stages
// its 128 rows of A and HALF of the B tile (128 of the 256 B rows); the MMA issued by the leader CTA reads both
// halves from both shared memories and writes 128 accumulator rows into each CTA's TMEM.  Per K-chunk an SM now
// pulls 32 KB of operands (A 16 KB + B/2 16 KB, hi + lo) instead of 48 KB for the same 128 x 256 outputs: the
// single-CTA kernel is limited by SM <-> L2 traffic once its epilogue stores are added (profiles/, 62 % tensor pipe).
// Protocol (CUTLASS sm100 2-SM pipelines): only the leader's `full` barriers are used, its producer posts
// expect_tx for BOTH CTAs' bytes and both producers' TMA loads (.cta_group::2) complete on it; the leader's MMA
// warp releases smem stages and publishes accumulators with multicast commits to both CTAs; both CTAs' epilogue
// warps arrive on the leader's `tempty`.  cosine logits, 3xTF32 only.
constexpr int TC2_EPI_WARPS = 8;                   // EPI/4 warps per TMEM lane quarter, 256/(EPI/4) columns each (measured: 4 -> 166 us, 8 -> 146 us, 16 -> 160 us)
constexpr int TC2_CPW = 32 / TC2_EPI_WARPS;        // 32-column chunks per epilogue warp and tile
constexpr int TC2_THREADS = 64 + TC2_EPI_WARPS * 32;
constexpr uint32_t TC2_STAGE_BYTES = 4 * TC_BM * TC_BK * 4;   // A_hi, A_lo, Bhalf_hi, Bhalf_lo: 4 x 8 KB
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;     // shared::cluster address of the same offset in CTA 0 of the pair

// TMAST = false: 5 operand stages of 32 KB + a padded 32x33 transpose buffer per epilogue warp (4-byte row stores).
// TMAST = true : 4 operand stages + two 4 KB TMA-store staging boxes per epilogue warp (32 rows x 128 B, SWIZZLE_128B).
template <bool TMAST>
struct __align__(1024) Tc2Smem {
  static constexpr int STAGES = TMAST ? 4 : 5;
  static constexpr int EPI_FLOATS = TMAST ? 2 * 32 * 32 : 32 * 33;
  float a_hi[STAGES][TC_BM * TC_BK];
  float a_lo[STAGES][TC_BM * TC_BK];
  float b_hi[STAGES][TC_BM * TC_BK];   // this CTA's half (128 rows) of the 256-row B tile
  float b_lo[STAGES][TC_BM * TC_BK];
  float epi[TC2_EPI_WARPS][EPI_FLOATS];   // TMAST: 8 KB per warp, every 4 KB box 1024-byte aligned
  unsigned long long full[STAGES], empty[STAGES], tfull[2], tempty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"((unsigned long long)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d_hint(const CUtensorMap* map, const void* src, int c0, int c1, int c2, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
               ::"l"((unsigned long long)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* map, void* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((unsigned long long)map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(void* bar) {   // arrives on `bar` (same offset) in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(z) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(void* bar) {   // arrive on CTA 0's barrier at this offset
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// instruction descriptor of the pair: D=F32, A=B=TF32, K-major, N=256, M=256
constexpr uint32_t kIdescTf32_2sm = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
                                    ((uint32_t)((2 * TC_BM) >> 4) << 24);

__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc), "r"(z) : "memory");
}
constexpr uint32_t kIdescF16_2sm = kIdescF16Base | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)((2 * TC_BM) >> 4) << 24);

// F16: the operands are fp16 (k_normalize_split<.., true>): a 64-byte stage row holds 32 K elements instead of 16, one
// tcgen05.mma.kind::f16 covers K = 16; the accumulator carries the 2^24 scale of the two operands.
// TMAST: the output is pitched so that element (1, 1) of every instance is 16-byte aligned and ldc % 4 == 0 (the padded
// layout compute_feature_similarity hands out for the fine shape): an epilogue warp then writes its 32 x 32 chunk into a
// SWIZZLE_128B staging box (8 conflict-free STS.128 per lane) and ONE elected lane issues a TMA tensor store
// (cp.async.bulk.tensor, clipped at the ragged edges by the hardware) instead of 32 row-segment store instructions per
// lane — the 4-byte-aligned row stores were what held the kernel below the MMA rate (profiles/r1_latency_microbench.txt).
template <bool STATS, bool F16 = false, bool TMAST = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
k_similarity_tc2(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const __grid_constant__ CUtensorMap map_c,
                 int batch, int M, int N, int K, float temp, int off, float* __restrict__ C, int ldc,
                 float* __restrict__ rowpart, float* __restrict__ colpart, float gref, int l2_keep) {
  extern __shared__ unsigned char smem_raw[];
  typedef Tc2Smem<TMAST> Smem;
  constexpr int TC2_STAGES = Smem::STAGES;
  Smem& sm = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const bool leader = rank == 0;
  const int mt2 = (M - off + 2 * TC_BM - 1) / (2 * TC_BM);    // 256-row tiles of the pair
  const int mt = 2 * mt2;                                     // 128-row tiles (statistics layout)
  const int nt = (N - off + TC_BN - 1) / TC_BN;
  const int total = batch * mt2 * nt;
  constexpr int KE = F16 ? 2 * TC_BK : TC_BK;   // K elements per 64-byte stage row
  const int kchunks = K / KE;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC2_STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&sm.tfull[a], 1); mbar_init(&sm.tempty[a], 2 * TC2_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // the same warp of both CTAs, same destination address
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = sm.tmem_base;

  if (warp == 0) {
    // ===== TMA producer (both CTAs) =====
    int s = 0;
    uint32_t ph = 0;
    for (int t = cid; t < total; t += ncl) {
      const int b = t / (mt2 * nt), rem = t - b * mt2 * nt;
      const int mi2 = rem / nt, ni = rem - mi2 * nt;
      const int arow = off + mi2 * 2 * TC_BM + (int)rank * TC_BM;
      const int brow = off + ni * TC_BN + (int)rank * TC_BM;
      for (int kc = 0; kc < kchunks; ++kc) {
        mbar_wait(&sm.empty[s], ph ^ 1);
        if (lane == 0) {
          if (leader) mbar_expect_tx(&sm.full[s], 2 * TC2_STAGE_BYTES);   // both CTAs' loads land on this barrier
          tma_load_3d_2sm(&map_a_hi, &sm.full[s], sm.a_hi[s], kc * KE, arow, b);
          tma_load_3d_2sm(&map_b_hi, &sm.full[s], sm.b_hi[s], kc * KE, brow, b);
          tma_load_3d_2sm(&map_a_lo, &sm.full[s], sm.a_lo[s], kc * KE, arow, b);
          tma_load_3d_2sm(&map_b_lo, &sm.full[s], sm.b_lo[s], kc * KE, brow, b);
        }
        __syncwarp();
        if (++s == TC2_STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (leader CTA only) =====
    if (leader) {
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = cid; t < total; t += ncl, ++it) {
        const int acc = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&sm.tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TC_BN;
        for (int kc = 0; kc < kchunks; ++kc) {
          mbar_wait(&sm.full[s], ph);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t ahi = make_desc_sw128(sm.a_hi[s]), alo = make_desc_sw128(sm.a_lo[s]);
            const uint64_t bhi = make_desc_sw128(sm.b_hi[s]), blo = make_desc_sw128(sm.b_lo[s]);
#pragma unroll
            for (int kk = 0; kk < TC_BK / 8; ++kk) {
              const uint64_t adv = (uint64_t)(kk * 8 * 4 >> 4);
              if (F16) {
                tc_mma_f16_2sm(d_tmem, alo + adv, bhi + adv, kIdescF16_2sm, (kc | kk) ? 1u : 0u);
                tc_mma_f16_2sm(d_tmem, ahi + adv, blo + adv, kIdescF16_2sm, 1u);
                tc_mma_f16_2sm(d_tmem, ahi + adv, bhi + adv, kIdescF16_2sm, 1u);
              } else {
              tc_mma_tf32_2sm(d_tmem, alo + adv, bhi + adv, kIdescTf32_2sm, (kc | kk) ? 1u : 0u);
              tc_mma_tf32_2sm(d_tmem, ahi + adv, blo + adv, kIdescTf32_2sm, 1u);
              tc_mma_tf32_2sm(d_tmem, ahi + adv, bhi + adv, kIdescTf32_2sm, 1u);
              }
            }
            tc_commit_2sm(&sm.empty[s]);                          // stage reusable in both CTAs
            if (kc == kchunks - 1) tc_commit_2sm(&sm.tfull[acc]);   // accumulators complete in both CTAs
          }
          __syncwarp();
          if (++s == TC2_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else {
    // ===== epilogue warps (both CTAs; this CTA's 128 rows of the 256-row tile) =====
    const int q = warp & 3;
    const int cq = (warp - 2) >> 2;          // column slice of the tile handled by this warp
    float* tr = &sm.epi[warp - 2][0];
    const float inv_temp = F16 ? 1.0f / (temp * kF16Scale * kF16Scale) : 1.0f / temp;
    const uint64_t pol_last = l2_policy_evict_last(), pol_first = l2_policy_evict_first();
    int it = 0;
    int cc = 0;   // TMAST: chunks stored so far by this warp (staging box = cc & 1)
    for (int t = cid; t < total; t += ncl, ++it) {
      const int b = t / (mt2 * nt), rem = t - b * mt2 * nt;
      const int mi2 = rem / nt, ni = rem - mi2 * nt;
      const int mi = 2 * mi2 + (int)rank;
      const int acc = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      mbar_wait(&sm.tfull[acc], acc_ph);
      tc_fence_after();
      const int row0 = off + mi * TC_BM + q * 32;
      const int nrows = min(32, M - row0);
      float* Cb = C + ((size_t)b * M + row0) * ldc;
      float rsum = 0.f;
#pragma unroll 1
      for (int cb = cq * TC2_CPW; cb < cq * TC2_CPW + TC2_CPW; ++cb) {
        const int col0 = off + ni * TC_BN + cb * 32;
        if (col0 >= N) break;
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + acc * TC_BN + cb * 32 + ((uint32_t)(q * 32) << 16), r);
        if (TMAST) {
          float* box = tr + ((cc & 1) << 10);
          if (lane == 0) bulk_wait_read<1>();   // the store issued from this box two chunks ago has read it
          __syncwarp();
          float ex[(STATS && UPK_EPI_XCHG) ? 32 : 1];   // XCHG: this row's exponentials of the chunk
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            float4 v;
            v.x = __uint_as_float(r[4 * c4]) * inv_temp;
            v.y = __uint_as_float(r[4 * c4 + 1]) * inv_temp;
            v.z = __uint_as_float(r[4 * c4 + 2]) * inv_temp;
            v.w = __uint_as_float(r[4 * c4 + 3]) * inv_temp;
            if (STATS) {
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(vv[e4], 1.4426950408889634f, -gref)));
                if (UPK_EPI_XCHG) ex[4 * c4 + e4] = e;
                else rsum += (col0 + 4 * c4 + e4 < N) ? e : 0.f;
              }
            }
            // row `lane` of the box (128 B), 16-byte chunk c4 at position c4 ^ (row & 7): the SWIZZLE_128B pattern
            *reinterpret_cast<float4*>(box + lane * 32 + ((c4 ^ (lane & 7)) << 2)) = v;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // l2_keep: the last l2_keep instances are what the assignment pass that follows reads FIRST (it walks the
            // instances in descending order): keep them in L2, stream the rest through
            if (l2_keep > 0) tma_store_3d_hint(&map_c, box, col0 - off, row0 - off, b, b >= batch - l2_keep ? pol_last : pol_first);
            else tma_store_3d(&map_c, box, col0 - off, row0 - off, b);
            bulk_commit();
          }
          if (STATS && UPK_EPI_XCHG) {
            // Experiment (compile with -DUPK_EPI_XCHG=1): column sums by halving exchanges between the lanes instead of
            // a second pass over the staging box.  Fewer issue slots and half the MUFU work, but measured SLOWER
            // (fine similarity 136 -> 145 us).  So was a variant of the column pass below with shared-window LDS / STS
            // (the compiler emits generic LD.E / ST.E.128 for the staging box) and no bound checks on whole chunks:
            // 25 % fewer issue slots, 136 -> 140 us.  The epilogue's instruction count is not what bounds the kernel;
            // ncu: SM -> L2 request port busy 64 % of the cycles, 302 MB of TMA stores + 4.9 TB/s of operand reads.
            if (col0 + 32 > N || nrows < 32) {   // ragged chunk (warp-uniform; never for the 2048 x 2048 main block)
#pragma unroll
              for (int j = 0; j < 32; ++j) ex[j] = (col0 + j < N && lane < nrows) ? ex[j] : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) rsum += ex[j];
            int o = 16;
#pragma unroll
            for (int n = 32; n > 1; n >>= 1, o >>= 1) {
              const bool upper = (lane & o) != 0;
#pragma unroll
              for (int i = 0; i < n / 2; ++i) {
                const float send = upper ? ex[i] : ex[i + n / 2];
                const float keep = upper ? ex[i + n / 2] : ex[i];
                ex[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
              }
            }
            if (col0 + lane < N && nrows > 0) colpart[((size_t)b * (4 * mt) + 4 * mi + q) * N + col0 + lane] = ex[0];
          } else if (STATS && col0 + lane < N && nrows > 0) {
            // column col0 + lane of the 32 rows: word (lane & 3) of chunk (lane >> 2) ^ (rr & 7) of row rr
            float csum = 0.f;
            for (int rr = 0; rr < nrows; ++rr) {
              const float v = box[rr * 32 + ((((lane >> 2) ^ (rr & 7)) << 2) | (lane & 3))];
              float e;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v, 1.4426950408889634f, -gref)));
              csum += e;
            }
            colpart[((size_t)b * (4 * mt) + 4 * mi + q) * N + col0 + lane] = csum;
          }
          ++cc;
        } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float v = __uint_as_float(r[j]) * inv_temp;
          tr[lane * 33 + j] = v;
          if (STATS) {
            float e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v, 1.4426950408889634f, -gref)));
            rsum += (col0 + j < N) ? e : 0.f;
          }
        }
        __syncwarp();
        if (col0 + lane < N && nrows > 0) {
          float* dst = Cb + col0 + lane;
          const float* src = tr + lane;
          float csum = 0.f;
          for (int rr = 0; rr < nrows; ++rr) {
            const float v = src[rr * 33];
            dst[(size_t)rr * ldc] = v;
            if (STATS) {
              float e;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(v, 1.4426950408889634f, -gref)));
              csum += e;
            }
          }
          if (STATS) colpart[((size_t)b * (4 * mt) + 4 * mi + q) * N + col0 + lane] = csum;
        }
        __syncwarp();
        }
      }
      if (STATS && lane < nrows) rowpart[((size_t)b * (4 * nt) + 4 * ni + cq * (16 / TC2_EPI_WARPS)) * M + row0 + lane] = rsum;
      tc_fence_before();
      if (lane == 0) mbar_arrive_leader(&sm.tempty[acc]);
    }
    if (TMAST && lane == 0) bulk_wait_all();   // every tensor store of this warp has completed
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody tears down while the peer may still signal or read
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  }
  return fn;
}

// fp16 operands (experimental 3xFP16 mode): same 64-byte rows, 32 halves wide
static int make_map_f16(CUtensorMap* map, const void* base, int batch, int rows, int K, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return UPK_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)K * 2, (cuuint64_t)rows * K * 2};
  cuuint32_t box[3] = {(cuuint32_t)(2 * TC_BK), (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? UPK_OK : UPK_ERR_INVALID_ARG;
}

static int make_map(CUtensorMap* map, const float* base, int batch, int rows, int K, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return UPK_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)K * 4, (cuuint64_t)rows * K * 4};
  cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? UPK_OK : UPK_ERR_INVALID_ARG;
}

// fp32 output tiles for the TMA-store epilogue: the main block [1, n) x [1, m) of every instance, row pitch ldc floats,
// boxes of 32 x 32 with SWIZZLE_128B (128-byte box rows).  `c11` = address of element (1, 1) of instance 0.
static int make_map_out(CUtensorMap* map, float* c11, int batch, int n, int m, int ldc) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return UPK_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)(m - 1), (cuuint64_t)(n - 1), (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ldc * 4, (cuuint64_t)n * ldc * 4};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)c11, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? UPK_OK : UPK_ERR_INVALID_ARG;
}

// How many of the LAST instances of a statistics-fused similarity launch stay in L2 for the assignment pass that
// follows (UPK_FINE_L2_KEEP_MB, default 50 MB of the 126 MB L2 — measured at B = 16: fine solve 149.9 us without the
// hints, 138.4 / 138.7 / 139.6 / 141.0 / 144.1 / 144.2 us at 40 / 50 / 60 / 72 / 88 / 110 MB; 0 disables them).
int fine_l2_keep(int b, int n, int m) {
  static int mb = -1;
  if (mb < 0) { const char* e = getenv("UPK_FINE_L2_KEEP_MB"); mb = e ? atoi(e) : 50; }
  if (mb <= 0) return 0;
  const double per = (double)n * m * 4.0 / 1048576.0;
  int k = (int)(mb / per);
  return k < b ? k : 0;      // everything fits anyway: no hints
}

static bool tma_store_ok(const float* out, int n, int m, int ldc) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("UPK_TC_TMA_STORE"); enabled = e ? atoi(e) : 1; }
  return enabled && n > 1 && m > 1 && ldc % 4 == 0 && ((reinterpret_cast<uintptr_t>(out + ldc + 1)) & 15) == 0;
}

// the same encoder for the other tensor-core translation units (geoembed.cu); `map` is a CUtensorMap*
int tc_make_map(void* map, const float* base, int batch, int rows, int K, int box_rows) {
  return make_map((CUtensorMap*)map, base, batch, rows, K, box_rows);
}

int tc_make_map_f16(void* map, const void* base, int batch, int rows, int K, int box_rows) {
  return make_map_f16((CUtensorMap*)map, base, batch, rows, K, box_rows);
}

// whether the driver hands out cuTensorMapEncodeTiled (callers keep their non-TMA kernels otherwise)
bool tc_tma_available() { return get_encode() != nullptr; }

// Plain (unswizzled) fp32 boxes of box_cols x box_rows out of a pitched (batch, rows, cols) matrix: the streaming
// passes of assign_fine.cu read `atten` through it.  `base` = element (0, 0) of the block the boxes tile, 16-byte aligned.
int tc_make_map_plane(void* map, const float* base, int batch, int rows, int cols, int ld, size_t batch_stride,
                      int box_cols, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) return UPK_ERR_UNSUPPORTED;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)batch_stride * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc((CUtensorMap*)map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? UPK_OK : UPK_ERR_INVALID_ARG;
}

static int g_sim_mode = -1;  // 16 = 3xFP16 on the CTA-pair shapes with normalised operands, 3xTF32 elsewhere (default); 3 = 3xTF32; 1 = 1xTF32; 0 = fp32 SIMT

int similarity_mode() {
  if (g_sim_mode < 0) {
    const char* e = getenv("UPK_SIMILARITY_MODE");
    g_sim_mode = e ? atoi(e) : 16;
    if (g_sim_mode != 0 && g_sim_mode != 1 && g_sim_mode != 3 && g_sim_mode != 16) g_sim_mode = 16;
  }
  return g_sim_mode;
}

size_t similarity_tc_workspace_bytes(int b, int n, int m, int c) {
  size_t a = (((size_t)b * n * c * sizeof(float)) + 1023) & ~(size_t)1023;
  size_t bb = (((size_t)b * m * c * sizeof(float)) + 1023) & ~(size_t)1023;
  return 2 * a + 2 * bb + 1024;
}

SimStatsGeom sim_stats_geom(int b, int n, int m) {
  SimStatsGeom g;
  g.npr = 4 * ((m - 1 + TC_BN - 1) / TC_BN);   // up to 4 column slices per 256-column tile (the CTA-pair kernel's epilogue)
  g.npc = 4 * 2 * ((n - 1 + 2 * TC_BM - 1) / (2 * TC_BM));   // 4 per 128-row tile, tiles rounded up to CTA pairs
  g.row_floats = (size_t)b * n * g.npr;
  g.col_off_floats = (g.row_floats + 63) & ~(size_t)63;
  g.total_bytes = (g.col_off_floats + (size_t)b * m * g.npc) * sizeof(float);
  return g;
}

// reference exponent (log2 units) of the fused statistics: |cosine| <= 1 up to rounding, so v log2e <= gref
float sim_stats_gref(float temp) { return 1.4426950408889634f / temp + 0.01f; }

bool similarity_tc_eligible(int n, int m, int c) {
  return similarity_mode() != 0 && c % TC_BK == 0 && c >= TC_BK && (long long)n * m >= 128LL * 128LL &&
         get_encode() != nullptr;
}

// 3xFP16 split on the CTA-pair kernel (mode 16, the default; validated on B200 in round 2: all fine-stage parity suites
// green, fine similarity 201 -> 158 us at B = 16).  Only for NORMALISED operands: the 2^12 operand scale would overflow
// fp16 for |x| >= 16.  Returns 1 if the call is not handled here (the caller continues with the 3xTF32 path), else a
// launch status.
static int run_similarity_f16(const float* f1, const float* f2, int b, int n, int m, int c, float temp, int normalize,
                              int sim_type, void* a_hi, void* a_lo, void* b_hi, void* b_lo, float* out, int ldc,
                              cudaStream_t st, float* stats_row, float* stats_col, float stats_gref) {
  const char* e = getenv("UPK_TC_2SM");
  if (e && atoi(e) == 0) return 1;
  if (sim_type != 0 || !normalize || n <= 1 || m <= 1 || c % (2 * TC_BK) != 0) return 1;
  const int off = 1;   // background row / column peeled, as on the CTA-pair 3xTF32 path at these shapes
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles2 = b * ((n - off + 2 * TC_BM - 1) / (2 * TC_BM)) * ((m - off + TC_BN - 1) / TC_BN);
  if (tiles2 < sms / 2) return 1;
  if (stats_row && !stats_col) return UPK_ERR_UNSUPPORTED;
  const dim3 g12(((n > m ? n : m) + 8 * NS_RPW - 1) / (8 * NS_RPW), b, 2);
  const size_t qs = (size_t)c * sizeof(float);
  const NsOperand oa{f1, n, (float*)a_hi, (float*)a_lo, f2, m, 1}, ob{f2, m, (float*)b_hi, (float*)b_lo, f1, n, 0};
  if (c <= 256) k_normalize_split<0, true, 4><<<g12, 256, qs, st>>>(oa, ob, c, normalize, temp, out, n, m, ldc);
  else k_normalize_split<0, true, 2><<<g12, 256, qs, st>>>(oa, ob, c, normalize, temp, out, n, m, ldc);
  count_launch(1);
  CUtensorMap fa_hi, fa_lo, fb_hi, fb_lo;
  int rc;
  if ((rc = make_map_f16(&fa_hi, a_hi, b, n, c, TC_BM))) return rc;
  if ((rc = make_map_f16(&fa_lo, a_lo, b, n, c, TC_BM))) return rc;
  if ((rc = make_map_f16(&fb_hi, b_hi, b, m, c, TC_BM))) return rc;
  if ((rc = make_map_f16(&fb_lo, b_lo, b, m, c, TC_BM))) return rc;
  const int grid2 = 2 * (tiles2 < sms / 2 ? tiles2 : sms / 2);
  CUtensorMap mc;
  const bool tmast = tma_store_ok(out, n, m, ldc);
  if (tmast) {
    if ((rc = make_map_out(&mc, out + ldc + 1, b, n, m, ldc))) return rc;
  } else {
    mc = fa_hi;   // unused by the kernel
  }
#define UPK_LAUNCH_TC2(ST, TM)                                                                                       \
  do {                                                                                                               \
    auto kern = k_similarity_tc2<ST, true, TM>;                                                                      \
    const size_t smem2 = sizeof(Tc2Smem<TM>) + 1024;                                                                 \
    UPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));              \
    kern<<<grid2, TC2_THREADS, smem2, st>>>(fa_hi, fa_lo, fb_hi, fb_lo, mc, b, n, m, c, temp, off, out, ldc,         \
                                            ST ? stats_row : nullptr, ST ? stats_col : nullptr, ST ? stats_gref : 0.f, ST ? fine_l2_keep(b, n, m) : 0); \
  } while (0)
  if (stats_row) { if (tmast) UPK_LAUNCH_TC2(true, true); else UPK_LAUNCH_TC2(true, false); }
  else { if (tmast) UPK_LAUNCH_TC2(false, true); else UPK_LAUNCH_TC2(false, false); }
#undef UPK_LAUNCH_TC2
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

// stats_row / stats_col (optional, both or neither; requires sim_type 0 and that the tiles start at (1,1)):
// per-tile partial sums of 2^(v log2e - stats_gref), layouts of SimStatsGeom.
int run_similarity_tc(const float* f1, const float* f2, int b, int n, int m, int c, float temp, int normalize,
                      int sim_type, void* workspace, size_t workspace_bytes, float* out, cudaStream_t st,
                      float* stats_row, float* stats_col, float stats_gref, int ldc) {
  if (ldc <= 0) ldc = m;
  if (ldc < m) return UPK_ERR_INVALID_ARG;
  if (workspace_bytes < similarity_tc_workspace_bytes(b, n, m, c)) return UPK_ERR_INVALID_ARG;
  char* w = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  size_t a = (((size_t)b * n * c * sizeof(float)) + 1023) & ~(size_t)1023;
  size_t bb = (((size_t)b * m * c * sizeof(float)) + 1023) & ~(size_t)1023;
  float* a_hi = (float*)w;
  float* a_lo = (float*)(w + a);
  float* b_hi = (float*)(w + 2 * a);
  float* b_lo = (float*)(w + 2 * a + bb);
  if (similarity_mode() == 16) {   // see run_similarity_f16; 1 = not applicable, fall through to 3xTF32
    const int rc16 = run_similarity_f16(f1, f2, b, n, m, c, temp, normalize, sim_type, a_hi, a_lo, b_hi, b_lo, out, ldc, st,
                                        stats_row, stats_col, stats_gref);
    if (rc16 != 1) return rc16;
  }
  // peel the first row and column off when that saves tiles (2049 = 2048 + the background token); the
  // peeled entries are produced by the operand-preparation kernels
  const int tiles0 = ((n + TC_BM - 1) / TC_BM) * ((m + TC_BN - 1) / TC_BN);
  const int tiles1 = ((n - 1 + TC_BM - 1) / TC_BM) * ((m - 1 + TC_BN - 1) / TC_BN);
  const int off = (n > 1 && m > 1 && (tiles1 < tiles0 || stats_row)) ? 1 : 0;   // the statistics assume the peel
  if (stats_row && (!off || sim_type != 0 || !stats_col)) return UPK_ERR_UNSUPPORTED;
  const dim3 g12(((n > m ? n : m) + 8 * NS_RPW - 1) / (8 * NS_RPW), b, 2);
  const size_t qs = off ? (size_t)c * sizeof(float) : 0;
  const NsOperand oa{f1, n, a_hi, a_lo, off ? f2 : nullptr, m, 1}, ob{f2, m, b_hi, b_lo, off ? f1 : nullptr, n, 0};
  if (sim_type == 0) {
    if (c <= 256) k_normalize_split<0, false, 4><<<g12, 256, qs, st>>>(oa, ob, c, normalize, temp, out, n, m, ldc);
    else k_normalize_split<0, false, 2><<<g12, 256, qs, st>>>(oa, ob, c, normalize, temp, out, n, m, ldc);
  } else {
    k_normalize_split<1, false, 2><<<g12, 256, qs, st>>>(oa, ob, c, normalize, temp, out, n, m, ldc);
  }
  count_launch(1);
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, a_hi, b, n, c, TC_BM))) return rc;
  if ((rc = make_map(&ma_lo, a_lo, b, n, c, TC_BM))) return rc;
  if ((rc = make_map(&mb_hi, b_hi, b, m, c, TC_BN))) return rc;
  if ((rc = make_map(&mb_lo, b_lo, b, m, c, TC_BN))) return rc;
  // CTA-pair kernel (cosine, 3xTF32, enough tiles to fill the clusters; UPK_TC_2SM=0 disables it)
  static int use_2sm = -1;
  if (use_2sm < 0) { const char* e = getenv("UPK_TC_2SM"); use_2sm = e ? atoi(e) : 1; }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = b * (off ? tiles1 : tiles0);
  const int grid = tiles < sms ? tiles : sms;
  const size_t smem = sizeof(TcSmem) + 1024;
  const int terms = similarity_mode() == 1 ? 1 : 3;
#define UPK_LAUNCH_TC(MODE, NT, ST)                                                                          \
  do {                                                                                                       \
    auto kern = k_similarity_tc<MODE, NT, ST>;                                                               \
    UPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
    kern<<<grid, TC_THREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, b, n, m, c, temp, off, out, stats_row,  \
                                         stats_col, stats_gref, ldc, nullptr, 0);                            \
  } while (0)
  const int tiles2 = b * ((n - off + 2 * TC_BM - 1) / (2 * TC_BM)) * ((m - off + TC_BN - 1) / TC_BN);
  if (use_2sm && sim_type == 0 && terms == 3 && tiles2 >= sms / 2) {
    CUtensorMap mb_hi2, mb_lo2;   // the pair's CTAs each fetch 128 of the 256 B rows of a tile
    if ((rc = make_map(&mb_hi2, b_hi, b, m, c, TC_BM))) return rc;
    if ((rc = make_map(&mb_lo2, b_lo, b, m, c, TC_BM))) return rc;
    const int grid2 = 2 * (tiles2 < sms / 2 ? tiles2 : sms / 2);
    CUtensorMap mc;
    const bool tmast = off == 1 && tma_store_ok(out, n, m, ldc);
    if (tmast) {
      if ((rc = make_map_out(&mc, out + ldc + 1, b, n, m, ldc))) return rc;
    } else {
      mc = ma_hi;   // unused by the kernel
    }
#define UPK_LAUNCH_TC2(ST, TM)                                                                                       \
  do {                                                                                                               \
    auto kern = k_similarity_tc2<ST, false, TM>;                                                                     \
    const size_t smem2 = sizeof(Tc2Smem<TM>) + 1024;                                                                 \
    UPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));              \
    kern<<<grid2, TC2_THREADS, smem2, st>>>(ma_hi, ma_lo, mb_hi2, mb_lo2, mc, b, n, m, c, temp, off, out, ldc,       \
                                            ST ? stats_row : nullptr, ST ? stats_col : nullptr, ST ? stats_gref : 0.f, ST ? fine_l2_keep(b, n, m) : 0); \
  } while (0)
    if (stats_row) { if (tmast) UPK_LAUNCH_TC2(true, true); else UPK_LAUNCH_TC2(true, false); }
    else { if (tmast) UPK_LAUNCH_TC2(false, true); else UPK_LAUNCH_TC2(false, false); }
#undef UPK_LAUNCH_TC2
  } else if (stats_row) {   // cosine logits + fused exponent sums (3xTF32 only: the statistics need fp32-level logits)
    UPK_LAUNCH_TC(0, 3, true);
  } else if (sim_type == 0) {
    if (terms == 3) UPK_LAUNCH_TC(0, 3, false); else UPK_LAUNCH_TC(0, 1, false);
  } else {
    if (terms == 3) UPK_LAUNCH_TC(1, 3, false); else UPK_LAUNCH_TC(1, 1, false);
  }
#undef UPK_LAUNCH_TC
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

// y[rows][out] = x[rows][in] W[out][in]^T + bias, optional ReLU: the single-CTA 3xTF32 kernel with batch 1, no peel,
// temp 1, operands split (not normalised) by the same preparation launch as the similarity path.
int run_linear_tc(const float* x, const float* w, const float* bias, int rows, int in, int out_f, int relu,
                  void* workspace, size_t workspace_bytes, float* y, cudaStream_t st) {
  if (workspace_bytes < similarity_tc_workspace_bytes(1, rows, out_f, in)) return UPK_ERR_INVALID_ARG;
  char* wsp = (char*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
  const size_t a = (((size_t)rows * in * sizeof(float)) + 1023) & ~(size_t)1023;
  const size_t bb = (((size_t)out_f * in * sizeof(float)) + 1023) & ~(size_t)1023;
  float* a_hi = (float*)wsp;
  float* a_lo = (float*)(wsp + a);
  float* b_hi = (float*)(wsp + 2 * a);
  float* b_lo = (float*)(wsp + 2 * a + bb);
  const dim3 g12(((rows > out_f ? rows : out_f) + 8 * NS_RPW - 1) / (8 * NS_RPW), 1, 2);
  const NsOperand oa{x, rows, a_hi, a_lo, nullptr, out_f, 1}, ob{w, out_f, b_hi, b_lo, nullptr, rows, 0};
  if (in <= 256) k_normalize_split<0, false, 4><<<g12, 256, 0, st>>>(oa, ob, in, 0, 1.0f, y, rows, out_f, out_f);
  else k_normalize_split<0, false, 2><<<g12, 256, 0, st>>>(oa, ob, in, 0, 1.0f, y, rows, out_f, out_f);
  count_launch(1);
  CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
  int rc;
  if ((rc = make_map(&ma_hi, a_hi, 1, rows, in, TC_BM))) return rc;
  if ((rc = make_map(&ma_lo, a_lo, 1, rows, in, TC_BM))) return rc;
  if ((rc = make_map(&mb_hi, b_hi, 1, out_f, in, TC_BN))) return rc;
  if ((rc = make_map(&mb_lo, b_lo, 1, out_f, in, TC_BN))) return rc;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles = ((rows + TC_BM - 1) / TC_BM) * ((out_f + TC_BN - 1) / TC_BN);
  const int grid = tiles < sms ? tiles : sms;
  const size_t smem = sizeof(TcSmem) + 1024;
  auto kern = k_similarity_tc<0, 3, false, true>;
  UPK_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<grid, TC_THREADS, smem, st>>>(ma_hi, ma_lo, mb_hi, mb_lo, 1, rows, out_f, in, 1.0f, 0, y, nullptr, nullptr, 0.f,
                                       out_f, bias, relu);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

}  // namespace upk

extern "C" size_t upk_linear_workspace_bytes(int rows, int in_features, int out_features) {
  if (rows <= 0 || in_features <= 0 || out_features <= 0) return 0;
  return upk::similarity_tc_workspace_bytes(1, rows, out_features, in_features);
}

extern "C" int upk_linear(const float* x, const float* weight, const float* bias, int rows, int in_features,
                          int out_features, int relu, void* workspace, size_t workspace_bytes, float* y,
                          upk_stream_t stream) {
  if (rows < 0 || in_features <= 0 || out_features <= 0) return UPK_ERR_INVALID_ARG;
  if (rows == 0) return UPK_OK;
  if (!x || !weight || !y || !workspace) return UPK_ERR_INVALID_ARG;
  if (in_features % upk::TC_BK != 0 || ((uintptr_t)x & 15) || ((uintptr_t)weight & 15) || !upk::similarity_tc_eligible(rows, out_features, in_features))
    return UPK_ERR_UNSUPPORTED;
  return upk::run_linear_tc(x, weight, bias, rows, in_features, out_features, relu, workspace, workspace_bytes, y,
                            (cudaStream_t)stream);
}

extern "C" int upk_set_similarity_mode(int mode) {
  int prev = upk::similarity_mode();
  if (mode == 0 || mode == 1 || mode == 3 || mode == 16) upk::g_sim_mode = mode;
  return prev;
}
