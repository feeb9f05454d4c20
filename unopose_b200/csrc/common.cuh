// Shared device/host helpers for the unopose_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define UPK_OK 0
#define UPK_ERR_INVALID_ARG (-1)
#define UPK_ERR_UNSUPPORTED (-2)

// Launch-error convention: return the cudaError_t as a positive int, never
// exit() (deliberate deviation from the reference's CUDA_CHECK_ERRORS,
// _ext_src/include/cuda_utils.h:35-44, which calls exit(-1)).
#define UPK_RETURN_LAST_ERROR()                    \
  do {                                             \
    cudaError_t _e = cudaGetLastError();           \
    return _e == cudaSuccess ? UPK_OK : (int)_e;   \
  } while (0)

#define UPK_CUDA_TRY(expr)                         \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

namespace upk {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Squared distance with the exact rounding sequence the reference's nvcc build
// produces (default -fmad=true contraction of dx*dx + dy*dy + dz*dz):
//   fma(dz, dz, fma(dx, dx, dy*dy))
// read off the SASS of the reference extension compiled unmodified for sm_100a
// (FMUL on the y difference, then FFMA x, then FFMA z — identical in
// sampling_gpu.cu, ball_query_gpu.cu and interpolate_gpu.cu).  SURVEY.md
// Appendix A.1 states the x/y roles the other way round; the SASS wins.
__device__ __forceinline__ float sqdist_ref(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// torch.sum(v ** 2, dim=-1) over a last dimension of 3 ON THE GPU: the products are rounded separately and ATen's
// reduction kernel adds them as (v0^2 + v2^2) + v1^2 — read off a B200 run (scripts/dev/diag_bmm_k3.py: 0 of 1.2 M
// values differ with this order, 22 % with (v0^2 + v1^2) + v2^2, which is what CPU torch computes).  Together with
// cuBLAS's K = 3 dot product, fma(x2, y2, fma(x1, y1, x0 * y0)) (same experiment, 0 mismatches over four shapes), this
// makes the expansion-form squared distances of pairwise_distance (model_utils.py:246-256) bit-identical to the
// reference's GPU path for identical operands.
__device__ __forceinline__ float sumsq3_torch(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(z, z)), __fmul_rn(y, y));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// Blackwell packed fp32 pairs (FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest operations per issue
// slot, bit-identical to the scalar instructions lane by lane.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// Nearest-model-point search in the reference's expansion form (pairwise_distance, model_utils.py:246-256):
//   d_j = (|x|^2 - 2 x.y_j) + |y_j|^2 ,  x.y = fma(x2,yz, fma(x1,yy, x0*yx))
// over a model staged in shared memory as SoA arrays padded to a multiple of 4 (sentinel points far
// away).  Four model points per step: 4 LDS.128 + 10 packed ops + 4 FMNMX.  Returns min_j d_j.
__device__ __forceinline__ float nn_min_expansion(const float* __restrict__ mx, const float* __restrict__ my,
                                                  const float* __restrict__ mz, const float* __restrict__ mn,
                                                  int nm_pad, float x0, float x1, float x2, float xx,
                                                  int first = 0, const int step = 4) {
  const unsigned long long X0 = pack2(x0, x0), X1 = pack2(x1, x1), X2 = pack2(x2, x2);
  const unsigned long long XX = pack2(xx, xx), M2 = pack2(-2.0f, -2.0f);
  float best = INFINITY;
#pragma unroll 2
  for (int j = first; j < nm_pad; j += step) {  // (first, step) = (4 p, 4 P): P threads interleave one cloud
    const ulonglong2 qx = *reinterpret_cast<const ulonglong2*>(mx + j);
    const ulonglong2 qy = *reinterpret_cast<const ulonglong2*>(my + j);
    const ulonglong2 qz = *reinterpret_cast<const ulonglong2*>(mz + j);
    const ulonglong2 qn = *reinterpret_cast<const ulonglong2*>(mn + j);
    unsigned long long ta = fma2(X2, qz.x, fma2(X1, qy.x, mul2(X0, qx.x)));
    unsigned long long tb = fma2(X2, qz.y, fma2(X1, qy.y, mul2(X0, qx.y)));
    ta = add2(fma2(M2, ta, XX), qn.x);
    tb = add2(fma2(M2, tb, XX), qn.y);
    float d0, d1, d2, d3;
    unpack2(ta, d0, d1);
    unpack2(tb, d2, d3);
    best = fminf(fminf(best, d0), fminf(d1, fminf(d2, d3)));
  }
  return best;
}

// Two query points against the same staged model in one sweep (the two hypotheses a k_score thread evaluates):
// the four LDS.128 of a step are shared, the two dependency chains interleave.  Same arithmetic per point as
// nn_min_expansion.
__device__ __forceinline__ void nn_min_expansion2(const float* __restrict__ mx, const float* __restrict__ my,
                                                  const float* __restrict__ mz, const float* __restrict__ mn,
                                                  int nm_pad, const float (&xa)[4], const float (&xb)[4],
                                                  float& best_a, float& best_b, int first = 0, const int step = 4) {
  const unsigned long long A0 = pack2(xa[0], xa[0]), A1 = pack2(xa[1], xa[1]), A2 = pack2(xa[2], xa[2]);
  const unsigned long long AX = pack2(xa[3], xa[3]);
  const unsigned long long B0 = pack2(xb[0], xb[0]), B1 = pack2(xb[1], xb[1]), B2 = pack2(xb[2], xb[2]);
  const unsigned long long BX = pack2(xb[3], xb[3]);
  const unsigned long long M2 = pack2(-2.0f, -2.0f);
  float ba = INFINITY, bb = INFINITY;
#pragma unroll 2
  for (int j = first; j < nm_pad; j += step) {
    const ulonglong2 qx = *reinterpret_cast<const ulonglong2*>(mx + j);
    const ulonglong2 qy = *reinterpret_cast<const ulonglong2*>(my + j);
    const ulonglong2 qz = *reinterpret_cast<const ulonglong2*>(mz + j);
    const ulonglong2 qn = *reinterpret_cast<const ulonglong2*>(mn + j);
    unsigned long long ta = fma2(A2, qz.x, fma2(A1, qy.x, mul2(A0, qx.x)));
    unsigned long long tb = fma2(A2, qz.y, fma2(A1, qy.y, mul2(A0, qx.y)));
    unsigned long long ua = fma2(B2, qz.x, fma2(B1, qy.x, mul2(B0, qx.x)));
    unsigned long long ub = fma2(B2, qz.y, fma2(B1, qy.y, mul2(B0, qx.y)));
    ta = add2(fma2(M2, ta, AX), qn.x);
    tb = add2(fma2(M2, tb, AX), qn.y);
    ua = add2(fma2(M2, ua, BX), qn.x);
    ub = add2(fma2(M2, ub, BX), qn.y);
    float d0, d1, d2, d3, e0, e1, e2, e3;
    unpack2(ta, d0, d1);
    unpack2(tb, d2, d3);
    unpack2(ua, e0, e1);
    unpack2(ub, e2, e3);
    ba = fminf(fminf(ba, d0), fminf(d1, fminf(d2, d3)));
    bb = fminf(fminf(bb, e0), fminf(e1, fminf(e2, e3)));
  }
  best_a = ba;
  best_b = bb;
}

// stage a model cloud (AoS global) into the SoA layout nn_min_expansion expects; call with all threads
__device__ __forceinline__ void stage_model_soa(const float* __restrict__ model, int nm, int nm_pad, float* mx,
                                                float* my, float* mz, float* mn) {
  for (int j = threadIdx.x; j < nm_pad; j += blockDim.x) {
    float x = 1e15f, y = 1e15f, z = 1e15f;
    if (j < nm) { x = model[j * 3 + 0]; y = model[j * 3 + 1]; z = model[j * 3 + 2]; }
    mx[j] = x; my[j] = y; mz[j] = z;
    mn[j] = sumsq3_torch(x, y, z);   // y2 = torch.sum(y ** 2, -1)
  }
}

// Streaming (read-once) 128-bit global load that does not pollute L1.
__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ld_stream_i4(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

}  // namespace upk
