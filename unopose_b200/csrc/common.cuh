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
