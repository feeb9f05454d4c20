// Dev microbenchmark: dependent-chain latencies (cycles) of the cross-lane / barrier primitives the FPS loop uses.
#include <cstdio>
#include <cuda_runtime.h>
#define FULL 0xffffffffu
template <int OP>
__global__ void k(int iters, int* out, long long* cyc) {
  __shared__ int sm[64];
  int v = threadIdx.x * 7 + 1;
  sm[threadIdx.x & 63] = v;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (OP == 0) v = __reduce_max_sync(FULL, v) + (threadIdx.x & 1);
    if (OP == 1) v = __popc(__ballot_sync(FULL, v & 1)) + v;
    if (OP == 2) v = __shfl_sync(FULL, v, (v + 1) & 31) + 1;
    if (OP == 3) { __syncthreads(); v += 1; }
    if (OP == 4) v = sm[v & 63] + 1;
    if (OP == 5) { v = __reduce_max_sync(FULL, v); v = __reduce_min_sync(FULL, v ^ threadIdx.x) + 1; }
    if (OP == 6) { if (__any_sync(FULL, v & 4)) v += 3; else v += 1; }
    if (OP == 7) { sm[threadIdx.x & 63] = v; __syncthreads(); v = sm[(threadIdx.x + 1) & 63] + 1; }
    if (OP == 8) { v = __float_as_int(sqrtf(__int_as_float(v & 0x3fffffff))) + 1; }
    if (OP == 9) { v = __ffs(v) + v; }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = v;
}
int main() {
  int* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  const char* names[] = {"redux.max", "ballot+popc", "shfl(var lane)", "bar.sync", "lds", "redux.max+redux.min", "any+branch", "sts+bar+lds", "sqrtf", "ffs"};
  const int iters = 4096;
  for (int threads : {32, 512, 1024}) {
    auto run = [&](auto kern, int op) {
      kern<<<1, threads>>>(iters, out, cyc);
      kern<<<1, threads>>>(iters, out, cyc);
      cudaDeviceSynchronize();
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("threads=%4d %-22s %.1f cyc/iter\n", threads, names[op], (double)c / iters);
    };
    run(k<0>, 0); run(k<1>, 1); run(k<2>, 2); run(k<3>, 3); run(k<4>, 4); run(k<5>, 5); run(k<6>, 6); run(k<7>, 7); run(k<8>, 8); run(k<9>, 9);
  }
  return 0;
}
