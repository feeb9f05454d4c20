// (1b)+(2) Coarse dual-softmax assignment and sampling CDF in ONE kernel, bit-exact with the reference's GPU path.
//
// Reference: compute_coarse_Rt[_overlap], core/unopose/utils/model_utils.py:443-461 —
//   A   = softmax(atten, 2) * softmax(atten, 1) * score1 * score2
//   w1  = argmax_j A[1:, :] > 0 ,  w2 = argmax_i A[:, 1:] > 0
//   P   = (A[1:, 1:] * w1 (x) w2) ** 1.5 , cdf = cumsum(P) / (cumsum(P)[-1] + 1e-8)
// i.e. ~14 ATen launches on the reference's GPU path.  The uniform draws of the hypothesis sampler are compared against
// this CDF (torch.searchsorted, :462), so every bit of it decides which correspondence a draw selects.  This kernel
// therefore reproduces the ARITHMETIC ORDER of the ATen CUDA kernels the reference runs (torch 2.11, verified bit for
// bit on a B200 by scripts/r2_parity_probe.py and tests/test_pose_gpu.py::test_coarse_cdf_bit_exact_vs_torch):
//   * softmax over the last dim (C <= 1024): ATen/native/cuda/PersistentSoftmax.cuh softmax_warp_forward — a warp per
//     row, lane l accumulates exp(x - max) of elements l, l + W, l + 2W, ... in that order, then a xor-butterfly
//     (offsets W/2 .. 1), result exp(x - max) / sum with an IEEE division;
//   * softmax over dim 1: cunn_SpatialSoftMaxForward with blockDim.x == 1 (inner size > 64): one thread per column,
//     sum over the rows 0 .. R-1 strictly in order;
//   * x ** 1.5: powf(x, 1.5f);
//   * cumsum over the flattened (N1*N2) row: ATen/native/cuda/ScanUtils.cuh tensor_kernel_scan_innermost_dim — blocks of
//     2^(lx+1) elements (lx from get_log_num_threads_x_inner_scan(num_rows = B, row_size)), the running total added to
//     element 0 of the next block, a Sklansky scan inside the block;
//   * the final division by (last + 1e-8f).
// For B == 1 torch routes cumsum to cub::DeviceScan instead (decoupled look-back; CUB documents run-to-run variation
// for floating-point addition), so there is no bit-exact reference there; the same scan is used.
//
// Parallelisation: a thread-block CLUSTER of 8 CTAs per instance (distributed shared memory).  Rows are dealt to the
// CTAs in slabs; the strictly sequential column sums are dealt by column (each CTA pulls its columns' exponentials from
// the other slabs through DSMEM and sums them in order); the scan is split into its carry-free part (Sklansky inside
// every dyadic sub-block [2^k, 2^(k+1)) of every block, all blocks in parallel) and the carry chain (11 dependent
// additions per block, one thread), which is exactly the data flow of the ATen kernel:
//   final[i] = q_i + T_msb(i),  T_0 = x_0 + carry,  T_m = r_m + T_(m-1),  carry' = T_(lx+1),
// with q_i the scan of i inside its dyadic sub-block and r_m = q_(2^m - 1).
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "launch_count.h"
#include "pose_internal.h"
#include "../../include/unopose_b200.h"

namespace cg = cooperative_groups;

namespace upk {

constexpr int CA_CL = 8;          // CTAs per instance (portable cluster size)
constexpr int CA_THREADS = 512;       // 2 CTAs per SM (64 registers, < 113 KB smem): all 16 clusters of a batch resident at once
constexpr int CA_WARPS = CA_THREADS / 32;
constexpr int CA_RVS = 12;        // floats per block in the carry exchange (x_0, r_1 .. r_(lx+1) with lx <= 9)
constexpr int CA_MAX_BLOCKS = 256;

struct CaGeom {
  int R, C, rpc, cpc, lx, wb, nblk, nown;
  size_t main_floats, smem_bytes;
};

static int ca_log_threads_x(int num_rows, int row_size) {   // ScanUtils.cuh get_log_num_threads_x_inner_scan
  int lx = 0, ly = 0;
  while ((1 << lx) < row_size) ++lx;
  while ((1 << ly) < num_rows) ++ly;
  lx = (9 + (lx - ly)) / 2;
  return lx < 4 ? 4 : (lx > 9 ? 9 : lx);
}

static bool ca_geometry(int b, int R, int C, CaGeom& g) {
  if (R < 2 || C < 2 || R > 1024 || C > 1024) return false;
  if (C <= 64 && R >= 64) return false;   // torch's spatial softmax then splits the reduction over dim: other order
  g.R = R;
  g.C = C;
  g.rpc = ceil_div(R, CA_CL);
  g.cpc = ceil_div(C, CA_CL);
  const long long L = (long long)(R - 1) * (C - 1);
  g.lx = ca_log_threads_x(b, (int)L);
  g.wb = 2 << g.lx;
  g.nblk = (int)((L + g.wb - 1) / g.wb);
  if (g.nblk > CA_MAX_BLOCKS) return false;
  g.nown = ceil_div(g.nblk, CA_CL);
  const size_t slabs = 2 * (size_t)g.rpc * C;
  const size_t scan = (size_t)g.nown * g.wb;
  g.main_floats = ((slabs > scan ? slabs : scan) + 3) & ~(size_t)3;
  const size_t small = 6 * (size_t)C + 3 * (size_t)g.rpc + 2 * (size_t)(CA_MAX_BLOCKS + 1) * CA_RVS + 16;
  g.smem_bytes = (g.main_floats + (size_t)R * g.cpc + small) * sizeof(float);
  return g.smem_bytes <= 200 * 1024;
}

__device__ __forceinline__ int ca_next_pow2(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

__global__ void __cluster_dims__(CA_CL, 1, 1) __launch_bounds__(CA_THREADS, 2)
k_coarse_assign_exact(const float* __restrict__ atten, const float* __restrict__ score1, int ld1,
                      const float* __restrict__ score2, int ld2, int R, int C, int rpc, int cpc, int lx,
                      int main_floats, float* __restrict__ w1_out, float* __restrict__ w2_out,
                      float* __restrict__ cdf, long long* __restrict__ stamps) {
  extern __shared__ __align__(16) float ca_smem[];
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N1 = R - 1, N2 = C - 1;
  const int r0 = rank * rpc, nr = max(0, min(rpc, R - r0));
  const int c0 = rank * cpc, nc = max(0, min(cpc, C - c0));
  // ---- carve
  float* xs = ca_smem;                        // [rpc][C]  logits of my rows, later exp(x - colmax)
  float* e2 = xs + (size_t)rpc * C;           // [rpc][C]  exp(x - rowmax), later A
  float* scan = ca_smem;                      // [nown][wb]  (aliases the slabs once P is in global memory)
  float* rv = ca_smem + main_floats;          // [(CA_MAX_BLOCKS + 1) * CA_RVS]  (rank 0: carry-free block values); 16-byte aligned
  float* tv = rv + (CA_MAX_BLOCKS + 1) * CA_RVS;   // same size (rank 0: the carry chain, + denominator)
  float* colbuf = tv + (CA_MAX_BLOCKS + 1) * CA_RVS;   // [R][cpc]  my columns of everybody's exp(x - colmax)
  float* cpart = colbuf + (size_t)R * cpc;    // [C]  per-CTA column partials (max of x, later max of A)
  float* cmax = cpart + C;                    // [C]
  float* csum = cmax + C;                     // [C]
  float* a0j = csum + C;                      // [C]  A[0][j] (filled by rank 0)
  float* w2s = a0j + C;                       // [C]
  float* s2v = w2s + C;                       // [C]
  float* rsum = s2v + C;                      // [rpc]
  float* w1s = rsum + rpc;                    // [rpc]
  float* s1v = w1s + rpc;                     // [rpc]
  const float* Ab = atten + (size_t)b * R * C;
  float* cdf_b = cdf + (size_t)b * N1 * N2;
  int n_stamp = 0;
  // profiling aid (upk_coarse_assignment_profile): SM clock of rank 0 / instance 0 at every phase boundary
#define CA_STAMP() do { if (stamps && b == 0 && rank == 0 && tid == 0) stamps[n_stamp++] = clock64(); } while (0)
  CA_STAMP();

  // ---- load my slab (rows are contiguous in global memory)
  for (int i = tid; i < nr * C; i += CA_THREADS) xs[i] = __ldg(Ab + (size_t)r0 * C + i);
  for (int i = tid; i < nr; i += CA_THREADS) {
    const int gi = r0 + i;
    s1v[i] = (gi > 0 && score1) ? score1[(size_t)b * ld1 + gi - 1] : 1.f;   // the background row / column carry 1.0
  }
  for (int j = tid; j < C; j += CA_THREADS) s2v[j] = (j > 0 && score2) ? score2[(size_t)b * ld2 + j - 1] : 1.f;
  __syncthreads();

  // ---- softmax over the last dim: statistics in softmax_warp_forward order
  {
    const int p2 = ca_next_pow2(C);
    const int W = p2 < 32 ? p2 : 32;
    const int iters = p2 / W;
    for (int i = warp; i < nr; i += CA_WARPS) {
      const float* row = xs + (size_t)i * C;
      float m = -INFINITY;
      for (int it = 0; it < iters; ++it) {
        const int idx = lane + it * W;
        const float v = (lane < W && idx < C) ? row[idx] : -INFINITY;
        m = m > v ? m : v;
      }
      for (int off = W >> 1; off > 0; off >>= 1) {
        const float o = __shfl_xor_sync(kFull, m, off);
        m = m < o ? o : m;
      }
      float s = 0.f;
      for (int it = 0; it < iters; ++it) {
        const int idx = lane + it * W;
        float e = 0.f;
        if (lane < W && idx < C) {
          e = expf(row[idx] - m);
          e2[(size_t)i * C + idx] = e;
        }
        s += e;
      }
      for (int off = W >> 1; off > 0; off >>= 1) s = s + __shfl_xor_sync(kFull, s, off);
      if (lane == 0) rsum[i] = s;
    }
  }
  // column maxima over my rows
  for (int j = tid; j < C; j += CA_THREADS) {
    float m = -INFINITY;
    for (int i = 0; i < nr; ++i) m = fmaxf(m, xs[(size_t)i * C + j]);
    cpart[j] = m;
  }
  cl.sync();   // #1: every CTA's column partials are visible
  CA_STAMP();

  for (int j = tid; j < C; j += CA_THREADS) {
    float m = -INFINITY;
#pragma unroll
    for (int r = 0; r < CA_CL; ++r) m = fmaxf(m, cl.map_shared_rank(cpart, r)[j]);
    cmax[j] = m;
  }
  __syncthreads();
  for (int r = warp; r < nr; r += CA_WARPS)
    for (int j = lane; j < C; j += 32) xs[(size_t)r * C + j] = expf(xs[(size_t)r * C + j] - cmax[j]);
  cl.sync();   // #2: exp(x - colmax) of all slabs ready; all reads of the x-max partials are done
  CA_STAMP();

  // ---- softmax over dim 1: sum_i exp(x_ij - colmax_j) for i = 0 .. R-1 IN ORDER, for my columns
  {
    const int total = R * nc;
    for (int k0 = tid; k0 < total; k0 += 4 * CA_THREADS) {   // four independent DSMEM loads in flight per thread
      float v[4];
      int dst[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * CA_THREADS;
        dst[u] = -1;
        if (k < total) {
          const int i = k / nc, jj = k - i * nc;
          const int rr = i / rpc;
          v[u] = cl.map_shared_rank(xs, rr)[(size_t)(i - rr * rpc) * C + c0 + jj];
          dst[u] = i * cpc + jj;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (dst[u] >= 0) colbuf[dst[u]] = v[u];
    }
  }
  __syncthreads();
  if (tid < nc) {
    float s = 0.f;
#pragma unroll 8
    for (int i = 0; i < R; ++i) s += colbuf[(size_t)i * cpc + tid];
#pragma unroll
    for (int r = 0; r < CA_CL; ++r) cl.map_shared_rank(csum, r)[c0 + tid] = s;
  }
  cl.sync();   // #3: column sums everywhere
  CA_STAMP();

  // ---- A = ((softmax_2 * softmax_1) * s1) * s2 for my rows
  // and w1_i = (max_{j >= 1} A_ij > A_i0): torch.max keeps the FIRST maximum, so label > 0 needs a strictly larger entry
  for (int r = warp; r < nr; r += CA_WARPS) {
    const float rs = rsum[r], s1 = s1v[r];
    float m = -INFINITY, a0 = 0.f;
    for (int j = lane; j < C; j += 32) {
      const size_t i = (size_t)r * C + j;
      float a = __fdiv_rn(e2[i], rs) * __fdiv_rn(xs[i], csum[j]);
      a = a * s1;
      a = a * s2v[j];
      e2[i] = a;
      if (j > 0) m = fmaxf(m, a); else a0 = a;
    }
    m = warp_max(m);
    a0 = __shfl_sync(kFull, a0, 0);
    if (lane == 0) w1s[r] = (r0 + r > 0 && m > a0) ? 1.f : 0.f;
  }
  __syncthreads();
  for (int j = tid; j < C; j += CA_THREADS) {
    float m = -INFINITY;
    for (int i = (r0 == 0 ? 1 : 0); i < nr; ++i) m = fmaxf(m, e2[(size_t)i * C + j]);
    cpart[j] = m;
    if (rank == 0) a0j[j] = e2[j];
  }
  cl.sync();   // #4: column maxima of A, A[0][:]
  CA_STAMP();

  for (int j = tid; j < C; j += CA_THREADS) {
    float m = -INFINITY;
#pragma unroll
    for (int r = 0; r < CA_CL; ++r) m = fmaxf(m, cl.map_shared_rank(cpart, r)[j]);
    w2s[j] = (j > 0 && m > cl.map_shared_rank(a0j, 0)[j]) ? 1.f : 0.f;
  }
  __syncthreads();
  for (int i = tid; i < nr; i += CA_THREADS)
    if (r0 + i > 0) w1_out[(size_t)b * N1 + r0 + i - 1] = w1s[i];
  if (rank == 0)
    for (int j = 1 + tid; j < C; j += CA_THREADS) w2_out[(size_t)b * N2 + j - 1] = w2s[j];

  // ---- P = ((A w1) w2) ** 1.5 over the foreground block, staged in the cdf buffer (global)
  for (int r = warp; r < nr; r += CA_WARPS) {
    const int gi = r0 + r;
    if (gi == 0) continue;
    const float wr = w1s[r];
    float* dst = cdf_b + (size_t)(gi - 1) * N2 - 1;
    for (int j = 1 + lane; j < C; j += 32) {
      const float a = (e2[(size_t)r * C + j] * wr) * w2s[j];
      dst[j] = a == 0.f ? 0.f : powf(a, 1.5f);     // powf(+0, 1.5) = +0: masked entries skip the slow path
    }
  }
  __threadfence();
  cl.sync();   // #5: P complete in global memory; the slabs are dead from here on
  CA_STAMP();

  // ---- cumsum, carry-free part: blocks of wb = 2^(lx+1) elements dealt round-robin to the CTAs
  const int wb = 2 << lx, ntx = 1 << lx;
  const int L = N1 * N2;
  const int nblk = (L + wb - 1) / wb;
  {
    int n = 0;
    for (int blk = rank; blk < nblk; blk += CA_CL) ++n;
    const int nmine = n;
    // all of this thread's loads first (independent L2 round trips), then the shared-memory stores
    {
      const int per = (nmine * wb + CA_THREADS - 1) / CA_THREADS;   // <= 8 for the coarse shape
      float v[8];
      for (int c = 0; c < per; c += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = (c + u) * CA_THREADS + tid;      // element of my concatenated blocks
          const int q = e >> (lx + 1), t = e & (wb - 1);
          const int g = (rank + q * CA_CL) * wb + t;
          v[u] = (c + u < per && q < nmine && g < L) ? __ldcg(cdf_b + g) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int e = (c + u) * CA_THREADS + tid;
          if (c + u < per && e < nmine * wb) scan[e] = v[u];
        }
      }
    }
    __syncthreads();
    // Sklansky levels: thread t of a block updates element ti(t, m) from si(t, m).  For m <= 5 both lie in the 64-element
    // segment of t's warp, so a warp barrier orders the levels; from m = 6 on the whole CTA synchronises.
    for (int m = 0; m <= lx; ++m) {
      const int s = 1 << m;
      if (tid < ntx && (tid >> m) != 0) {   // (t >> m) == 0: targets in [2^m, 2^(m+1)) receive T_m of the carry chain instead
        const int a = ((tid >> m) << (m + 1)) | s;
        const int ti = a + (tid & (s - 1));
        for (int q = 0; q < nmine; ++q) {
          float* buf = scan + (size_t)q * wb;
          buf[ti] = buf[ti] + buf[a - 1];
        }
      }
      if (m < 5) __syncwarp(); else __syncthreads();
    }
    // hand x_0, r_1 .. r_(lx+1) of my blocks (and q of the last element of the row) to rank 0
    float* rv0 = cl.map_shared_rank(rv, 0);
    n = 0;
    for (int blk = rank; blk < nblk; blk += CA_CL, ++n) {
      const float* buf = scan + (size_t)n * wb;
      if (tid <= lx + 1) rv0[blk * CA_RVS + tid] = tid == 0 ? buf[0] : buf[(1 << tid) - 1];
      if (blk == (L - 1) / wb && tid == 32) rv0[nblk * CA_RVS] = buf[(L - 1) % wb];
    }
  }
  cl.sync();   // #6: block values at rank 0
  CA_STAMP();

  if (rank == 0 && tid == 0) {
    // 11 dependent additions per block (128-bit shared-memory accesses; the next block's operands are fetched meanwhile)
    float carry = 0.f;
    const float4* rv4 = reinterpret_cast<const float4*>(rv);
    float4* tv4 = reinterpret_cast<float4*>(tv);
    float4 n0 = rv4[0], n1 = rv4[1], n2 = rv4[2];
    for (int blk = 0; blk < nblk; ++blk) {
      const float r[CA_RVS] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w, n2.x, n2.y, n2.z, n2.w};
      if (blk + 1 < nblk) { n0 = rv4[(blk + 1) * 3]; n1 = rv4[(blk + 1) * 3 + 1]; n2 = rv4[(blk + 1) * 3 + 2]; }
      float T = r[0] + carry;                // row_buf[0] = row_buf[0] + block_total
      float out[CA_RVS];
      out[0] = T;
#pragma unroll
      for (int m = 1; m < CA_RVS; ++m) {
        if (m <= lx + 1) T = r[m] + T;
        out[m] = T;
      }
      carry = T;
      tv4[blk * 3] = make_float4(out[0], out[1], out[2], out[3]);
      tv4[blk * 3 + 1] = make_float4(out[4], out[5], out[6], out[7]);
      tv4[blk * 3 + 2] = make_float4(out[8], out[9], out[10], out[11]);
    }
    const int lb = (L - 1) / wb, pos = (L - 1) % wb;
    const float last = pos == 0 ? tv[lb * CA_RVS] : rv[nblk * CA_RVS] + tv[lb * CA_RVS + (31 - __clz(pos))];
    tv[nblk * CA_RVS] = last + 1e-8f;      // cumsum[:, -1] + 1e-8
  }
  cl.sync();   // #7: carry chain and denominator at rank 0
  CA_STAMP();

  {
    const float* tv0 = cl.map_shared_rank(tv, 0);
    const float denom = tv0[nblk * CA_RVS];
    float* tl = rv;      // this CTA's copy of its blocks' chain values (rv is free: unused on ranks != 0, consumed on rank 0)
    int n = 0;
    for (int blk = rank; blk < nblk; blk += CA_CL, ++n)
      if (tid <= lx + 1) tl[n * CA_RVS + tid] = tv0[blk * CA_RVS + tid];
    __syncthreads();
    n = 0;
    for (int blk = rank; blk < nblk; blk += CA_CL, ++n) {
      const float* buf = scan + (size_t)n * wb;
      for (int t = tid; t < wb; t += CA_THREADS) {
        const int g = blk * wb + t;
        if (g < L) {
          const float cs = t == 0 ? tl[n * CA_RVS] : buf[t] + tl[n * CA_RVS + (31 - __clz(t))];
          cdf_b[g] = __fdiv_rn(cs, denom);
        }
      }
    }
  }
  cl.sync();   // #8: nobody exits while a peer may still read its shared memory
  CA_STAMP();
#undef CA_STAMP
}

// w1 [b][R-1], w2 [b][C-1], cdf [b][(R-1)(C-1)].  UPK_ERR_UNSUPPORTED when the geometry is outside what the cluster kernel
// handles (the caller then takes the tile pipeline of assign.cu, which is close to but not bit-identical with torch).
int run_coarse_assign_exact(const float* atten, const float* score1, int ld1, const float* score2, int ld2, int b,
                            int R, int C, float* w1, float* w2, float* cdf, cudaStream_t st, long long* stamps) {
  static int enabled = -1;
  if (enabled < 0) { const char* e = getenv("UPK_COARSE_EXACT"); enabled = e ? atoi(e) : 1; }
  CaGeom g;
  if (!enabled || !ca_geometry(b, R, C, g)) return UPK_ERR_UNSUPPORTED;
  UPK_CUDA_TRY(cudaFuncSetAttribute(k_coarse_assign_exact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_bytes));
  k_coarse_assign_exact<<<dim3(CA_CL, b), CA_THREADS, g.smem_bytes, st>>>(atten, score1, ld1, score2, ld2, R, C, g.rpc,
                                                                          g.cpc, g.lx, (int)g.main_floats, w1, w2, cdf, stamps);
  count_launch();
  UPK_RETURN_LAST_ERROR();
}

}  // namespace upk

// Profiling aid: the same kernel as upk_coarse_assignment's, additionally recording the SM clock (clock64) of instance
// 0 / cluster rank 0 at kernel entry and after each of its 8 cluster barriers into stamps_out[9] (device memory).
extern "C" int upk_coarse_assignment_profile(const float* atten, const float* score1, int score1_ld, const float* score2,
                                             int score2_ld, int b, int n1, int n2, float* w1_out, float* w2_out,
                                             float* cdf_out, long long* stamps_out, upk_stream_t stream) {
  if (b <= 0 || n1 <= 0 || n2 <= 0 || !atten || !w1_out || !w2_out || !cdf_out || !stamps_out) return UPK_ERR_INVALID_ARG;
  return upk::run_coarse_assign_exact(atten, score1, score1_ld, score2, score2_ld, b, n1 + 1, n2 + 1, w1_out, w2_out,
                                      cdf_out, (cudaStream_t)stream, stamps_out);
}
