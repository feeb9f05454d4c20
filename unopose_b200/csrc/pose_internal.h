// Internal (non-ABI) declarations shared by the pose translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace upk {

// Bump allocator over a caller-provided workspace; with base == nullptr it only
// measures (used by the *_workspace_bytes entry points).
struct Carver {
  char* base;
  size_t off;
  explicit Carver(void* p) : base((char*)p), off(0) {}
  template <class T>
  T* take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* r = base ? (T*)(base + off) : (T*)nullptr;
    off += n * sizeof(T);
    return r;
  }
  size_t bytes() const { return (off + 255) & ~(size_t)255; }
};

// Tile geometry of the assignment passes (assign.cu)
struct AssignGeom {
  int R, C;      // rows / cols of atten (incl. background row/col 0)
  int TR, TC;    // tile shape
  int ntr, ntc;  // tile counts
};
AssignGeom assign_geom(int R, int C);

// Workspace of the dual-softmax assignment passes.
struct AssignWs {
  float2* rowpart;  // [b][R][ntc]  (max, sumexp) partials
  float2* colpart;  // [b][C][ntr]
  float* rmax;      // [b][R]
  float* rsum;
  float* cmax;      // [b][C]
  float* csum;
  float* rowpm;     // [b][R][ntc]  partial row max of A over cols >= 1
  float* colpm;     // [b][C][ntr]  partial col max of A over rows >= 1
  float* ai0;       // [b][R]  A[i][0]
  float* a0j;       // [b][C]  A[0][j]
  float* bsum;      // [b][2] fused-statistics path: exponent sums of the background row / column
  int* flags;       // [b]  large geometry: pass 1's single-reference sums were not trustworthy -> exact redo
};
void carve_assign(Carver& cv, int b, const AssignGeom& g, AssignWs& ws);

// stats + labels: fills ws.{rmax,rsum,cmax,csum} and w1 [b][R-1], w2 [b][C-1].
// for_fine_solve = false (the coarse solver): natural-log maxima / sums from the exact tile pipeline, which is what
// run_coarse_P consumes.  for_fine_solve = true on a large geometry: the streaming passes of assign_fine.cu, which
// leave log2-domain constants (rml, rmul, cml, cmul) in the same buffers for run_fine_rowsums.
// atten_ld: row pitch of atten in floats (== g.C for a contiguous tensor; only the large-geometry fine passes accept
// a pitched one, UPK_ERR_UNSUPPORTED otherwise).
int run_assignment_labels(const float* atten, const float* score1, int ld1, const float* score2, int ld2,
                          int b, const AssignGeom& g, const AssignWs& ws, float* w1, float* w2,
                          cudaStream_t st, bool for_fine_solve = false, int atten_ld = 0);

// coarse_assign.cu: masks + sampling CDF in one cluster kernel, bit-exact with the ATen CUDA kernels of the reference's
// GPU path.  UPK_ERR_UNSUPPORTED for geometries it does not handle (-> the tile pipeline below).
int run_coarse_assign_exact(const float* atten, const float* score1, int ld1, const float* score2, int ld2, int b,
                            int R, int C, float* w1, float* w2, float* cdf, cudaStream_t st, long long* stamps = nullptr);

// coarse: P = (A w1 w2)^1.5 over the foreground block -> pmat [b][N1*N2], row-sum partials (double)
int run_coarse_P(const float* atten, const float* score1, int ld1, const float* score2, int ld2, int b,
                 const AssignGeom& g, const AssignWs& ws, const float* w1, const float* w2, float* pmat,
                 double* prow /*[b][N1][ntc]*/, cudaStream_t st);
// cdf[k] = float(prefix_k) / (float(total) + 1e-8)
int run_cdf(const float* pmat, const double* prow, int b, int n1, int n2, int ntc, float* cdf, cudaStream_t st);

// fine: per-row  sum_j A_ij w2_j {x,y,z,1}_j  (times w1_i)  -> soft [b][N1][3], asum [b][N1]
int run_fine_rowsums(const float* atten, const float* score1, int ld1, const float* score2, int ld2, int b,
                     const AssignGeom& g, const AssignWs& ws, const float* w1, const float* w2,
                     const float* pts2, float4* rowpart4 /*[b][N1][ntc]*/, float* soft, float* asum,
                     cudaStream_t st, int atten_ld = 0);

// Large-geometry (fine) streaming passes, assign_fine.cu.  They reuse the AssignWs buffers: rmax/rsum/cmax/csum
// hold rml (row reference, log2 units) / rmul (score1 / row sum) / cml / cmul there.
struct FineGeom2 {
  int nstrip, nrt;  // 256-column strips / 128-row tiles of the main block (background row and column peeled off)
};
FineGeom2 fine_geom2(int R, int C);
int run_fine_labels2(const float* atten, int ld, const float* score1, int ld1, const float* score2, int ld2, int b,
                     const AssignGeom& g, const AssignWs& ws, float* w1, float* w2, cudaStream_t st);
int run_fine_rows2(const float* atten, int ld, int b, const AssignGeom& g, const AssignWs& ws, const float* w1,
                   const float* w2, const float* pts2, float4* rowpart4, float* soft, float* asum, cudaStream_t st);
// helpers living in assign.cu
int launch_labels_merge(const float* rowpm, const float* colpm, const float* ai0, const float* a0j, int b, int R,
                        int C, int ntr, int ntc, float* w1, float* w2, cudaStream_t st);
int launch_fine_rows_merge(const float4* rowpart4, const float* w1, int b, int n1, int ntc, float* soft,
                           float* asum, cudaStream_t st);

// tensor-core similarity path (similarity_tc.cu)
int similarity_mode();  // 16 = 3xFP16/3xTF32 tcgen05 (default), 3 = 3xTF32, 1 = 1xTF32 tcgen05, 0 = fp32 SIMT
bool similarity_tc_eligible(int n, int m, int c);
size_t similarity_tc_workspace_bytes(int b, int n, int m, int c);
// ldc: row pitch of `out` in floats (0 = contiguous, m).  With ldc % 4 == 0 and element (1, 1) 16-byte aligned the
// CTA-pair kernel stores its tiles with TMA.
int run_similarity_tc(const float* f1, const float* f2, int b, int n, int m, int c, float temp, int normalize,
                      int sim_type, void* workspace, size_t workspace_bytes, float* out, cudaStream_t st,
                      float* stats_row = nullptr, float* stats_col = nullptr, float stats_gref = 0.f, int ldc = 0);

// TMA descriptor of a (batch, rows, K) fp32 K-major operand, boxes of TC_BK x box_rows, SWIZZLE_64B
// (`map` is a CUtensorMap*; UPK_ERR_UNSUPPORTED if the driver entry point is unavailable)
int tc_make_map(void* map, const float* base, int batch, int rows, int K, int box_rows);
// same for fp16 operands (32 halves per 64-byte box row); experimental 3xFP16 paths
int tc_make_map_f16(void* map, const void* base, int batch, int rows, int K, int box_rows);
int fine_l2_keep(int b, int n, int m);   // instances of a fine-shape atten kept in L2 between the fine-stage kernels (0 = no hints)
bool tc_tma_available();   // the driver entry point for tensor-map encoding exists
// unswizzled fp32 boxes (box_cols x box_rows) of a pitched (batch, rows, cols) matrix; `base` 16-byte aligned, `ld` and
// `batch_stride` in floats (ld % 4 == 0)
int tc_make_map_plane(void* map, const float* base, int batch, int rows, int cols, int ld, size_t batch_stride,
                      int box_cols, int box_rows);

// Exponent-sum partials produced by the similarity GEMM's epilogue (cosine logits; fused pass 1 of the fine
// assignment): rowpart [b][npr][n] then colpart [b][npc][m] (256-byte aligned), all relative to the ONE
// reference exponent sim_stats_gref(temp).
struct SimStatsGeom {
  int npr, npc;        // partials per row (2 per 256-column tile) / per column (4 per 128-row tile)
  size_t row_floats, col_off_floats, total_bytes;
};
SimStatsGeom sim_stats_geom(int b, int n, int m);
float sim_stats_gref(float temp);
int run_fine_labels2_fused(const float* atten, int ld, const float* stats, float temp, const float* score1, int ld1,
                           const float* score2, int ld2, int b, const AssignGeom& g, const AssignWs& ws, float* w1,
                           float* w2, cudaStream_t st);

}  // namespace upk
