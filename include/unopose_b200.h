/*
 * unopose_b200 — C ABI of the B200-native (sm_100a) correspondence-and-pose hot path.
 *
 * Drop-in boundary for the reference's native extension
 *   core/unopose/model/pointnet2/_ext   (pybind module, _ext_src/src/bindings.cpp:11-24)
 * and for the torch-op sequences of
 *   core/unopose/utils/model_utils.py   (compute_*_Rt*, weighted_procrustes, ...).
 *
 * Conventions
 *   - plain device pointers + sizes + a CUDA stream; no torch types.
 *   - every tensor is contiguous, fp32 data / int32 indices, row-major with the
 *     shapes given per function (same layouts as the reference).
 *   - outputs are caller-allocated and FULLY overwritten (the reference
 *     zero-initialises with torch::zeros, e.g. sampling.cpp:30-32; here the
 *     kernels write every element, including the zero rows of ball_query).
 *     *_grad entry points clear their output themselves (cudaMemsetAsync).
 *   - inputs are never modified.
 *   - asynchronous: work is enqueued on `stream`, no host synchronisation.
 *   - return value: 0 on success, a negative UPK_ERR_* for bad arguments, or a
 *     positive cudaError_t.  Never exit()s (the reference's
 *     CUDA_CHECK_ERRORS, _ext_src/include/cuda_utils.h:35-44, does).
 */
#ifndef UNOPOSE_B200_H
#define UNOPOSE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* upk_stream_t; /* == cudaStream_t */

#define UPK_OK 0
#define UPK_ERR_INVALID_ARG (-1)
#define UPK_ERR_UNSUPPORTED (-2)

/* Library/ABI version and the SM architecture the kernels were built for (100). */
int upk_abi_version(void);
int upk_built_sm(void);
/* Launch counter: number of kernel launches this library has enqueued in the
 * calling process since load (used by bench.py for "gpu_launches"). */
unsigned long long upk_launch_count(void);

/* ------------------------------------------------------------------------- *
 * (5) pointnet2 ops — replace the 9 functions of _ext_src/src/bindings.cpp
 * ------------------------------------------------------------------------- */

/* furthest_point_sampling(points[b,n,3], nsamples) -> idx[b,m] int32
 * replaces _ext.furthest_point_sampling (sampling.cpp:70-91,
 * sampling_gpu.cu:74-234).  Start index 0; bit-exact tie order of the
 * reference's block-size-dependent tree reduction.  The reference's (b,n)
 * `temp` scratch lives in registers here; no workspace needed. */
int upk_furthest_point_sampling(const float* xyz, int b, int n, int m,
                                int* idx_out, upk_stream_t stream);

/* gather_points(points[b,c,n], idx[b,m]) -> out[b,c,m]
 * replaces _ext.gather_points (sampling.cpp:20-43, sampling_gpu.cu:13-35). */
int upk_gather_points(const float* points, const int* idx, int b, int c, int n,
                      int m, float* out, upk_stream_t stream);

/* gather_points_grad(grad_out[b,c,m], idx[b,m], n) -> grad_points[b,c,n]
 * replaces _ext.gather_points_grad (sampling.cpp:45-68, sampling_gpu.cu:39-60). */
int upk_gather_points_grad(const float* grad_out, const int* idx, int b, int c,
                           int n, int m, float* grad_points, upk_stream_t stream);

/* ball_query(new_xyz[b,m,3], xyz[b,n,3], radius, nsample) -> idx[b,m,nsample]
 * replaces _ext.ball_query (ball_query.cpp:13-37, ball_query_gpu.cu:14-58).
 * First `nsample` hits in ascending point index with d2 < radius*radius,
 * remaining slots filled with the first hit, all-zero row if no hit. */
int upk_ball_query(const float* new_xyz, const float* xyz, int b, int n, int m,
                   float radius, int nsample, int* idx_out, upk_stream_t stream);

/* group_points(points[b,c,n], idx[b,npoints,nsample]) -> out[b,c,npoints,nsample]
 * replaces _ext.group_points (group_points.cpp:17-40, group_points_gpu.cu:13-44). */
int upk_group_points(const float* points, const int* idx, int b, int c, int n,
                     int npoints, int nsample, float* out, upk_stream_t stream);

/* group_points_grad(grad_out[b,c,npoints,nsample], idx, n) -> grad_points[b,c,n]
 * replaces _ext.group_points_grad (group_points.cpp:42-65, group_points_gpu.cu:48-80). */
int upk_group_points_grad(const float* grad_out, const int* idx, int b, int c,
                          int n, int npoints, int nsample, float* grad_points,
                          upk_stream_t stream);

/* three_nn(unknown[b,n,3], known[b,m,3]) -> dist2[b,n,3], idx[b,n,3]
 * replaces _ext.three_nn (interpolate.cpp:19-45, interpolate_gpu.cu:14-73).
 * dist2 is the SQUARED distance (the Python wrapper takes the sqrt). */
int upk_three_nn(const float* unknown, const float* known, int b, int n, int m,
                 float* dist2_out, int* idx_out, upk_stream_t stream);

/* three_interpolate(points[b,c,m], idx[b,n,3], weight[b,n,3]) -> out[b,c,n]
 * replaces _ext.three_interpolate (interpolate.cpp:47-74, interpolate_gpu.cu:77-116). */
int upk_three_interpolate(const float* points, const int* idx, const float* weight,
                          int b, int c, int m, int n, float* out,
                          upk_stream_t stream);

/* three_interpolate_grad(grad_out[b,c,n], idx, weight, m) -> grad_points[b,c,m]
 * replaces _ext.three_interpolate_grad (interpolate.cpp:76-104, interpolate_gpu.cu:121-160). */
int upk_three_interpolate_grad(const float* grad_out, const int* idx,
                               const float* weight, int b, int c, int n, int m,
                               float* grad_points, upk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* UNOPOSE_B200_H */
