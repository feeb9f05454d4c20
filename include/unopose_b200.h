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

#include <stddef.h>

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

/* Fused ball_query + group_points of the xyz channels, for ONE or TWO (radius, nsample) scales in one
 * scan of the cloud: what QueryAndGroup / QueryAndLRFGroup (pointnet2_utils.py:292-378, :484-584) and the
 * two-scale PositionalEncoding (oneref_predator_fine_point_matching.py:159-178) compute with
 * ball_query -> transpose -> grouping_operation per scale.
 *   idxK[b,m,nsampleK] int32     == upk_ball_query(new_xyz, xyz, radiusK, nsampleK)          (bit-exact)
 *   groupedK[b,3,m,nsampleK]     == upk_group_points(xyz^T[b,3,n], idxK)   (may be NULL: indices only)
 * A scale with nsampleK == 0 is skipped (single-scale call). */
int upk_ball_query_group(const float* new_xyz, const float* xyz, int b, int n, int m,
                         float radius0, int nsample0, int* idx0, float* grouped0,
                         float radius1, int nsample1, int* idx1, float* grouped1,
                         upk_stream_t stream);

/* group_points(points[b,c,n], idx[b,npoints,nsample]) -> out[b,c,npoints,nsample]
 * replaces _ext.group_points (group_points.cpp:17-40, group_points_gpu.cu:13-44). */
int upk_group_points(const float* points, const int* idx, int b, int c, int n,
                     int npoints, int nsample, float* out, upk_stream_t stream);

/* group_points_grad(grad_out[b,c,npoints,nsample], idx, n) -> grad_points[b,c,n]
 * replaces _ext.group_points_grad (group_points.cpp:42-65, group_points_gpu.cu:48-80). */
int upk_group_points_grad(const float* grad_out, const int* idx, int b, int c,
                          int n, int npoints, int nsample, float* grad_points,
                          upk_stream_t stream);

/* Row gather, channel-last: x[b,n,c], idx[b,m] -> out[b,m,c] (and its scatter-add gradient).
 * Replaces the transpose -> _ext.gather_points -> transpose sequences of
 * sample_pts_feats / sample_pts_feats_wlrf / gather_pts_feats* (model_utils.py:137-212). */
int upk_gather_rows(const float* x, const int* idx, int b, int n, int m, int c, float* out,
                    upk_stream_t stream);
int upk_gather_rows_grad(const float* grad_out, const int* idx, int b, int n, int m, int c,
                         float* grad_x, upk_stream_t stream);

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

/* ------------------------------------------------------------------------- *
 * (1) feature similarity — compute_feature_similarity, model_utils.py:260-282
 * ------------------------------------------------------------------------- */

/* feat1[b,n,c], feat2[b,m,c] -> atten[b,n,m].
 * normalize != 0: rows L2-normalised first (F.normalize, eps 1e-12) — needs the
 * workspace (two normalised copies).  sim_type 0 = "cosine": f1.f2^T / temp;
 * 1 = "L2": sqrt(clamp(2 - 2 f1.f2^T, 0)) / temp. */
size_t upk_feature_similarity_workspace_bytes(int b, int n, int m, int c, int normalize);
/* Arithmetic of the tensor-core path (n*m >= 128^2, c % 16 == 0): 16 (default) = tcgen05 tensor cores with a
 * three-product split of fp32 operands, fp32-level accuracy either way: 3xFP16 (operands scaled by 2^12) for
 * normalised cosine logits on the CTA-pair shapes (enough 256x256 tiles to fill the GPU, c % 32 == 0), 3xTF32
 * elsewhere; 3 = 3xTF32 everywhere; 1 = single TF32 pass; 0 = fp32 SIMT everywhere.
 * Also settable with the environment variable UPK_SIMILARITY_MODE.  Returns the previous mode.
 * (The fused-statistics variant below has its own, larger threshold: the Python wrapper requests it above 512^2
 * logits, where the fine solver runs its streaming passes.) */
int upk_set_similarity_mode(int mode);
int upk_feature_similarity(const float* feat1, const float* feat2, int b, int n, int m,
                           int c, float temp, int normalize, int sim_type,
                           void* workspace, size_t workspace_bytes, float* atten_out,
                           upk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (2)(3)(4) coarse pose — compute_coarse_Rt[_overlap], model_utils.py:336-490
 * ------------------------------------------------------------------------- */

/* Optional device buffers receiving the intermediates (any member may be NULL;
 * pass dbg == NULL for none).  Used by the stage-wise parity tests. */
typedef struct upk_coarse_debug {
  float* w1;     /* [b,n1]   foreground mask of the query points (label1 > 0)   */
  float* w2;     /* [b,n2]                                                       */
  float* cdf;    /* [b,n1*n2] normalised sampling CDF                            */
  int* idx1;     /* [b,n_hyp,3] sampled query indices  (both or neither)         */
  int* idx2;     /* [b,n_hyp,3] sampled reference indices                        */
  float* Rs;     /* [b,n_hyp,9] */
  float* ts;     /* [b,n_hyp,3] */
  float* resid;  /* [b,n_hyp]   mean triplet residual                            */
  int* top;      /* [b,n_keep]  pool indices of the n_keep smallest residuals, ascending index */
  float* scores; /* [b,n_keep]  */
} upk_coarse_debug;

/* atten[b,n1+1,n2+1] (background token at row/col 0); score1[b,n1] / score2[b,n2]
 * with row strides score*_ld (both NULL => compute_coarse_Rt, no overlap scores);
 * pts1[b,n1,3] query, pts2[b,n2,3] reference, model_pts[b,n_model,3] or NULL (= pts2);
 * u[b,3*n_hyp] uniform draws in [0,1) (the reference's torch.rand, :462 — drawn by
 * the HOST so the Philox stream position is unchanged).
 * Outputs R[b,9] row-major, t[b,3], score[b], pool_idx[b] (index in [0,n_hyp) of the
 * selected hypothesis; may be NULL).  Pose convention: p_ref ~= (p_query - t) @ R. */
size_t upk_coarse_pose_workspace_bytes(int b, int n1, int n2, int n_hyp, int n_keep);
int upk_coarse_pose(const float* atten, const float* score1, int score1_ld,
                    const float* score2, int score2_ld, const float* pts1,
                    const float* pts2, const float* model_pts, int n_model,
                    const float* u, int b, int n1, int n2, int n_hyp, int n_keep,
                    void* workspace, size_t workspace_bytes, float* R_out, float* t_out,
                    float* score_out, int* pool_idx_out, const upk_coarse_debug* dbg,
                    upk_stream_t stream);

/* Stage-wise entry points (same kernels; used for identical-input parity tests
 * and for hypothesis sharding across GPUs). */

/* a2 + first half of a3 alone: masks w1[b,n1], w2[b,n2] and the normalised sampling CDF
 * cdf[b,n1*n2] (model_utils.py:443-461). */
size_t upk_coarse_assignment_workspace_bytes(int b, int n1, int n2);
int upk_coarse_assignment(const float* atten, const float* score1, int score1_ld,
                          const float* score2, int score2_ld, int b, int n1, int n2,
                          void* workspace, size_t workspace_bytes, float* w1_out,
                          float* w2_out, float* cdf_out, upk_stream_t stream);

/* searchsorted(cdf,u) -> triplets -> Kabsch -> residual for hypotheses
 * [h_begin,h_end) of every instance (model_utils.py:462-475). Rs/ts/resid are
 * indexed by the global hypothesis index ([b,n_hyp,...]). */
/* Profiling aid: upk_coarse_assignment's cluster kernel, additionally recording the SM clock of instance 0 / cluster
 * rank 0 at kernel entry and after each of its 8 cluster barriers into stamps_out[9] (device memory, int64). */
int upk_coarse_assignment_profile(const float* atten, const float* score1, int score1_ld, const float* score2,
                                  int score2_ld, int b, int n1, int n2, float* w1_out, float* w2_out,
                                  float* cdf_out, long long* stamps_out, upk_stream_t stream);

int upk_sample_hypotheses(const float* cdf, const float* u, const float* pts1,
                          const float* pts2, int b, int n1, int n2, int n_hyp,
                          int h_begin, int h_end, int* idx1_out, int* idx2_out,
                          float* Rs, float* ts, float* resid, upk_stream_t stream);
/* Kabsch on n explicit triplets: p1[n,3,3] (query), p2[n,3,3] (reference)
 * == WeightedProcrustes()(p2, p1, None), model_utils.py:469. resid may be NULL. */
int upk_kabsch_triplets(const float* p1, const float* p2, int n, float* Rs, float* ts,
                        float* resid, upk_stream_t stream);
/* indices of the k smallest of vals[b,n], ascending index order (torch.topk largest=False, :476) */
int upk_topk_smallest(const float* vals, int b, int n, int k, int* idx_out,
                      upk_stream_t stream);
/* scores[b,n_keep] for kept hypotheses [k_begin,k_end) (model_utils.py:481-485);
 * top == NULL means hypothesis k is pool index k. */
int upk_score_hypotheses(const float* pts1, const float* model_pts, const float* w1,
                         const float* Rs, const float* ts, const int* top, int b, int n1,
                         int n_model, int n_hyp, int n_keep, int k_begin, int k_end,
                         float* scores, upk_stream_t stream);
/* first arg-max over scores[b,n_keep] and gather (model_utils.py:486-488) */
int upk_select_best(const float* scores, const int* top, const float* Rs, const float* ts,
                    int b, int n_hyp, int n_keep, float* R_out, float* t_out,
                    float* score_out, int* pool_idx_out, upk_stream_t stream);

/* Hypothesis sharding across GPUs (SURVEY.md §8e partitioning B; no counterpart in the reference, which has no
 * collective on this path).  Ranks that share an instance batch each sample a slice [h_begin, h_end) of the pool
 * (upk_sample_hypotheses), keep the n_local smallest residuals of the slice (upk_topk_smallest on the slice) and
 * exchange fixed-size candidate lists: upk_pack_candidates writes n_slots records of 14 floats per instance
 * {residual, pool index (int bits), R[9], t[3]}, padded with {+inf, -1}; after an all_gather into
 * gathered[world][b][n_slots][14], upk_unpack_candidates scatters the OTHER ranks' records into the dense pool arrays
 * (resid[b][n_hyp], Rs, ts), on which the same upk_topk_smallest re-selects the global top-K.  upk_fill_f32 resets
 * resid to +inf (and a score buffer to -inf) before a solve. */
int upk_fill_f32(float* p, size_t n, float value, upk_stream_t stream);
int upk_pack_candidates(const float* resid, const float* Rs, const float* ts, const int* top_local, int b, int n_hyp,
                        int h_begin, int n_local, int n_slots, float* cand_out, upk_stream_t stream);
int upk_unpack_candidates(const float* gathered, int world, int my_rank, int b, int n_hyp, int n_slots, float* resid,
                          float* Rs, float* ts, upk_stream_t stream);
/* Compact merge (round 2): the gathered lists as a pool of world * n_slots candidates per instance —
 * resid_c[b][world*n_slots] (padding records get the largest key), Rs_c [..][9], ts_c [..][3], pool_c = pool index of
 * each entry (-1 for padding).  Compact order == pool-index order, so upk_topk_smallest / upk_score_hypotheses /
 * upk_select_best_map on these arrays (n_hyp = world * n_slots) select exactly what the dense merge selects, without
 * touching H-sized arrays.  upk_topk_smallest_ld is upk_topk_smallest on rows with a pitch (a slice of the pool). */
int upk_unpack_candidates_compact(const float* gathered, int world, int b, int n_slots, float* resid_c, float* Rs_c,
                                  float* ts_c, int* pool_c, upk_stream_t stream);
int upk_topk_smallest_ld(const float* vals, int b, int n, int ld, int k, int* idx_out, upk_stream_t stream);
int upk_select_best_map(const float* scores, const int* top, const float* Rs, const float* ts, const int* pool_map, int b,
                        int n_hyp, int n_keep, float* R_out, float* t_out, float* score_out, int* pool_idx_out,
                        upk_stream_t stream);

/* Relative-position score term of the geometric self-attention (core/unopose/model/transformer.py:392-395 with the
 * queries projected by W_p, see modules/transformer.py): out[b][h][n][m] = sum_c embed[b][n][m][c] * q2[b][n][c][h].
 * One pass over the (B,N,M,C) embedding at HBM speed; heads == 4, c in {128, 256}; UPK_ERR_UNSUPPORTED otherwise. */
int upk_rpe_scores(const float* embed, const float* q2, int b, int n, int m, int c, int heads, float* out,
                   upk_stream_t stream);

/* Fully-connected layer on the tensor cores at fp32-level accuracy: y[rows][out] = x[rows][in] W[out][in]^T + bias[out]
 * (bias may be NULL), optional ReLU — torch.nn.functional.linear as the matching modules' transformer blocks call it
 * (core/unopose/model/transformer.py:94-201,517-612: proj_q/k/v/p, linear, expand, squeeze; in_proj / out_proj of the
 * matching modules).  The reference runs these as fp32 SIMT SGEMMs (TF32 is off, main_unopose.py:139-141); this is the
 * 3xTF32 tcgen05 GEMM of upk_feature_similarity with a bias / ReLU epilogue.  in_features % 16 == 0, 16-byte aligned
 * operands; UPK_ERR_UNSUPPORTED otherwise (callers then use their own GEMM). */
size_t upk_linear_workspace_bytes(int rows, int in_features, int out_features);
int upk_linear(const float* x, const float* weight, const float* bias, int rows, int in_features, int out_features,
               int relu, void* workspace, size_t workspace_bytes, float* y, upk_stream_t stream);

/* Peer exchange (NVLink / NVSwitch peer memory) — the fused compute + collective form of the two exchange steps of
 * hypothesis sharding and of the result gather of instance sharding (BASELINE.json north_star: "an NCCL
 * gather over NVLink of per-shard best scores").  No counterpart in the reference (single process, no collective).
 * The host maps one SYMMETRIC allocation per rank into every process (torch.distributed._symmetric_memory, CUDA VMM)
 * and describes it with upk_peer_t; producer kernels store their results into every rank's copy and publish an epoch
 * flag, consumer kernels wait on their local flags (csrc/peer.cuh has the protocol).  Channels are independent
 * exchange steps; data of epoch e lives in slab (e & 1) of its channel. */
#define UPK_MAX_PEERS 8
#define UPK_PEER_CHANNELS 4
typedef struct upk_peer {
  int world, rank;
  void* data[UPK_MAX_PEERS];                /* rank r's exchange buffer as mapped in THIS process (data[rank] = local) */
  unsigned long long* flags[UPK_MAX_PEERS]; /* rank r's flag array [UPK_PEER_CHANNELS][UPK_MAX_PEERS], zero-initialised */
  unsigned long long* epoch;                /* local: [UPK_PEER_CHANNELS], zero-initialised                            */
  unsigned int* done;                       /* local: [UPK_PEER_CHANNELS] CTA-completion counters, zero-initialised     */
  unsigned int* status;                     /* local: [1], set to 1 if a wait timed out (a peer never published)        */
} upk_peer_t;

/* upk_pack_candidates that writes this rank's candidate list into EVERY rank's gathered[world][b][n_slots][14] array
 * (at data[r] + data_offset + slab * slab_bytes) and publishes channel `channel`. */
int upk_pack_candidates_peer(const float* resid, const float* Rs, const float* ts, const int* top_local, int b, int n_hyp,
                             int h_begin, int n_local, int n_slots, const upk_peer_t* peer, size_t data_offset,
                             size_t slab_bytes, int channel, upk_stream_t stream);
/* upk_unpack_candidates that first waits for every rank's publication on `channel`, then reads the local array. */
int upk_unpack_candidates_peer(const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel, int b, int n_hyp,
                               int n_slots, float* resid, float* Rs, float* ts, upk_stream_t stream);
/* upk_score_hypotheses whose scores[b][n_keep] entries [k_begin,k_end) are stored into every rank's score table. */
int upk_score_hypotheses_peer(const float* pts1, const float* model_pts, const float* w1, const float* Rs, const float* ts,
                              const int* top, int b, int n1, int n_model, int n_hyp, int n_keep, int k_begin, int k_end,
                              const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel,
                              upk_stream_t stream);
/* upk_unpack_candidates_compact after the same wait. */
int upk_unpack_candidates_compact_peer(const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel, int b,
                                       int n_slots, float* resid_c, float* Rs_c, float* ts_c, int* pool_c,
                                       upk_stream_t stream);
/* upk_select_best[_map] on the local score table after waiting for every rank's slice (pool_map may be NULL). */
int upk_select_best_peer(const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel, const int* top,
                         const float* Rs, const float* ts, const int* pool_map, int b, int n_hyp, int n_keep, float* R_out,
                         float* t_out, float* score_out, int* pool_idx_out, upk_stream_t stream);
/* Generic all-gather of `bytes` (multiple of 16) per rank: src -> every rank's data[r] + data_offset + slab * slab_bytes
 * + rank * bytes, then publish; upk_peer_wait blocks the stream until every rank has published and (optionally) copies
 * the gathered world * bytes out of the slab into dst (may be NULL). */
int upk_peer_all_gather(const void* src, size_t bytes, const upk_peer_t* peer, size_t data_offset, size_t slab_bytes,
                        int channel, upk_stream_t stream);
int upk_peer_wait(const upk_peer_t* peer, size_t data_offset, size_t slab_bytes, int channel, void* dst, size_t bytes,
                  upk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * fine pose — compute_fine_Rt[_overlap], model_utils.py:493-566
 * ------------------------------------------------------------------------- */
typedef struct upk_fine_debug {
  float* w1;   /* [b,n1] */
  float* w2;   /* [b,n2] */
  float* soft; /* [b,n1,3] soft correspondences (assignment-weighted reference points) */
  float* asum; /* [b,n1]   row sums of the masked assignment (Procrustes weights)       */
  float* nn;   /* [b,n1]   nearest-neighbour distance of each transformed query point   */
} upk_fine_debug;

/* weight_thresh: 0.001 for the _overlap variant (:528), 0.0 for compute_fine_Rt (:502). */
size_t upk_fine_pose_workspace_bytes(int b, int n1, int n2);
int upk_fine_pose(const float* atten, const float* score1, int score1_ld,
                  const float* score2, int score2_ld, const float* pts1, const float* pts2,
                  const float* model_pts, int n_model, int b, int n1, int n2,
                  float dis_thres, float weight_thresh, void* workspace,
                  size_t workspace_bytes, float* R_out, float* t_out, float* score_out,
                  const upk_fine_debug* dbg, upk_stream_t stream);

/* Fused pass 1 (cosine logits): the similarity GEMM's epilogue also emits the exponent sums that the dual-softmax
 * assignment of compute_fine_Rt[_overlap] needs (model_utils.py:542), so the fine solve reads `atten` twice
 * instead of three times.  |cosine| <= 1 bounds every logit by 1/temp: one fixed reference exponent, no maxima.
 *   upk_feature_similarity_stats == upk_feature_similarity(normalize = 1, sim_type = 0) + stats_out
 *   upk_fine_pose_stats          == upk_fine_pose, given the stats of the SAME atten and temp
 * Returns UPK_ERR_UNSUPPORTED when the tensor-core path does not apply (small shapes, c % 16 != 0, unaligned
 * features, similarity mode != 3): callers then use the plain pair. */
size_t upk_similarity_stats_bytes(int b, int n, int m);
int upk_feature_similarity_stats(const float* feat1, const float* feat2, int b, int n, int m, int c, float temp,
                                 void* workspace, size_t workspace_bytes, float* atten_out, float* stats_out,
                                 size_t stats_bytes, upk_stream_t stream);
int upk_fine_pose_stats(const float* atten, const float* stats, size_t stats_bytes, float temp,
                        const float* score1, int score1_ld, const float* score2, int score2_ld,
                        const float* pts1, const float* pts2, const float* model_pts, int n_model, int b, int n1,
                        int n2, float dis_thres, float weight_thresh, void* workspace, size_t workspace_bytes,
                        float* R_out, float* t_out, float* score_out, const upk_fine_debug* dbg,
                        upk_stream_t stream);

/* Pitched variants.  atten[b][i][j] lives at atten + (b * n + i) * atten_ld + j (atten_ld >= m floats per row).  With
 * atten_ld % 4 == 0 and element (0, 1, 1) 16-byte aligned — the layout unopose_b200.model_utils.compute_feature_similarity
 * allocates for the fine shape: 3 pad floats in front of every row of 2049 — the GEMM stores its tiles with TMA tensor
 * stores and the assignment passes read them with 128-bit loads; any other pitch takes the 4-byte paths.
 *   upk_feature_similarity_stats_ld: stats_out may be NULL (logits only).  UPK_ERR_UNSUPPORTED when the tensor-core path
 *     does not apply, or (with stats_out) when temp is so small that the single-reference exponent sums would underflow
 *     (2 log2e / temp > 60): callers then use upk_feature_similarity + upk_fine_pose.
 *   upk_fine_pose_ld: stats may be NULL (three-pass solve). */
int upk_feature_similarity_stats_ld(const float* feat1, const float* feat2, int b, int n, int m, int c, float temp,
                                    void* workspace, size_t workspace_bytes, float* atten_out, int atten_ld,
                                    float* stats_out, size_t stats_bytes, upk_stream_t stream);
int upk_fine_pose_ld(const float* atten, int atten_ld, const float* stats, size_t stats_bytes, float temp,
                     const float* score1, int score1_ld, const float* score2, int score2_ld,
                     const float* pts1, const float* pts2, const float* model_pts, int n_model, int b, int n1,
                     int n2, float dis_thres, float weight_thresh, void* workspace, size_t workspace_bytes,
                     float* R_out, float* t_out, float* score_out, const upk_fine_debug* dbg,
                     upk_stream_t stream);

/* weighted_procrustes(src[b,n,3], ref[b,n,3], weights[b,n] or NULL, thresh, eps)
 * -> R[b,9], t[b,3] with ref ~= R src + t  (model_utils.py:667-743). */
int upk_weighted_procrustes(const float* src, const float* ref, const float* weights, int b,
                            int n, float weight_thresh, float eps, float* R_out,
                            float* t_out, upk_stream_t stream);

/* Global local-reference-frame coordinates of a cloud: `get_batch_lrf(pts)` of the reference models
 * (oneref_grf_predator_pose_estimation_model.py:78-93), i.e. LRF(r)(centroid, pts) of model_utils.py:766-823.
 * pts[b,n,3] -> out[b,n,3].  radius[b] or NULL (then r = max |p - centroid|, the `use_ref_rad = False` branch; pass
 * ones for `use_ref_rad = True`).  frame_out (optional, [b,13]): the frame columns x|y|z row-major (9), the centre (3)
 * and the radius (1). */
int upk_global_lrf(const float* pts, const float* radius, int b, int n, float eps, float* out, float* frame_out,
                   upk_stream_t stream);

/* QueryAndLRFGroup's frame stage in one kernel: LRF_batch (pointnet2_utils.py:429-481) + the feature assembly of
 * QueryAndLRFGroup.forward (:556-571).  grouped[b,3,n,ns] are the ABSOLUTE neighbour coordinates (grouping_operation /
 * upk_ball_query_group output), centres[b,n,3] the points the frames are anchored at (`xyz`), new_xyz[b,n,3] the points
 * subtracted for the raw offsets.  out[b,6,n,ns] = [grouped - new_xyz (/ r_lrf if normalize_xyz) | frame coordinates / r_lrf]
 * when use_xyz, else out[b,3,n,ns] = frame coordinates.  Replaces a cuSOLVER batched SVD + ~15 elementwise passes; when
 * the +-1e-3 sign vote of a frame's z axis is exactly 0 the reference's z sign is whatever its SVD returned — here it is
 * the Jacobi solver's, deterministically. */
int upk_lrf_group(const float* centres, const float* new_xyz, const float* grouped, int b, int n, int ns,
                  float r_lrf, float eps, int use_xyz, int normalize_xyz, float* out, upk_stream_t stream);

/* out[b,i,:] = (pts[b,i,:] - t[b]) @ R[b]: a cloud moved by a pose, `p1_ = (p1 - init_t) @ init_R` of the fine
 * module (oneref_predator_fine_point_matching.py:65-72) and the scoring transforms of model_utils.py:483,:558. */
int upk_transform_points(const float* pts, const float* R, const float* t, int b, int n, float* out,
                         upk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (f2) GeometricStructureEmbedding — replaces the torch-op sequence of
 *      core/unopose/model/transformer.py:287-350 (+ SinusoidalPositionalEmbedding, :261-284)
 * ------------------------------------------------------------------------- */

/* 1 if the fused kernel handles (hidden_dim c, angle_k): c a multiple of 32 in [32,256], 1 <= angle_k <= 8. */
int upk_geometric_embedding_supported(int c, int angle_k);
size_t upk_geometric_embedding_workspace_bytes(int b, int n, int c, int angle_k);

/* get_embedding_indices(points[b,n,3]) -> d_idx[b,n,n], a_idx[b,n,n,angle_k]   (transformer.py:303-336)
 * d = sqrt(clamp(x2 - 2xy + y2, 0)) / sigma_d (the expansion form of pairwise_distance, model_utils.py:230-257);
 * neighbours = the angle_k nearest after the nearest (ties: lower index); angle = atan2(|ref x anc|, ref.anc) * factor_a. */
int upk_geometric_embedding_indices(const float* points, int b, int n, int angle_k, float sigma_d, float factor_a,
                                    float* d_idx, float* a_idx, upk_stream_t stream);

/* forward(points[b,n,3]) -> out[b,n,n,c]   (transformer.py:338-350)
 * out = proj_d(emb(d_idx)) + red_k proj_a(emb(a_idx)), red = max (reduction_mean = 0) or mean (1).
 * div_term[c/2] is the module's `embedding.div_term` buffer; w_d/w_a [c,c] row-major (out, in) and b_d/b_a [c] are
 * `proj_d` / `proj_a`.  The sinusoid rows are generated in shared memory as the A operand of tcgen05.mma (3xTF32);
 * only `out` is written (the "a" phase adds into what the "d" phase stored). */
int upk_geometric_embedding(const float* points, int b, int n, int c, int angle_k, float sigma_d, float factor_a,
                            const float* div_term, const float* w_d, const float* b_d, const float* w_a,
                            const float* b_a, int reduction_mean, void* workspace, size_t workspace_bytes,
                            float* out, upk_stream_t stream);

/* ------------------------------------------------------------------------- *
 * (f1) SharedMLP + max over the ball — the MLP half of PositionalEncoding
 *      (oneref_predator_fine_point_matching.py:167-176: `self.mlp1(group(...)).max(dim=3)[0]`;
 *       SharedMLP: pointnet2/pytorch_utils.py:25-48, three 1x1 Conv2d + BatchNorm2d + ReLU)
 * ------------------------------------------------------------------------- */

/* 1 if the fused kernel handles the geometry: cin <= 16, c1 in {16,32}, c2 in {32,64}, c3 in {32,...,128 step 32},
 * nsample >= 32 with nsample % 128 == 0 or 128 % nsample == 0, (m * nsample) % 128 == 0. */
int upk_shared_mlp_max_supported(int cin, int c1, int c2, int c3, int m, int nsample);

/* x[b,cin,m,nsample] -> out[b,c3,m] = max_s relu(W3 relu(W2 relu(W1 x + b1) + b2) + b3).
 * W_l [c_l, c_{l-1}] row-major and b_l [c_l] are the conv weights with the eval-mode batch norm FOLDED IN by the caller
 * (W' = W * gamma / sqrt(var + eps) per output row, b' = beta - mean * gamma / sqrt(var + eps)).  The three GEMMs run
 * on tcgen05 (3xTF32) with the activations kept in shared memory / TMEM between layers. */
int upk_shared_mlp_max(const float* x, int b, int cin, int m, int nsample, int c1, int c2, int c3,
                       const float* w1, const float* b1, const float* w2, const float* b2, const float* w3,
                       const float* b3, float* out, upk_stream_t stream);

/* HOST function (no GPU): the 3x3 Procrustes rotation solver of kernel family (3),
 * compiled from the same source as the device code.  H[n,9] row-major -> R[n,9]. */
int upk_host_procrustes_rotation(const double* H, int n, double* R_out);

/* HOST function (no GPU): the z axis (least-variance eigenvector, raw sign) of the local-reference-frame kernels
 * (upk_lrf_group, upk_global_lrf), compiled from the same source as the device code.  The raw sign is the one
 * cuSOLVER's batched SVD returns for the last column of V (the reference keeps it when the sign vote ties,
 * core/unopose/model/pointnet2/pointnet2_utils.py:451-456).  cov[n,6] = (xx, yy, zz, xy, xz, yz) -> z[n,3]. */
int upk_host_lrf_z_axis(const double* cov, int n, double* z_out);

#ifdef __cplusplus
}
#endif
#endif /* UNOPOSE_B200_H */
