"""One hot-path step (the workload of bench.py / unopose_b200.pipeline.run_hot_path) executed by the REFERENCE's own
functions: `ns` is the namespace returned by baseline.refgpu.load() — its unmodified model_utils / pointnet2_utils and,
on a GPU, its own `_ext` compiled for sm_100a.  Bench / test infrastructure; never imported by the product.

Same stages, same order, same inputs as run_hot_path (UNOPose.forward):
  sample_pts_feats x3 -> compute_feature_similarity -> compute_coarse_Rt_overlap -> ball_query + grouping_operation
  (both PositionalEncoding scales, both clouds; the query cloud moved by the coarse pose as the fine module does,
  oneref_predator_fine_point_matching.py:65-72) -> compute_feature_similarity -> compute_fine_Rt_overlap.
"""


def ref_step(ns, inp, cfg, record=None):
    mu, pu = ns.model_utils, ns.pointnet2_utils
    out = {}
    out["tem_sub"], tem_f, out["tem_idx"] = mu.sample_pts_feats(inp["tem_pts"], inp["tem_feats"], cfg.n_fine, True)
    out["sp1"], out["sf1"], out["fps_idx1"] = mu.sample_pts_feats(inp["pts"], inp["pts_feats"], cfg.n_coarse, True)
    out["sp2"], out["sf2"], out["fps_idx2"] = mu.sample_pts_feats(out["tem_sub"], tem_f, cfg.n_coarse, True)
    out["c_atten"] = mu.compute_feature_similarity(inp["c_f1"], inp["c_f2"], "cosine", cfg.temp, True)
    out["init_R"], out["init_t"], out["init_pose_score"] = mu.compute_coarse_Rt_overlap(
        out["c_atten"], inp["c_score"], inp["c_pts1"], inp["c_pts2"], None, cfg.n_proposal1, cfg.n_proposal2)
    out["pts_moved"] = (inp["pts"] - out["init_t"].unsqueeze(1)) @ out["init_R"]
    for name, cloud in (("q", out["pts_moved"]), ("r", out["tem_sub"])):
        cloud = cloud.contiguous()
        feats = cloud.transpose(1, 2).contiguous()
        for i, (r, nsample) in enumerate(cfg.pe):
            idx = pu.ball_query(r, nsample, cloud, cloud)
            out["pe_idx_%s%d" % (name, i)] = idx
            out["pe_%s%d" % (name, i)] = pu.grouping_operation(feats, idx)
    out["f_atten"] = mu.compute_feature_similarity(inp["f_f1"], inp["f_f2"], "cosine", cfg.temp, True)
    out["pred_R"], out["pred_t"], out["pred_pose_score"] = mu.compute_fine_Rt_overlap(
        out["f_atten"], inp["f_score"], inp["f_pts1"], inp["f_pts2"], None, cfg.dis_thres)
    return out


def ref_step_cpu(ns, inp, cfg, threads=None):
    """The same step on the host cores with the reference's own torch code.  Its three pointnet2 ops are CUDA-only
    (`CPU not supported`, sampling.cpp:39), so those calls are served by the C restatement oracle/pointnet2_oracle.c
    (pinned bit-exactly against the reference extension; instances spread over host threads); everything else is the
    reference's Python, unmodified."""
    import os
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    import torch

    from oracle import pointnet2_oracle as O

    mu, pu = ns.model_utils, ns.pointnet2_utils
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)

    def per_instance(fn, b):
        if threads <= 1 or b <= 1:
            return np.concatenate([fn(i) for i in range(b)])
        with ThreadPoolExecutor(max_workers=min(threads, b)) as ex:
            return np.concatenate(list(ex.map(fn, range(b))))

    def fps(pts, n):
        a = pts.contiguous().numpy()
        return torch.from_numpy(per_instance(lambda i: O.furthest_point_sampling(a[i:i + 1], n), a.shape[0]))

    def gather(feats, idx):
        a, j = feats.contiguous().numpy(), idx.numpy()
        return torch.from_numpy(per_instance(lambda i: O.gather_points(a[i:i + 1], j[i:i + 1]), a.shape[0]))

    def ball_query(r, nsample, xyz, new_xyz):
        a, q = xyz.contiguous().numpy(), new_xyz.contiguous().numpy()
        return torch.from_numpy(per_instance(lambda i: O.ball_query(q[i:i + 1], a[i:i + 1], r, nsample), a.shape[0]))

    def group(feats, idx):
        a, j = feats.contiguous().numpy(), idx.numpy()
        return torch.from_numpy(per_instance(lambda i: O.group_points(a[i:i + 1], j[i:i + 1]), a.shape[0]))

    saved = (mu.furthest_point_sample, mu.gather_operation, pu.ball_query, pu.grouping_operation)
    mu.furthest_point_sample, mu.gather_operation, pu.ball_query, pu.grouping_operation = fps, gather, ball_query, group
    try:
        return ref_step(ns, inp, cfg)
    finally:
        mu.furthest_point_sample, mu.gather_operation, pu.ball_query, pu.grouping_operation = saved
