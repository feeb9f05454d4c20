"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of one hot-path step (the same stage list as unopose_b200/pipeline.py), built
from the oracle pieces: the pose math is the torch-op restatement of the reference
(oracle/pose_oracle.py, run on CPU tensors with all host threads), the pointnet2 ops are the C
restatement (oracle/pointnet2_oracle.c — the reference has NO CPU path for them,
"CPU not supported", sampling.cpp:39; instances are spread over host threads).

Used by bench.py (`cpu_baseline` leg and `--impl reference`) and by __graft_entry__.smoke() as
the checker.  Never imported by the product package.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from . import pointnet2_oracle as O
from . import pose_oracle as PO


def _per_instance(fn, b, threads):
    if threads <= 1 or b <= 1:
        return [fn(i) for i in range(b)]
    with ThreadPoolExecutor(max_workers=min(threads, b)) as ex:
        return list(ex.map(fn, range(b)))


def _sample_pts_feats(pts, feats, npoint, threads):
    """model_utils.py:137-153 on numpy arrays: FPS + gathers (channel-first gather kernel semantics)."""
    b = pts.shape[0]

    def one(i):
        idx = O.furthest_point_sampling(pts[i:i + 1], npoint)
        p = O.gather_points(np.ascontiguousarray(pts[i:i + 1].transpose(0, 2, 1)), idx).transpose(0, 2, 1)
        f = O.gather_points(np.ascontiguousarray(feats[i:i + 1].transpose(0, 2, 1)), idx).transpose(0, 2, 1)
        return idx, p, f

    r = _per_instance(one, b, threads)
    return (np.concatenate([x[1] for x in r]), np.concatenate([x[2] for x in r]), np.concatenate([x[0] for x in r]))


def run_hot_path_cpu(inp, cfg, threads=None):
    """inp: dict of CPU torch tensors (unopose_b200.pipeline.synthetic_inputs(..., device=None))."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    npf = {k: v.numpy() for k, v in inp.items() if not k.startswith("_")}
    out = {}
    tem_sub, tem_sub_f, tem_idx = _sample_pts_feats(npf["tem_pts"], npf["tem_feats"], cfg.n_fine, threads)
    sp1, sf1, i1 = _sample_pts_feats(npf["pts"], npf["pts_feats"], cfg.n_coarse, threads)
    sp2, sf2, i2 = _sample_pts_feats(tem_sub, tem_sub_f, cfg.n_coarse, threads)
    out.update(tem_idx=tem_idx, fps_idx1=i1, fps_idx2=i2, tem_sub=tem_sub)
    c_atten = PO.feature_similarity(inp["c_f1"], inp["c_f2"], "cosine", cfg.temp, True)
    out["init_R"], out["init_t"], out["init_pose_score"] = PO.coarse_pose(
        c_atten, inp["c_score"], inp["c_pts1"], inp["c_pts2"], None, cfg.n_proposal1, cfg.n_proposal2)
    b = npf["pts"].shape[0]
    p1_ = ((inp["pts"] - out["init_t"].unsqueeze(1)) @ out["init_R"]).contiguous().numpy()   # fine module :65-72
    for name, cloud in (("q", p1_), ("r", tem_sub)):
        for k, (r, ns) in enumerate(cfg.pe):
            def one(i, cloud=cloud, r=r, ns=ns):
                idx = O.ball_query(cloud[i:i + 1], cloud[i:i + 1], r, ns)
                return O.group_points(np.ascontiguousarray(cloud[i:i + 1].transpose(0, 2, 1)), idx)
            out["pe_%s%d" % (name, k)] = np.concatenate(_per_instance(one, b, threads))
    f_atten = PO.feature_similarity(inp["f_f1"], inp["f_f2"], "cosine", cfg.temp, True)
    out["pred_R"], out["pred_t"], out["pred_pose_score"] = PO.fine_pose(
        f_atten, inp["f_score"], inp["f_pts1"], inp["f_pts2"], None, cfg.dis_thres)
    return out
