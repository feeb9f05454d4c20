"""Dev tool: CUDA-event timings of individual ops at the bench shapes (B instances).

    python scripts/kbench.py [--b 16] [--only bqg,fps,...] [--once]

--once runs every selected op a single time after one warm-up (for `ncu`)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from unopose_b200 import model_utils as MU  # noqa: E402
from unopose_b200.pipeline import HotPathConfig, synthetic_inputs  # noqa: E402
from unopose_b200.pointnet2 import _ext as X  # noqa: E402


def timeit(fn, it=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=16)
    ap.add_argument("--only", default="")
    ap.add_argument("--once", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    cfg = HotPathConfig()
    inp = synthetic_inputs(1, a.b, cfg, device=dev)
    pts, tem = inp["pts"].contiguous(), inp["tem_pts"].contiguous()
    f3 = pts.transpose(1, 2).contiguous()
    idx256 = X.ball_query(pts, pts, 0.2, 256)
    c_att = MU.compute_feature_similarity(inp["c_f1"], inp["c_f2"], "cosine", cfg.temp, True)
    f_att = MU.compute_feature_similarity(inp["f_f1"], inp["f_f2"], "cosine", cfg.temp, True)
    f_att2, f_stats = MU.compute_feature_similarity(inp["f_f1"], inp["f_f2"], "cosine", cfg.temp, True, return_stats=True)
    ops = {
        "fps5000": lambda: X.furthest_point_sampling(tem, 2048),
        "fps2048": lambda: X.furthest_point_sampling(pts, 196),
        "bq64": lambda: X.ball_query(pts, pts, 0.1, 64),
        "bq256": lambda: X.ball_query(pts, pts, 0.2, 256),
        "group256": lambda: X.group_points(f3, idx256),
        "bqg2": lambda: X.ball_query_group(pts, pts, [(0.1, 64), (0.2, 256)]),
        "bqg1": lambda: X.ball_query_group(pts, pts, [(0.2, 256)]),
        "bqg2_noidx": lambda: X.ball_query_group(pts, pts, [(0.1, 64), (0.2, 256)], group=False),
        "csim": lambda: MU.compute_feature_similarity(inp["c_f1"], inp["c_f2"], "cosine", cfg.temp, True),
        "fsim": lambda: MU.compute_feature_similarity(inp["f_f1"], inp["f_f2"], "cosine", cfg.temp, True),
        "cpose": lambda: MU.compute_coarse_Rt_overlap(c_att, inp["c_score"], inp["c_pts1"], inp["c_pts2"], None,
                                                      cfg.n_proposal1, cfg.n_proposal2),
        "fpose": lambda: MU.compute_fine_Rt_overlap(f_att, inp["f_score"], inp["f_pts1"], inp["f_pts2"], None,
                                                    cfg.dis_thres),
        "fsim_stats": lambda: MU.compute_feature_similarity(inp["f_f1"], inp["f_f2"], "cosine", cfg.temp, True,
                                                            return_stats=True),
        "fpose_stats": lambda: MU.compute_fine_Rt_overlap(f_att2, inp["f_score"], inp["f_pts1"], inp["f_pts2"], None,
                                                          cfg.dis_thres, stats=f_stats),
    }
    sel = [s for s in a.only.split(",") if s] or list(ops)
    for name in sel:
        if a.once:
            ops[name]()
            torch.cuda.synchronize()
            ops[name]()
            torch.cuda.synchronize()
        else:
            print("B=%d %-12s %9.1f us" % (a.b, name, timeit(ops[name])), flush=True)


if __name__ == "__main__":
    main()
