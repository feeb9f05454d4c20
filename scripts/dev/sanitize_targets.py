"""Small invocations of the kernels added in the second half of round 2, for compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/dev/sanitize_targets.py
(TMA-fed fine passes incl. L2 hints on 4 instances of 1024 x 1024, coarse solve with register-cached top-K + fused
selection, streamed top-K, fused ball query / grouping, FPS 8-warp configuration via UPK_FPS_CFG if set.)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from unopose_b200 import _lib as L  # noqa: E402
from unopose_b200 import model_utils as MU  # noqa: E402
from unopose_b200.pointnet2 import pointnet2_utils as P  # noqa: E402
from unopose_b200.synthetic import batch_clouds, matching_batch  # noqa: E402

dev = torch.device("cuda:0")
with torch.no_grad():
    mb = matching_batch(1, 4, 1024, 64)
    d = {k: torch.from_numpy(np.ascontiguousarray(mb[k])).to(dev) for k in ("f1", "f2", "pts1", "pts2", "score")}
    atten, stats = MU.compute_feature_similarity(d["f1"], d["f2"], "cosine", 0.1, True, return_stats=True)
    assert not atten.is_contiguous()
    os.environ.setdefault("UPK_FINE_L2_KEEP_MB", "6")       # 4 MB per instance here: two instances get evict_last
    R, t, s = MU.compute_fine_Rt_overlap(atten, d["score"], d["pts1"], d["pts2"], None, 0.15, stats=stats)
    R2, t2, s2 = MU.compute_fine_Rt_overlap(atten, d["score"], d["pts1"], d["pts2"], None, 0.15)
    mc = matching_batch(2, 3, 196, 64, kind="ball")
    c = {k: torch.from_numpy(np.ascontiguousarray(mc[k])).to(dev) for k in ("f1", "f2", "pts1", "pts2", "score")}
    ca = MU.compute_feature_similarity(c["f1"], c["f2"], "cosine", 0.1, True)
    Rc, tc, sc = MU.compute_coarse_Rt_overlap(ca, c["score"], c["pts1"], c["pts2"], None, 3000, 301)
    lib = L.load()
    v = torch.rand(2, 20000, device=dev)
    top = torch.empty((2, 300), dtype=torch.int32, device=dev)
    L.check(lib.upk_topk_smallest(L.ptr(v), 2, 20000, 300, L.ptr(top), L.stream_ptr(v)), "topk")
    cloud = torch.from_numpy(batch_clouds(3, 2, 2048, "surface")).to(dev)
    outs = P.ball_query_and_group(cloud, cloud, [(0.1, 64), (0.2, 256)])
    idx = P.furthest_point_sample(torch.from_numpy(batch_clouds(4, 2, 5000, "surface")).to(dev), 256)
    torch.cuda.synchronize()
print("sanitize targets ok", float(s.sum()), float(sc.sum()), int(top.sum()), int(idx.sum()))
