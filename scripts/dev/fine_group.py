"""Dev experiment: does running the fine stage (similarity GEMM -> labels -> rows -> Kabsch -> inliers) in groups of G
instances keep the group's `atten` L2-resident between the passes?  Times one B = 16 batch as 16/G groups inside one
CUDA graph, for G in {16, 8, 4, 2}, rotating over input sets so nothing survives in L2 between replays.
usage: python scripts/dev/fine_group.py [B] [reps]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from unopose_b200 import model_utils as MU  # noqa: E402
from unopose_b200.pipeline import HotPathConfig, synthetic_inputs  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
only = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [16, 8, 4, 2]
cfg = HotPathConfig()
dev = torch.device("cuda:0")
sets = [synthetic_inputs(s, B, cfg, device=dev) for s in range(3)]


def fine(inp, G):
    res = []
    for g0 in range(0, B, G):
        sl = slice(g0, g0 + G)
        atten, stats = MU.compute_feature_similarity(inp["f_f1"][sl], inp["f_f2"][sl], "cosine", cfg.temp, True,
                                                     return_stats=True)
        res.append(MU.compute_fine_Rt_overlap(atten, inp["f_score"][sl], inp["f_pts1"][sl], inp["f_pts2"][sl], None,
                                              cfg.dis_thres, stats=stats))
    return [torch.cat([r[k] for r in res]) for k in range(3)]


out = {}
ref = None
with torch.no_grad():
    for G in only:
        graphs = []
        for inp in sets:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    r = fine(inp, G)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                r = fine(inp, G)
            graphs.append((g, r))
        for g, _ in graphs:
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            graphs[i % len(graphs)][0].replay()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        r0 = [t.clone() for t in graphs[0][1]]
        if ref is None:
            ref = r0
        same = all(torch.equal(a, b) for a, b in zip(ref, r0))
        out["G=%d" % G] = {"ms_per_batch": ms, "bit_identical_to_first": same}
        print("G=%2d  %.4f ms per %d-instance batch  identical=%s" % (G, ms, B, same), flush=True)
        del graphs
print(json.dumps(out))
