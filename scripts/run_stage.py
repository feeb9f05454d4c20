"""Dev tool: run one stage of the hot path a few times (for ncu captures).
usage: python scripts/run_stage.py <fine_pose|fine_sim|coarse|fps|pe|all> [batch] [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unopose_b200 import model_utils as MU  # noqa: E402
from unopose_b200.pipeline import HotPathConfig, run_hot_path, synthetic_inputs  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = HotPathConfig()
dev = torch.device("cuda:0")
inp = synthetic_inputs(0, B, cfg, device=dev)
plan = []
out = run_hot_path(inp, cfg, stages=plan)
names = {"fps": ["fps_template+gather", "fps_sparse+gather"], "coarse": ["coarse_similarity", "coarse_pose"],
         "pe": ["ball_query+group"], "fine_sim": ["fine_similarity"], "fine_pose": ["fine_similarity", "fine_pose"],
         "all": [n for n, _ in plan]}[what]
# prerequisites
for n, fn in plan:
    fn()
torch.cuda.synchronize()
for _ in range(reps):
    for n, fn in plan:
        if n in names:
            fn()
torch.cuda.synchronize()
print("done", what)
