"""Dev: GEMM-like ops of the product's matching forward at B=32 grouped by input shapes (torch profiler, record_shapes)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from baseline import refgpu  # noqa: E402
from unopose_b200.model import UNOPose  # noqa: E402
from unopose_b200.synthetic import forward_batch  # noqa: E402
from util_state import keyed_state_dict  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
model = UNOPose(refgpu.real_model_cfg()).eval()
model.load_state_dict(keyed_state_dict(model.state_dict(), 12))
model = model.to(dev)
inp = {k: torch.from_numpy(v).to(dev) for k, v in forward_batch(3, B).items()}
feed = lambda: {k: v for k, v in inp.items() if k not in ("R", "t")}  # noqa: E731
with torch.no_grad():
    feats = model.feature_extraction(feed())
    for _ in range(2):
        model.matching_forward(*feats, feed())
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
        model.matching_forward(*feats, feed())
        torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True)
        if e.key in ("aten::addmm", "aten::bmm", "aten::mm", "aten::conv1d", "aten::einsum", "aten::matmul", "aten::linear",
                     "aten::layer_norm", "aten::copy_", "aten::cat")]
rows.sort(key=lambda e: -e.device_time_total)
for e in rows[:40]:
    print("%-14s %9.1f us  x%-4d %s" % (e.key, e.device_time_total, e.count, str(e.input_shapes)[:150]))
