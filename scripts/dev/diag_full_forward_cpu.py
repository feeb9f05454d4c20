"""Dev diagnostic (this container only, needs /root/reference): the REFERENCE UNOPose.forward on CPU (timm stubbed, the
pointnet2 extension served by the C oracle) on one synthetic forward_batch instance; prints how degenerate the coarse
and fine assignments are (foreground label fractions, pose scores, error vs the planted pose)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_pose_golden import import_reference  # noqa: E402
from oracle import pointnet2_oracle as O  # noqa: E402
from oracle import pose_oracle as PO  # noqa: E402
from unopose_b200.synthetic import forward_batch  # noqa: E402
from util_state import keyed_state_dict  # noqa: E402

mu = import_reference()
import core.unopose.model.pointnet2.pointnet2_utils as pu  # noqa: E402
import core.unopose.model.transformer as tr  # noqa: E402

pu.ball_query = lambda radius, nsample, xyz, new_xyz: torch.from_numpy(O.ball_query(new_xyz.numpy(), xyz.numpy(), radius, nsample))
pu.grouping_operation = lambda f, idx: torch.from_numpy(O.group_points(f.detach().numpy(), idx.numpy()))
pu.gather_operation = lambda f, idx: torch.from_numpy(O.gather_points(f.detach().numpy(), idx.numpy()))
pu.furthest_point_sample = lambda xyz, n: torch.from_numpy(O.furthest_point_sampling(xyz.numpy(), n))
tr.gather_operation = pu.gather_operation
mu.gather_operation, mu.furthest_point_sample = pu.gather_operation, pu.furthest_point_sample

from baseline import refgpu  # noqa: E402

refgpu._stub_timm()
import importlib  # noqa: E402

model_mod = importlib.import_module("core.unopose.model.oneref_grf_predator_pose_estimation_model")
fine_mod = importlib.import_module("core.unopose.model.oneref_predator_fine_point_matching")
coarse_mod = importlib.import_module("core.unopose.model.oneref_predator_coarse_point_matching")

orig_fine, orig_coarse = mu.compute_fine_Rt_overlap, mu.compute_coarse_Rt_overlap


def spy_fine(atten, score, pts1, pts2, *a, **k):
    A = torch.softmax(atten, 2) * torch.softmax(atten, 1)
    l1 = A[:, 1:, :].max(2)[1]
    l2 = A[:, :, 1:].max(1)[1]
    print("  fine: logits range [%.2f, %.2f], fg rows %.3f, fg cols %.3f, score range [%.3f, %.3f]" % (
        float(atten.min()), float(atten.max()), float((l1 > 0).float().mean()), float((l2 > 0).float().mean()),
        float(score.min()), float(score.max())))
    return orig_fine(atten, score, pts1, pts2, *a, **k)


def spy_coarse(atten, score, pts1, pts2, *a, **k):
    A = torch.softmax(atten, 2) * torch.softmax(atten, 1)
    l1 = A[:, 1:, :].max(2)[1]
    print("  coarse: logits range [%.2f, %.2f], fg rows %.3f" % (float(atten.min()), float(atten.max()), float((l1 > 0).float().mean())))
    return orig_coarse(atten, score, pts1, pts2, *a, **k)


fine_mod.compute_fine_Rt_overlap = spy_fine
coarse_mod.compute_coarse_Rt_overlap = spy_coarse
torch.set_num_threads(8)
cfg = refgpu.real_model_cfg()
cfg["coarse_point_matching"]["nproposal1"] = 1000
if os.environ.get("TEMP"):
    cfg["coarse_point_matching"]["temp"] = cfg["fine_point_matching"]["temp"] = float(os.environ["TEMP"])
model = model_mod.UNOPose(cfg).eval()
model.load_state_dict(keyed_state_dict(model.state_dict(), 12))
d = forward_batch(int(sys.argv[1]) if len(sys.argv) > 1 else 3, 1, view_jitter_deg=float(os.environ.get("JIT", 3.0)))
inp = {k: torch.from_numpy(v) for k, v in d.items()}
with torch.no_grad():
    out = model({k: v for k, v in inp.items() if k not in ("R", "t")})
print("coarse err %.2f deg, final err %.2f deg, score %.4f" % (
    float(PO.rotation_geodesic_deg(out["init_R"], inp["R"])), float(PO.rotation_geodesic_deg(out["pred_R"], inp["R"])),
    float(out["pred_pose_score"])))
