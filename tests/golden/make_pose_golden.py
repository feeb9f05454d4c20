"""Generate golden vectors for the pose path by IMPORTING THE REFERENCE ITSELF
(/root/reference, read-only) and running its own functions on CPU tensors.

    python tests/golden/make_pose_golden.py      (only works where /root/reference exists)

Writes tests/golden/pose_*.npz (inputs + the reference's outputs).  The import
shims are the ones SURVEY.md Appendix B verified: the pointnet2 extension is
not needed by these functions (builtins.__POINTNET2_SETUP__) and detectron2's
logger (pulled in by loss_utils) is stubbed.
"""
import builtins
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from unopose_b200.synthetic import matching_batch  # noqa: E402


def import_reference():
    sys.path.insert(0, "/root/reference")
    builtins.__POINTNET2_SETUP__ = True
    d2l = types.ModuleType("detectron2.utils.logger")
    d2l.log_first_n = d2l.log_every_n = lambda *a, **k: None
    sys.modules.update({"detectron2": types.ModuleType("detectron2"),
                        "detectron2.utils": types.ModuleType("detectron2.utils"),
                        "detectron2.utils.logger": d2l})
    from core.unopose.utils import model_utils

    return model_utils


def main():
    mu = import_reference()
    out = os.path.join(ROOT, "tests", "golden")
    torch.set_num_threads(1)
    # ---- coarse: 196 x 196, H hypotheses, K kept
    for name, seed, n, H, K in (("coarse_a", 101, 196, 900, 60), ("coarse_b", 102, 196, 5000, 300),
                                ("coarse_c", 103, 64, 300, 40)):
        d = matching_batch(seed, 2, n, 64 if n == 64 else 256)
        f1, f2 = torch.from_numpy(d["f1"]), torch.from_numpy(d["f2"])
        atten = mu.compute_feature_similarity(f1, f2, "cosine", 0.1, True)
        pts1, pts2, score = map(torch.from_numpy, (d["pts1"], d["pts2"], d["score"]))
        torch.manual_seed(seed)
        R, t, s = mu.compute_coarse_Rt_overlap(atten, score, pts1, pts2, None, H, K)
        torch.manual_seed(seed)
        u = torch.rand(2, H * 3)  # the draw the reference consumed (model_utils.py:462)
        torch.manual_seed(seed)
        R0, t0, s0 = mu.compute_coarse_Rt(atten, pts1, pts2, None, H, K)
        np.savez_compressed(os.path.join(out, "pose_%s.npz" % name), f1=d["f1"], f2=d["f2"],
                            atten=atten.numpy(), score=d["score"], pts1=d["pts1"], pts2=d["pts2"],
                            H=H, K=K, seed=seed, u=u.numpy(), R=R.numpy(), t=t.numpy(), s=s.numpy(),
                            R_plain=R0.numpy(), t_plain=t0.numpy(), s_plain=s0.numpy(),
                            R_gt=d["R"], t_gt=d["t"])
        print(name, "score", s.tolist())
    # ---- fine: n x n
    for name, seed, n in (("fine_a", 201, 256), ("fine_b", 202, 500)):
        d = matching_batch(seed, 2, n, 64)
        f1, f2 = torch.from_numpy(d["f1"]), torch.from_numpy(d["f2"])
        atten = mu.compute_feature_similarity(f1, f2, "cosine", 0.1, True)
        pts1, pts2, score = map(torch.from_numpy, (d["pts1"], d["pts2"], d["score"]))
        R, t, s = mu.compute_fine_Rt_overlap(atten, score, pts1, pts2, None)
        R0, t0, s0 = mu.compute_fine_Rt(atten, pts1, pts2, None)
        np.savez_compressed(os.path.join(out, "pose_%s.npz" % name), f1=d["f1"], f2=d["f2"],
                            atten=atten.numpy(), score=d["score"], pts1=d["pts1"], pts2=d["pts2"],
                            R=R.numpy(), t=t.numpy(), s=s.numpy(),
                            R_plain=R0.numpy(), t_plain=t0.numpy(), s_plain=s0.numpy(),
                            R_gt=d["R"], t_gt=d["t"])
        print(name, "score", s.tolist())
    # ---- weighted procrustes / pairwise distance known answers
    g = torch.Generator().manual_seed(7)
    src = torch.randn(5, 40, 3, generator=g)
    Rg = torch.linalg.qr(torch.randn(5, 3, 3, generator=g))[0]
    Rg = Rg * torch.sign(torch.det(Rg)).reshape(5, 1, 1)
    tg = torch.randn(5, 3, generator=g)
    ref = src @ Rg.transpose(1, 2) + tg.unsqueeze(1) + 0.01 * torch.randn(5, 40, 3, generator=g)
    w = torch.rand(5, 40, generator=g)
    R1, t1 = mu.weighted_procrustes(src, ref, w, weight_thresh=0.3)
    R2, t2 = mu.weighted_procrustes(src[:, :3], ref[:, :3], None, weight_thresh=0.5)
    pd = mu.pairwise_distance(src, ref)
    np.savez_compressed(os.path.join(out, "pose_procrustes.npz"), src=src.numpy(), ref=ref.numpy(), w=w.numpy(),
                        R1=R1.numpy(), t1=t1.numpy(), R2=R2.numpy(), t2=t2.numpy(), pd=pd.numpy())
    print("procrustes ok")


if __name__ == "__main__":
    main()
