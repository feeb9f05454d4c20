"""Deterministic fake model + test loader for the inference-driver tests (shared by the golden generator,
which feeds them to the REFERENCE driver, and by tests/test_inference_cpu.py)."""
import torch
import torch.nn as nn


class FakePoseModel(nn.Module):
    """end_points from the inputs alone: a rotation about z by the mean x of each instance's cloud,
    t = per-instance mean, score = sigmoid(mean z)."""

    def __init__(self):
        super().__init__()
        self.dummy = nn.Parameter(torch.zeros(1))
        self.calls = []

    def forward(self, inputs):
        pts = inputs["pts"].float()
        self.calls.append(pts.shape[0])
        a = pts[..., 0].mean(1)
        c, s = torch.cos(a), torch.sin(a)
        R = torch.zeros(pts.shape[0], 3, 3)
        R[:, 0, 0], R[:, 0, 1], R[:, 1, 0], R[:, 1, 1], R[:, 2, 2] = c, -s, s, c, 1.0
        return {"pred_R": R, "pred_t": pts.mean(1), "pred_pose_score": torch.sigmoid(pts[..., 2].mean(1))}


class _Dataset:
    def __init__(self, dets):
        self.dets = dets


class FakeLoader(list):
    pass


def make_loader(seed=0, n_images=3, with_tem_pose=True):
    g = torch.Generator().manual_seed(seed)
    samples, dets = [], {}
    for im in range(n_images):
        n = [5, 37, 16][im % 3]            # fewer than / more than / exactly a multiple of the chunk size
        scene_id, img_id = 48 + im, 7 * im + 1
        d = {
            "pts": torch.randn(1, n, 64, 3, generator=g),
            "rgb": torch.rand(1, n, 3, 8, 8, generator=g),
            "rgb_choose": torch.randint(0, 64, (1, n, 64), generator=g),
            "tem1_rgb": torch.rand(1, n, 3, 8, 8, generator=g),
            "tem1_choose": torch.randint(0, 64, (1, n, 64), generator=g),
            "tem1_pts": torch.randn(1, n, 64, 3, generator=g),
            "score": torch.rand(1, n, 1, generator=g),
            "scene_id": torch.tensor([scene_id]),
            "img_id": torch.tensor([img_id]),
            "inst_ids": torch.arange(n).flip(0).unsqueeze(0),
            "obj_id": torch.randint(1, 22, (1, n), generator=g),
            "seg_time": torch.tensor([0.125 + im]),
        }
        if with_tem_pose:
            q, _ = torch.linalg.qr(torch.randn(n, 3, 3, generator=g))
            T = torch.zeros(n, 4, 4)
            T[:, :3, :3], T[:, :3, 3], T[:, 3, 3] = q, torch.randn(n, 3, generator=g), 1.0
            d["tem1_pose"] = T.unsqueeze(0)
        samples.append(d)
        dets[f"{scene_id:06d}_{img_id:06d}"] = [{"obj_id": int(d["obj_id"][0][k]), "score": float(d["score"][0, k, 0])}
                                               for k in range(n)]
    loader = FakeLoader(samples)
    loader.dataset = _Dataset(dets)
    return loader
