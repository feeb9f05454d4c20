"""Generate golden vectors for the pointnet2 ops by running the REFERENCE'S OWN
extension (compiled unmodified, oracle/build_ref_ext.py) on a B200.

    gpurun -- python tests/golden/make_pointnet2_golden.py   (writes gpurun_out/golden/*.npz)

then copy gpurun_out/golden/pointnet2_*.npz into tests/golden/.  The reference
extension has no CPU path, so these fixtures can only be produced on a GPU box.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_ext  # noqa: E402
from util_clouds import batch_clouds  # noqa: E402

CASES = [  # name, seed, b, n, m_fps, (r1, ns1), (r2, ns2), kind
    ("a", 11, 2, 777, 64, (0.25, 16), (0.5, 32), "ball"),
    ("b", 12, 1, 2048, 196, (0.1, 64), (0.2, 256), "surface"),
    ("c", 13, 3, 196, 50, (0.3, 8), (2.5, 24), "ball"),
    ("dup", 14, 1, 600, 40, (0.2, 16), (0.4, 16), "ball"),
]


def main():
    ext = ref_ext.load()
    assert ext is not None, "oracle/_ref/ref_pointnet2_ext.so missing"
    dev = torch.device("cuda:0")
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, seed, b, n, m, (r1, ns1), (r2, ns2), kind in CASES:
        rng = np.random.default_rng(seed)
        xyz = batch_clouds(seed, b, n, kind)
        if name == "dup":  # duplicated points and quantised coordinates -> distance ties
            xyz = np.round(xyz * 8) / 8
            xyz[:, 300:] = xyz[:, :300]
        feat = rng.standard_normal((b, 5, n)).astype(np.float32)
        unknown = batch_clouds(seed + 100, b, 37, kind)
        nn_w = rng.random((b, 37, 3)).astype(np.float32)
        t = lambda a: torch.from_numpy(a).to(dev)
        fps = ext.furthest_point_sampling(t(xyz), m)
        bq1 = ext.ball_query(t(xyz), t(xyz), r1, ns1)
        bq2 = ext.ball_query(t(xyz), t(xyz), r2, ns2)
        grouped = ext.group_points(t(feat), bq1)
        gathered = ext.gather_points(t(feat), fps)
        d2, i3 = ext.three_nn(t(unknown), t(xyz))
        interp = ext.three_interpolate(t(feat), i3, t(nn_w))
        np.savez_compressed(
            os.path.join(out_dir, "pointnet2_%s.npz" % name),
            xyz=xyz, feat=feat, unknown=unknown, nn_w=nn_w,
            fps_idx=fps.cpu().numpy(), bq1_idx=bq1.cpu().numpy(), bq1_radius=np.float32(r1),
            bq2_idx=bq2.cpu().numpy(), bq2_radius=np.float32(r2),
            grouped=grouped.cpu().numpy(), gathered=gathered.cpu().numpy(),
            nn_dist2=d2.cpu().numpy(), nn_idx=i3.cpu().numpy(), interp=interp.cpu().numpy(),
        )
        print("wrote", name)


if __name__ == "__main__":
    main()
