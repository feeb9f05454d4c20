"""Dev tool: CUDA-event timings of the pointnet2 kernels vs the reference extension (when present)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_ext  # noqa: E402
from unopose_b200.pointnet2 import _ext as mine  # noqa: E402
from util_clouds import batch_clouds  # noqa: E402


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
    dev = torch.device("cuda:0")
    ref = ref_ext.load()
    for B in (1, 16, 64):
        tem = torch.from_numpy(batch_clouds(1, B, 5000, "surface")).to(dev)
        pts = torch.from_numpy(batch_clouds(2, B, 2048, "surface")).to(dev)
        f3 = pts.transpose(1, 2).contiguous()
        rows = []
        for name, impl in (("mine", mine), ("ref", ref)):
            if impl is None:
                continue
            t1 = timeit(lambda: impl.furthest_point_sampling(tem, 2048), it=5)
            t2 = timeit(lambda: impl.furthest_point_sampling(pts, 196))
            t3 = timeit(lambda: impl.ball_query(pts, pts, 0.1, 64))
            t4 = timeit(lambda: impl.ball_query(pts, pts, 0.2, 256))
            idx = impl.ball_query(pts, pts, 0.2, 256)
            t5 = timeit(lambda: impl.group_points(f3, idx))
            fi = impl.furthest_point_sampling(pts, 196)
            feats = torch.randn(B, 256, 2048, device=dev)
            t6 = timeit(lambda: impl.gather_points(feats, fi))
            rows.append((name, t1, t2, t3, t4, t5, t6))
        for r in rows:
            print("B=%d %-5s fps5000->2048 %.1fus  fps2048->196 %.1fus  bq(.1,64) %.1fus  bq(.2,256) %.1fus  group256 %.1fus  gather256ch %.1fus" % ((B,) + r))


if __name__ == "__main__":
    main()
