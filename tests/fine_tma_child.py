"""Child process of test_pose_gpu.py::test_fine_tma_passes_bit_identical_to_streaming: runs the fine solve on seeded
inputs for a list of shapes and saves every intermediate.  The TMA-fed / register-streaming choice (UPK_FINE_TMA) is
read once per process, hence one process per mode.

    python tests/fine_tma_child.py out.pt "B,N1,N2;B,N1,N2;..."
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from unopose_b200 import model_utils as MU  # noqa: E402


def main():
    out_path, spec = sys.argv[1], sys.argv[2]
    dev = torch.device("cuda:0")
    res = {}
    for si, item in enumerate(spec.split(";")):
        B, N1, N2 = (int(v) for v in item.split(","))
        rng = np.random.default_rng(100 + si)
        C = 64
        if N1 == N2:   # planted matches: a realistic mix of foreground / background rows and columns
            from unopose_b200.synthetic import matching_batch
            mb = matching_batch(100 + si, B, N1, C)
            f1, f2, p1, p2, score = (torch.from_numpy(np.ascontiguousarray(mb[k])).to(dev) for k in ("f1", "f2", "pts1", "pts2", "score"))
        else:
            f1 = torch.from_numpy(rng.standard_normal((B, N1 + 1, C), dtype=np.float32)).to(dev)
            f2 = torch.from_numpy(rng.standard_normal((B, N2 + 1, C), dtype=np.float32)).to(dev)
            # correlated features so that a good part of the rows / columns is foreground
            k = min(N1, N2)
            f2[:, 1:k + 1] += 1.5 * f1[:, 1:k + 1]
            p1 = torch.from_numpy(rng.standard_normal((B, N1, 3), dtype=np.float32)).to(dev)
            p2 = torch.from_numpy(rng.standard_normal((B, N2, 3), dtype=np.float32)).to(dev)
            score = torch.from_numpy(rng.uniform(0.2, 1.0, (B, N1 + N2)).astype(np.float32)).to(dev)
        with torch.no_grad():
            atten, stats = MU.compute_feature_similarity(f1, f2, "cosine", 0.1, True, return_stats=True)
            for tag, st in (("fused", stats), ("plain", None)):
                if tag == "fused" and st is None:
                    continue
                R, t, sc, dbg = MU._fine(atten, score, p1, p2, None, 0.15, 0.001, return_debug=True, stats=st)
                for k2, v in dict(R=R, t=t, sc=sc, **dbg).items():
                    res["%s/%s/%s" % (item, tag, k2)] = v.cpu()
            res["%s/pitched" % item] = torch.tensor(int(not atten.is_contiguous()))
    torch.cuda.synchronize()
    torch.save(res, out_path)


if __name__ == "__main__":
    main()
