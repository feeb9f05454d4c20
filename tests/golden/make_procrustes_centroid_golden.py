"""Golden vectors for `weighted_procrustes(..., src_centroid=, ref_centroid=)` (reference model_utils.py:710-721), made by
IMPORTING THE REFERENCE (/root/reference, read-only) and running it on CPU tensors.

    python tests/golden/make_procrustes_centroid_golden.py      (only works where /root/reference exists)

Writes tests/golden/pose_procrustes_centroids.npz: inputs + the reference's (R, t) for both / src-only / ref-only
centroids, in the (B,3) and (B,1,3) forms the reference accepts.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_pose_golden import ROOT, import_reference  # noqa: E402


def main():
    mu = import_reference()
    torch.set_num_threads(1)
    g = torch.Generator().manual_seed(11)
    B, N = 6, 64
    src = torch.randn(B, N, 3, generator=g)
    Rg = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    Rg = Rg * torch.sign(torch.det(Rg)).reshape(B, 1, 1)
    tg = torch.randn(B, 3, generator=g)
    ref = src @ Rg.transpose(1, 2) + tg.unsqueeze(1) + 0.01 * torch.randn(B, N, 3, generator=g)
    w = torch.rand(B, N, generator=g)
    # centroids that are NOT the weighted means (plain means + an offset): the branch must really be taken
    cs = src.mean(1) + 0.05 * torch.randn(B, 3, generator=g)
    cr = ref.mean(1) + 0.05 * torch.randn(B, 3, generator=g)
    Rb, tb = mu.weighted_procrustes(src, ref, w, weight_thresh=0.2, src_centroid=cs, ref_centroid=cr.unsqueeze(1))
    Rs, ts = mu.weighted_procrustes(src, ref, w, weight_thresh=0.2, src_centroid=cs.unsqueeze(1))
    Rr, tr = mu.weighted_procrustes(src, ref, None, ref_centroid=cr)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pose_procrustes_centroids.npz"), src=src.numpy(),
                        ref=ref.numpy(), w=w.numpy(), cs=cs.numpy(), cr=cr.numpy(), Rb=Rb.numpy(), tb=tb.numpy(),
                        Rs=Rs.numpy(), ts=ts.numpy(), Rr=Rr.numpy(), tr=tr.numpy())
    print("centroid goldens ok")


if __name__ == "__main__":
    main()
