"""Golden vectors for the fused GeometricStructureEmbedding: the REFERENCE module
(core/unopose/model/transformer.py:287-350) run on CPU in this container.

    python tests/golden/make_geo_golden.py         (needs /root/reference)

Writes tests/golden/geo_real.npz (hidden 256, N = 197, k = 3, max — the real config, main_cfg.py:142-148; the
(197,197,256) output is kept for 3 rows i only, plus its sum per i) and tests/golden/geo_small.npz (hidden 64,
N = 32, k = 2, mean reduction, complete).  Inputs are regenerated from the seed by oracle.geo_oracle.make_inputs; a
checksum of them is stored so that drift of the CPU generator would be noticed.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_pose_golden import import_reference  # noqa: E402
from oracle import geo_oracle as G  # noqa: E402

ROWS = [0, 57, 196]


class Cfg(dict):
    __getattr__ = dict.__getitem__


def run(seed, B, N, C, k, red):
    from core.unopose.model.transformer import GeometricStructureEmbedding

    pts, dterm, w_d, b_d, w_a, b_a = G.make_inputs(seed, B, N, C)
    m = GeometricStructureEmbedding(Cfg(sigma_d=0.2, sigma_a=15, angle_k=k, reduction_a=red, hidden_dim=C)).eval()
    m.load_state_dict({"embedding.div_term": dterm, "proj_d.weight": w_d, "proj_d.bias": b_d, "proj_a.weight": w_a,
                       "proj_a.bias": b_a})
    with torch.no_grad():
        d_idx, a_idx = m.get_embedding_indices(pts)
        out = m(pts)
    chk = float(sum(t.double().sum() for t in (pts, w_d, b_d, w_a, b_a)))
    return pts, d_idx, a_idx, out, chk


def main():
    import_reference()
    torch.set_num_threads(8)
    pts, d_idx, a_idx, out, chk = run(11, 1, 197, 256, 3, "max")
    rows = [r for r in ROWS if r < 197]
    np.savez_compressed(os.path.join(HERE, "geo_real.npz"), seed=11, B=1, N=197, C=256, k=3, red="max", checksum=chk,
                        rows=np.array(rows), d_idx=d_idx.numpy(), a_idx=a_idx.numpy(),
                        out_rows=out[0, rows].numpy(), out_rowsum=out.double().sum(dim=(2, 3)).numpy())
    pts, d_idx, a_idx, out, chk = run(12, 2, 32, 64, 2, "mean")
    np.savez_compressed(os.path.join(HERE, "geo_small.npz"), seed=12, B=2, N=32, C=64, k=2, red="mean", checksum=chk,
                        d_idx=d_idx.numpy(), a_idx=a_idx.numpy(), out=out.numpy())
    for f in ("geo_real.npz", "geo_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
