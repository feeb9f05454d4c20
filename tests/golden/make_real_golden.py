"""Golden vectors at the REAL sizes of the path (configs/main_cfg.py:130-178), produced by IMPORTING THE REFERENCE and
running its own functions / modules on CPU tensors in this container.

    python tests/golden/make_real_golden.py            (needs /root/reference; ~2 min of CPU)

Inputs are regenerated at test time from seeds (numpy PCG64 generators of unopose_b200/synthetic.py and
tests/util_state.py are platform-independent), so only the reference's OUTPUTS are stored:
  pose_real.npz     compute_fine_Rt_overlap / compute_fine_Rt at 2048 x 2048 x 256 (B = 2);
                    compute_coarse_Rt_overlap at 196 x 196 x 256, H = 6000, K = 300 (B = 2) with the uniforms it drew
  modules_real.npz  GeometricStructureEmbedding (N = 197, hidden 256, k = 3) and CoarsePointMatchingOneRef
                    (3 blocks, hidden 256) forward features with key-addressed weights (seed 5)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
from make_pose_golden import import_reference  # noqa: E402
from util_state import keyed_state_dict  # noqa: E402
from unopose_b200.synthetic import matching_batch  # noqa: E402


class Cfg(dict):
    __getattr__ = dict.__getitem__


def main():
    mu = import_reference()
    torch.set_num_threads(8)
    out = {}
    # ---- fine pose at the real size
    d = matching_batch(211, 2, 2048, 256)
    f1, f2 = torch.from_numpy(d["f1"]), torch.from_numpy(d["f2"])
    atten = mu.compute_feature_similarity(f1, f2, "cosine", 0.1, True)
    pts1, pts2, score = map(torch.from_numpy, (d["pts1"], d["pts2"], d["score"]))
    R, t, s = mu.compute_fine_Rt_overlap(atten, score, pts1, pts2, None)
    R0, t0, s0 = mu.compute_fine_Rt(atten, pts1, pts2, None)
    out.update(fine_seed=211, fine_R=R.numpy(), fine_t=t.numpy(), fine_s=s.numpy(), fine_R_plain=R0.numpy(),
               fine_t_plain=t0.numpy(), fine_s_plain=s0.numpy(),
               fine_atten_sample=atten[:, ::97, ::89].numpy())
    print("fine", s.tolist())
    # ---- coarse pose at the real config (nproposal1 = 6000, main_cfg.py:160)
    d = matching_batch(212, 2, 196, 256)
    f1, f2 = torch.from_numpy(d["f1"]), torch.from_numpy(d["f2"])
    atten = mu.compute_feature_similarity(f1, f2, "cosine", 0.1, True)
    pts1, pts2, score = map(torch.from_numpy, (d["pts1"], d["pts2"], d["score"]))
    torch.manual_seed(212)
    R, t, s = mu.compute_coarse_Rt_overlap(atten, score, pts1, pts2, None, 6000, 300)
    torch.manual_seed(212)
    u = torch.rand(2, 6000 * 3)
    out.update(coarse_seed=212, coarse_u=u.numpy(), coarse_R=R.numpy(), coarse_t=t.numpy(), coarse_s=s.numpy())
    print("coarse", s.tolist())
    np.savez_compressed(os.path.join(HERE, "pose_real.npz"), **out)

    # ---- modules at the real config
    from core.unopose.model.oneref_predator_coarse_point_matching import CoarsePointMatchingOneRef
    from core.unopose.model.transformer import GeometricStructureEmbedding

    real_c = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, temp=0.1, sim_type="cosine", normalize_feat=True,
                 loss_predator_thres=0.15, loss_dis_thres=0.3, nproposal1=6000, nproposal2=300)
    real_g = Cfg(sigma_d=0.2, sigma_a=15, angle_k=3, reduction_a="max", hidden_dim=256)
    geo = GeometricStructureEmbedding(real_g).eval()
    geo.load_state_dict(keyed_state_dict(geo.state_dict(), 5))
    coarse = CoarsePointMatchingOneRef(real_c, return_feat=True).eval()
    coarse.load_state_dict(keyed_state_dict(coarse.state_dict(), 5))
    d = matching_batch(213, 1, 196, 256, kind="ball")
    sp1, sp2 = torch.from_numpy(d["pts1"]), torch.from_numpy(d["pts2"])
    sf1, sf2 = torch.from_numpy(d["f1"][:, 1:]), torch.from_numpy(d["f2"][:, 1:])
    bgp = torch.ones(1, 1, 3)
    with torch.no_grad():
        geo1 = geo(torch.cat([bgp, sp1], 1))
        geo2 = geo(torch.cat([bgp, sp2], 1))
        torch.manual_seed(7)
        ep, g1, g2 = coarse(sp1, sf1, geo1, sp2, sf2, geo2, torch.ones(1), {})
    m = dict(seed_weights=5, seed_inputs=213, geo1_sample=geo1[:, ::7, ::5].numpy(), geo1_mean=float(geo1.mean()),
             geo1_absmax=float(geo1.abs().max()), coarse_g1=g1.numpy(), coarse_g2=g2.numpy(),
             init_R=ep["init_R"].numpy(), init_t=ep["init_t"].numpy(), init_score=ep["init_pose_score"].numpy(),
             R_gt=d["R"], t_gt=d["t"])
    np.savez_compressed(os.path.join(HERE, "modules_real.npz"), **m)
    print("coarse module score", ep["init_pose_score"].tolist(),
          {k: os.path.getsize(os.path.join(HERE, k)) for k in ("pose_real.npz", "modules_real.npz")})


if __name__ == "__main__":
    main()
