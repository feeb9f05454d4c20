"""Golden vectors for the module wrappers (a14, a16, a17, a18): the REFERENCE modules run on CPU in this
container, with the pointnet2 extension calls (CUDA-only in the reference) served by the C oracle, which
is itself pinned bit-exactly against the reference extension (tests/golden/pointnet2_*.npz).

    python tests/golden/make_module_golden.py         (needs /root/reference)

Writes tests/golden/modules_small.pt (small hidden size: state dicts + inputs + outputs) and
tests/golden/modules_state_keys.json (parameter names and shapes at the real config, main_cfg.py:130-178).
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_pose_golden import import_reference  # noqa: E402
from oracle import pointnet2_oracle as O  # noqa: E402
from unopose_b200.synthetic import matching_batch  # noqa: E402


class Cfg(dict):
    __getattr__ = dict.__getitem__


def patch_pointnet2():
    import core.unopose.model.pointnet2.pointnet2_utils as pu
    import core.unopose.model.transformer as tr

    def ball_query(radius, nsample, xyz, new_xyz):
        return torch.from_numpy(O.ball_query(new_xyz.numpy(), xyz.numpy(), radius, nsample))

    def grouping_operation(features, idx):
        return torch.from_numpy(O.group_points(features.detach().numpy(), idx.numpy()))

    def gather_operation(features, idx):
        return torch.from_numpy(O.gather_points(features.detach().numpy(), idx.numpy()))

    pu.ball_query, pu.grouping_operation, pu.gather_operation = ball_query, grouping_operation, gather_operation
    tr.gather_operation = gather_operation


def main():
    mu = import_reference()
    patch_pointnet2()
    from core.unopose.model.oneref_predator_coarse_point_matching import CoarsePointMatchingOneRef
    from core.unopose.model.oneref_predator_fine_point_matching import FinePointMatchingOneRef
    from core.unopose.model.pointnet2.pointnet2_utils import LRF_batch
    from core.unopose.model.transformer import GeometricStructureEmbedding

    torch.manual_seed(0)
    torch.set_num_threads(4)
    # ---- parameter names/shapes at the real config
    real_c = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, temp=0.1, sim_type="cosine", normalize_feat=True,
                 loss_predator_thres=0.15, loss_dis_thres=0.3, nproposal1=6000, nproposal2=300)
    real_f = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, pe_radius1=0.1, pe_radius2=0.2, focusing_factor=3,
                 temp=0.1, sim_type="cosine", normalize_feat=True, loss_predator_thres=0.15, loss_dis_thres=0.3,
                 use_lrf=True, use_xyz=True, nsample1=64, nsample2=256)
    real_g = Cfg(sigma_d=0.2, sigma_a=15, angle_k=3, reduction_a="max", hidden_dim=256)
    keys = {
        "coarse": {k: list(v.shape) for k, v in CoarsePointMatchingOneRef(real_c).state_dict().items()},
        "fine": {k: list(v.shape) for k, v in FinePointMatchingOneRef(real_f).state_dict().items()},
        "geo": {k: list(v.shape) for k, v in GeometricStructureEmbedding(real_g).state_dict().items()},
    }
    json.dump(keys, open(os.path.join(HERE, "modules_state_keys.json"), "w"), indent=0, sort_keys=True)

    # ---- small instances with outputs
    C, n_c, n_f, B = 32, 24, 160, 2
    cc = Cfg(real_c, input_dim=C, hidden_dim=C, out_dim=C, nblock=2, nproposal1=300, nproposal2=30)
    cf = Cfg(real_f, input_dim=C, hidden_dim=C, out_dim=C, nblock=2, pe_radius1=0.45, pe_radius2=0.7, nsample1=16, nsample2=32)
    cg = Cfg(real_g, hidden_dim=C)
    geo = GeometricStructureEmbedding(cg).eval()
    coarse = CoarsePointMatchingOneRef(cc, return_feat=True).eval()
    fine = FinePointMatchingOneRef(cf, return_feat=True).eval()
    for m in (geo, coarse, fine):                       # non-trivial BN stats / scale params
        for name, p in m.named_parameters():
            if "scale" in name or "bn" in name:
                p.data.normal_(0.5, 0.2)
        for name, b in m.named_buffers():
            if "running_mean" in name:
                b.normal_(0, 0.1)
            if "running_var" in name:
                b.uniform_(0.5, 1.5)
    # volumetric clouds: on a thin surface every local height is ~0, the z-sign vote of the LRF ties and the
    # frame sign is then whatever the SVD backend returns (LAPACK here vs cuSOLVER on the GPU box)
    d = matching_batch(77, B, n_f, C, kind="ball")
    p1, p2 = torch.from_numpy(d["pts1"]), torch.from_numpy(d["pts2"])
    f1, f2 = torch.from_numpy(d["f1"][:, 1:]), torch.from_numpy(d["f2"][:, 1:])
    fps1 = torch.from_numpy(O.furthest_point_sampling(d["pts1"], n_c))
    fps2 = torch.from_numpy(O.furthest_point_sampling(d["pts2"], n_c))
    take = lambda x, idx: torch.gather(x, 1, idx.long().unsqueeze(2).expand(-1, -1, x.shape[2]))
    sp1, sp2, sf1, sf2 = take(p1, fps1), take(p2, fps2), take(f1, fps1), take(f2, fps2)
    bgp = torch.ones(B, 1, 3)
    with torch.no_grad():
        geo1 = geo(torch.cat([bgp, sp1], 1))
        geo2 = geo(torch.cat([bgp, sp2], 1))
        radius = torch.tensor([0.9, 1.3])
        torch.manual_seed(5)
        ep, cg1, cg2 = coarse(sp1, sf1, geo1, sp2, sf2, geo2, radius, {})
        ep_f, fg1, fg2 = fine(p1, f1, geo1, fps1, p2, f2, geo2, fps2, radius, dict(ep))
        pe_p2 = fine.PE(p2)
        grp_p2 = fine.PE.group1(p2.contiguous(), p2.contiguous(), p2.transpose(1, 2).contiguous())
        # LRF_batch / LRF
        idx = torch.from_numpy(O.ball_query(d["pts1"], d["pts1"], 0.4, 12))
        grouped = torch.from_numpy(O.group_points(np.ascontiguousarray(d["pts1"].transpose(0, 2, 1)), idx.numpy()))
        lrfb = LRF_batch(r_lrf=0.4)(p1, grouped.transpose(1, 2))
        lrf_r = torch.tensor([1.1, 0.8])
        lrfg = mu.LRF(r_lrf=lrf_r)(p1.mean(1, keepdim=True).transpose(1, 2), p1.transpose(1, 2).contiguous())
    out = dict(
        cfg_coarse=dict(cc), cfg_fine=dict(cf), cfg_geo=dict(cg),
        sd_geo=geo.state_dict(), sd_coarse=coarse.state_dict(), sd_fine=fine.state_dict(),
        p1=p1, p2=p2, f1=f1, f2=f2, fps1=fps1, fps2=fps2, sp1=sp1, sp2=sp2, sf1=sf1, sf2=sf2, radius=radius,
        geo1=geo1, geo2=geo2, coarse_g1=cg1, coarse_g2=cg2, init_R=ep["init_R"], init_t=ep["init_t"],
        init_score=ep["init_pose_score"], fine_g1=fg1, fine_g2=fg2, pred_R=ep_f["pred_R"], pred_t=ep_f["pred_t"],
        pred_score=ep_f["pred_pose_score"], pe_p2=pe_p2, grp_p2=grp_p2, lrf_grouped=grouped, lrf_batch=lrfb, lrf_r=lrf_r, lrf_global=lrfg,
        R_gt=torch.from_numpy(d["R"]), t_gt=torch.from_numpy(d["t"]),
    )
    torch.save(out, os.path.join(HERE, "modules_small.pt"))
    print("fine score", ep_f["pred_pose_score"], "size", os.path.getsize(os.path.join(HERE, "modules_small.pt")))


if __name__ == "__main__":
    main()
