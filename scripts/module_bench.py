"""Dev tool: device time of the matching modules at the real config (main_cfg.py:130-178), random-init weights,
synthetic clouds: geometric embedding (fused vs the reference torch-op sequence), coarse module, fine module, and
the RPE score term in the reference's formulation vs the restructured one (modules/transformer.py).
    python scripts/module_bench.py [B]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unopose_b200.modules import CoarsePointMatchingOneRef, FinePointMatchingOneRef, GeometricStructureEmbedding  # noqa: E402
from unopose_b200.pointnet2 import pointnet2_utils as PU  # noqa: E402


class Cfg(dict):
    __getattr__ = dict.__getitem__


def timeit(fn, it=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    torch.backends.cuda.matmul.allow_tf32 = False
    cc = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, temp=0.1, sim_type="cosine", normalize_feat=True,
             nproposal1=6000, nproposal2=300)
    cf = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, pe_radius1=0.1, pe_radius2=0.2, focusing_factor=3,
             temp=0.1, sim_type="cosine", normalize_feat=True, use_lrf=True, use_xyz=True, nsample1=64, nsample2=256)
    cg = Cfg(sigma_d=0.2, sigma_a=15, angle_k=3, reduction_a="max", hidden_dim=256)
    geo = GeometricStructureEmbedding(cg).to(dev).eval()
    coarse = CoarsePointMatchingOneRef(cc).to(dev).eval()
    fine = FinePointMatchingOneRef(cf).to(dev).eval()

    def cloud(n):
        p = torch.randn(B, n, 3, device=dev)
        return p / p.norm(dim=2).max(dim=1)[0].view(B, 1, 1)

    p2 = cloud(2048)
    R = torch.linalg.qr(torch.randn(B, 3, 3, device=dev))[0]
    R = R * torch.sign(torch.det(R)).view(B, 1, 1)
    p1 = p2 @ R.transpose(1, 2) + 0.01 * torch.randn(B, 2048, 3, device=dev)
    f2 = torch.randn(B, 2048, 256, device=dev)
    f1 = f2 + 0.5 * torch.randn(B, 2048, 256, device=dev)
    radius = torch.ones(B, device=dev)
    res = {"B": B}
    with torch.no_grad():
        i1 = PU.furthest_point_sample(p1.contiguous(), 196)
        i2 = PU.furthest_point_sample(p2.contiguous(), 196)
        gat = lambda x, i: torch.gather(x, 1, i.long().unsqueeze(-1).expand(-1, -1, x.shape[2]))  # noqa: E731
        sp1, sp2, sf1, sf2 = gat(p1, i1), gat(p2, i2), gat(f1, i1), gat(f2, i2)
        bg = torch.ones(B, 1, 3, device=dev)
        g_in1, g_in2 = torch.cat([bg, sp1], 1), torch.cat([bg, sp2], 1)
        res["geo_embedding_x2_fused_ms"] = timeit(lambda: (geo(g_in1), geo(g_in2)))
        geo1, geo2 = geo(g_in1), geo(g_in2)

        def geo_torch(pts):     # the reference op sequence (the module's non-fused branch)
            d_idx, a_idx = geo.get_embedding_indices(pts)
            d = geo.proj_d(geo.embedding(d_idx))
            a = geo.proj_a(geo.embedding(a_idx)).max(dim=3)[0]
            return d + a
        res["geo_embedding_x2_torch_ms"] = timeit(lambda: (geo_torch(g_in1), geo_torch(g_in2)), it=2, warm=1)
        try:
            res["coarse_module_ms"] = timeit(lambda: coarse(sp1, sf1, geo1, sp2, sf2, geo2, radius, {}))
            ep = coarse(sp1, sf1, geo1, sp2, sf2, geo2, radius, {})
            res["fine_module_ms"] = timeit(lambda: fine(p1, f1, geo1, i1, p2, f2, geo2, i2, radius, dict(ep)), it=3, warm=1)
        except Exception as e:  # noqa: BLE001
            res["module_error"] = repr(e)[:300]
        # the RPE score term of one self-attention layer call, both formulations
        att = coarse.transformers[0].layers[0].attention.attention
        h, c = att.num_heads, att.head_dim
        x = torch.randn(B, 197, 256, device=dev)
        q = att.proj_q(x).reshape(B, 197, h, c).permute(0, 2, 1, 3)

        def rpe_ref():
            p = att.proj_p(geo1).reshape(B, 197, 197, h, c)
            return torch.einsum("bhnc,bnmhc->bhnm", q, p)

        def rpe_new():
            q2 = torch.einsum("bhnc,hck->bnkh", q, att.proj_p.weight.view(h, c, -1))
            return torch.matmul(geo1, q2).permute(0, 3, 1, 2) + torch.einsum("bhnc,hc->bhn", q, att.proj_p.bias.view(h, c)).unsqueeze(-1)
        res["rpe_term_reference_formulation_ms"] = timeit(rpe_ref)
        res["rpe_term_restructured_ms"] = timeit(rpe_new)
        res["rpe_term_max_abs_diff"] = float((rpe_ref() - rpe_new()).abs().max())
    if "fine_module_ms" in res:
        res["instances_per_s_matching_modules"] = B / (res["geo_embedding_x2_fused_ms"] + res["coarse_module_ms"]
                                                       + res["fine_module_ms"]) * 1e3
    print(json.dumps(res))


if __name__ == "__main__":
    main()
