"""CPU: the host-side (torch) parts of the module wrappers against golden vectors produced by the
REFERENCE modules (tests/golden/make_module_golden.py), and state-dict compatibility at the real config."""
import json
import os

import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


class Cfg(dict):
    __getattr__ = dict.__getitem__


def _g():
    return torch.load(os.path.join(GOLD, "modules_small.pt"), weights_only=False)


def test_state_dict_names_and_shapes_match_reference():
    from unopose_b200.modules import CoarsePointMatchingOneRef, FinePointMatchingOneRef, GeometricStructureEmbedding

    keys = json.load(open(os.path.join(GOLD, "modules_state_keys.json")))
    real_c = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, temp=0.1, sim_type="cosine", normalize_feat=True,
                 nproposal1=6000, nproposal2=300)
    real_f = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, pe_radius1=0.1, pe_radius2=0.2, focusing_factor=3,
                 temp=0.1, sim_type="cosine", normalize_feat=True, use_lrf=True, use_xyz=True, nsample1=64, nsample2=256)
    real_g = Cfg(sigma_d=0.2, sigma_a=15, angle_k=3, reduction_a="max", hidden_dim=256)
    for name, mod in (("coarse", CoarsePointMatchingOneRef(real_c)), ("fine", FinePointMatchingOneRef(real_f)),
                      ("geo", GeometricStructureEmbedding(real_g))):
        mine = {k: list(v.shape) for k, v in mod.state_dict().items()}
        assert mine == keys[name], (name, set(mine) ^ set(keys[name]))


def test_geometric_embedding_matches_reference():
    from unopose_b200.modules import GeometricStructureEmbedding

    g = _g()
    m = GeometricStructureEmbedding(Cfg(g["cfg_geo"])).eval()
    m.load_state_dict(g["sd_geo"])
    with torch.no_grad():
        out = m(torch.cat([torch.ones(2, 1, 3), g["sp1"]], 1))
    assert torch.allclose(out, g["geo1"], atol=2e-5, rtol=1e-5)


def test_coarse_module_features_match_reference():
    from unopose_b200.modules import CoarsePointMatchingOneRef

    g = _g()
    m = CoarsePointMatchingOneRef(Cfg(g["cfg_coarse"])).eval()
    m.load_state_dict(g["sd_coarse"])           # strict: every reference parameter has a home
    with torch.no_grad():
        g1, g2, score = m.matching_features(g["sf1"], g["geo1"], g["sf2"], g["geo2"])
    assert torch.allclose(g1, g["coarse_g1"], atol=1e-5, rtol=1e-4)
    assert torch.allclose(g2, g["coarse_g2"], atol=1e-5, rtol=1e-4)
    assert score.shape == (2, 2 * g["sp1"].shape[1]) and (score >= 0).all() and (score <= 1).all()


def test_lrf_batch_and_global_lrf_match_reference():
    from unopose_b200.model_utils import LRF
    from unopose_b200.pointnet2.lrf import LRF_batch

    g = _g()
    out = LRF_batch(r_lrf=0.4)(g["p1"], g["lrf_grouped"].transpose(1, 2))
    assert torch.allclose(out, g["lrf_batch"], atol=1e-4, rtol=1e-4)
    p1 = g["p1"]
    out = LRF(r_lrf=g["lrf_r"])(p1.mean(1, keepdim=True).transpose(1, 2), p1.transpose(1, 2).contiguous())
    assert torch.allclose(out, g["lrf_global"], atol=1e-4, rtol=1e-4)


def test_training_branch_is_explicitly_unsupported():
    import pytest

    from unopose_b200.modules import CoarsePointMatchingOneRef

    g = _g()
    m = CoarsePointMatchingOneRef(Cfg(g["cfg_coarse"])).train()
    with pytest.raises(NotImplementedError):
        m(g["sp1"], g["sf1"], g["geo1"], g["sp2"], g["sf2"], g["geo2"], g["radius"], {})


def test_batch_norm_folding_of_shared_mlp():
    """modules/pe.py::fold_shared_mlp (what the fused MLP kernel is fed): relu(W' x + b') per layer == the eval-mode
    conv + batch norm + ReLU stack (pointnet2/pytorch_utils.py:25-48)."""
    from unopose_b200.modules.layers import SharedMLP
    from unopose_b200.modules.pe import fold_shared_mlp

    torch.manual_seed(3)
    mlp = SharedMLP([6, 32, 64, 128], bn=True)
    for layer in mlp:
        bn = layer.normlayer.bn
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.normal_(0, 0.2)
    mlp.eval()
    x = torch.randn(2, 6, 5, 7)
    with torch.no_grad():
        ref = mlp(x)
        h = x
        for w, b in fold_shared_mlp(mlp):
            h = torch.relu(torch.einsum("oc,bcmn->bomn", w, h) + b.view(1, -1, 1, 1))
    assert torch.allclose(h, ref, atol=1e-5, rtol=1e-5)
    plain = SharedMLP([3, 32, 64, 128], bn=False).eval()     # no batch norm: conv bias passes through
    with torch.no_grad():
        h = x[:, :3]
        for w, b in fold_shared_mlp(plain):
            h = torch.relu(torch.einsum("oc,bcmn->bomn", w, h) + b.view(1, -1, 1, 1))
        assert torch.allclose(h, plain(x[:, :3]), atol=1e-6, rtol=1e-5)


def test_rpe_score_term_equals_reference_formulation():
    """modules/transformer.py::_DotAttention projects the queries by W_p; the reference projects the (B,N,M,C)
    embedding (transformer.py:392-395).  Same scores and outputs."""
    from unopose_b200.modules.transformer import _DotAttention, _heads, _merge

    torch.manual_seed(4)
    att = _DotAttention(64, 4, relative=True).eval()
    x = torch.randn(2, 9, 64)
    emb = torch.randn(2, 9, 9, 64)
    with torch.no_grad():
        out, attn = att(x, x, emb)
        h = att.num_heads
        q, k, v = _heads(att.proj_q(x), h), _heads(att.proj_k(x), h), _heads(att.proj_v(x), h)
        p = att.proj_p(emb).reshape(2, 9, 9, h, att.head_dim)
        scores = (q @ k.transpose(-1, -2) + torch.einsum("bhnc,bnmhc->bhnm", q, p)) / att.head_dim ** 0.5
        ref_attn = torch.softmax(scores, dim=-1)
    assert torch.allclose(attn, ref_attn, atol=1e-6, rtol=1e-5)
    assert torch.allclose(out, _merge(ref_attn @ v), atol=1e-6, rtol=1e-5)


def test_fused_kernel_eligibility_is_answered_without_a_gpu():
    from unopose_b200 import _lib

    lib = _lib.load()
    assert lib.upk_geometric_embedding_supported(256, 3) == 1 and lib.upk_geometric_embedding_supported(32, 8) == 1
    assert lib.upk_geometric_embedding_supported(48, 3) == 0 and lib.upk_geometric_embedding_supported(256, 9) == 0
    assert lib.upk_shared_mlp_max_supported(6, 32, 64, 128, 2048, 256) == 1
    assert lib.upk_shared_mlp_max_supported(6, 32, 64, 128, 2048, 64) == 1
    assert lib.upk_shared_mlp_max_supported(6, 32, 64, 128, 2048, 48) == 0      # 48 neither divides 128 nor is a multiple
    assert lib.upk_shared_mlp_max_supported(17, 32, 64, 128, 2048, 64) == 0
    assert lib.upk_shared_mlp_max_supported(6, 32, 64, 256, 2048, 64) == 0
    assert lib.upk_geometric_embedding_workspace_bytes(16, 197, 256, 3) > 16 * 197 * 197 * 4 * 4


def _real_inputs():
    import numpy as np

    from unopose_b200.synthetic import matching_batch

    g = np.load(os.path.join(GOLD, "modules_real.npz"))
    d = matching_batch(int(g["seed_inputs"]), 1, 196, 256, kind="ball")
    T = torch.from_numpy
    return g, T(d["pts1"]), T(d["pts2"]), T(d["f1"][:, 1:].copy()), T(d["f2"][:, 1:].copy())


REAL_C = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, temp=0.1, sim_type="cosine", normalize_feat=True,
             nproposal1=6000, nproposal2=300)
REAL_G = Cfg(sigma_d=0.2, sigma_a=15, angle_k=3, reduction_a="max", hidden_dim=256)


def test_real_config_modules_match_reference_features():
    """GeometricStructureEmbedding + CoarsePointMatchingOneRef at the REAL config (hidden 256, 3 blocks, 196 points)
    against the reference modules' outputs (tests/golden/make_real_golden.py); weights are key-addressed
    (tests/util_state.py), inputs regenerated from the seed.  CPU: the torch branches of the product modules."""
    from unopose_b200.modules import CoarsePointMatchingOneRef, GeometricStructureEmbedding
    from util_state import keyed_state_dict

    g, sp1, sp2, sf1, sf2 = _real_inputs()
    geo = GeometricStructureEmbedding(REAL_G).eval()
    geo.load_state_dict(keyed_state_dict(geo.state_dict(), int(g["seed_weights"])))
    m = CoarsePointMatchingOneRef(REAL_C).eval()
    m.load_state_dict(keyed_state_dict(m.state_dict(), int(g["seed_weights"])))
    bgp = torch.ones(1, 1, 3)
    with torch.no_grad():
        geo1 = geo(torch.cat([bgp, sp1], 1))
        geo2 = geo(torch.cat([bgp, sp2], 1))
        g1, g2, score = m.matching_features(sf1, geo1, sf2, geo2)
    assert torch.allclose(geo1[:, ::7, ::5], torch.from_numpy(g["geo1_sample"]), atol=2e-5, rtol=1e-5)
    assert abs(float(geo1.mean()) - float(g["geo1_mean"])) < 1e-6
    assert torch.allclose(g1, torch.from_numpy(g["coarse_g1"]), atol=2e-5, rtol=1e-4)
    assert torch.allclose(g2, torch.from_numpy(g["coarse_g2"]), atol=2e-5, rtol=1e-4)


def test_linear_multi_cpu_fallback_and_weight_cache():
    """`linear_multi` (q / k / v projections as one GEMM on CUDA) falls back to one `F.linear` per layer on CPU, and its
    cached weight concatenation follows in-place parameter updates and reloaded state dicts."""
    import torch.nn as nn
    from unopose_b200.modules.linear import _cat_params, linear_multi

    torch.manual_seed(0)
    layers = [nn.Linear(32, 16), nn.Linear(32, 16), nn.Linear(32, 48)]
    x = torch.randn(2, 5, 32)
    with torch.no_grad():
        outs = linear_multi(layers, x)
        for l, o in zip(layers, outs):
            assert torch.equal(o, l(x))
        w, b = _cat_params(layers)
        assert w.shape == (80, 32) and torch.equal(w[16:32], layers[1].weight) and torch.equal(b[32:], layers[2].bias)
        w_again, _ = _cat_params(layers)
        assert w_again is w                                      # cached
        layers[1].weight.mul_(2.0)                               # in-place update: version counter moves
        w2, _ = _cat_params(layers)
        assert w2 is not w and torch.equal(w2[16:32], layers[1].weight)
        layers[2].load_state_dict({"weight": torch.ones(48, 32), "bias": torch.zeros(48)})
        w3, b3 = _cat_params(layers)
        assert torch.equal(w3[32:], torch.ones(48, 32)) and torch.equal(b3[32:], torch.zeros(48))
