"""CPU: the geometric-embedding oracle (oracle/geo_oracle.py) against golden vectors produced by the REFERENCE
module (tests/golden/make_geo_golden.py; `geo1` of modules_small.pt comes from make_module_golden.py)."""
import os

import numpy as np
import torch

from oracle import geo_oracle as G

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _case(name):
    z = np.load(os.path.join(GOLD, name))
    inp = G.make_inputs(int(z["seed"]), int(z["B"]), int(z["N"]), int(z["C"]))
    chk = float(sum(t.double().sum() for t in (inp[0], *inp[2:])))
    assert abs(chk - float(z["checksum"])) < 1e-9, "CPU generator drifted: regenerate the golden vectors"
    return z, inp


def test_indices_and_embedding_real_config():
    z, (pts, dterm, w_d, b_d, w_a, b_a) = _case("geo_real.npz")
    d_idx, a_idx = G.embedding_indices(pts, 0.2, 15, int(z["k"]))
    assert np.array_equal(d_idx.numpy(), z["d_idx"]) and np.array_equal(a_idx.numpy(), z["a_idx"])
    out = G.embed_from_indices(d_idx, a_idx, dterm, w_d, b_d, w_a, b_a, str(z["red"]))
    rows = z["rows"].tolist()
    assert torch.allclose(out[0, rows], torch.from_numpy(z["out_rows"]), atol=2e-6, rtol=1e-6)
    assert np.allclose(out.double().sum(dim=(2, 3)).numpy(), z["out_rowsum"], atol=1e-3, rtol=0)


def test_small_mean_reduction():
    z, (pts, dterm, w_d, b_d, w_a, b_a) = _case("geo_small.npz")
    out = G.geometric_embedding(pts, dterm, w_d, b_d, w_a, b_a, 0.2, 15, int(z["k"]), str(z["red"]))
    assert torch.allclose(out, torch.from_numpy(z["out"]), atol=2e-6, rtol=1e-6)


def test_module_golden_geo1():
    g = torch.load(os.path.join(GOLD, "modules_small.pt"), weights_only=False)
    sd, cfg = g["sd_geo"], g["cfg_geo"]
    pts = torch.cat([torch.ones(2, 1, 3), g["sp1"]], 1)
    out = G.geometric_embedding(pts, sd["embedding.div_term"], sd["proj_d.weight"], sd["proj_d.bias"], sd["proj_a.weight"],
                                sd["proj_a.bias"], cfg["sigma_d"], cfg["sigma_a"], cfg["angle_k"], cfg["reduction_a"])
    assert torch.allclose(out, g["geo1"], atol=2e-6, rtol=1e-6)
