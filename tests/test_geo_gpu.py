"""GPU: the fused GeometricStructureEmbedding kernels (unopose_b200/csrc/geoembed.cu) against the
oracle (torch ops of the reference, on the same device and on CPU) and against golden vectors of the REFERENCE module.

Tolerances.  Indices: 2e-5 absolute off the diagonal (the expansion-form distance cancels: 1e-7 on d^2 is 5e-6 on a
d_idx of 0.1/0.2; cuBLAS and CPU matmul differ by as much); the diagonal d_ii is sqrt(rounding noise) in every
implementation (0 ... 3e-3) and is compared loosely.  Embedding given identical indices: 2e-5 absolute on values of
magnitude ~1 (3xTF32 products accumulate in fp32 like cuBLAS SGEMM; argument of the sinusoids up to ~40 rad).
"""
import math
import os

import numpy as np
import pytest
import torch

from oracle import geo_oracle as G

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _offdiag(n, dev):
    return ~torch.eye(n, dtype=torch.bool, device=dev)


def _check_indices(d, a, d_ref, a_ref, pts, k):
    n = d.shape[1]
    m = _offdiag(n, d.device)
    assert (d - d_ref)[:, m].abs().max() < 2e-5
    assert (d - d_ref).abs().max() < 5e-3                      # diagonal: sqrt of cancellation noise / sigma_d
    # angles: rows i whose neighbour set agrees (ties / near-ties of the k-th neighbour may legitimately differ)
    bad = ((a - a_ref).abs() > 2e-4).flatten(2).any(2)         # (B,N) rows with any disagreeing angle
    assert bad.float().mean() < 0.02
    return bad


@pytest.mark.parametrize("B,N,C,k,red", [(1, 197, 256, 3, "max"), (3, 50, 64, 2, "mean"), (2, 33, 32, 3, "max"),
                                          (2, 70, 128, 1, "max"), (1, 41, 96, 4, "mean"), (1, 40, 64, 5, "max"),
                                          (1, 36, 32, 8, "mean")])
def test_fused_embedding_vs_oracle(cuda, B, N, C, k, red):
    from unopose_b200.modules import geo

    pts, dterm, w_d, b_d, w_a, b_a = G.make_inputs(100 + N, B, N, C, cuda)
    fa = 180.0 / (15 * math.pi)
    d, a = geo.geometric_embedding_indices(pts, 0.2, fa, k)
    d_ref, a_ref = G.embedding_indices(pts, 0.2, 15, k)
    bad = _check_indices(d, a, d_ref, a_ref, pts, k)
    # the distance indices are BIT-identical to the reference's GPU path, diagonal (cancellation noise) included: the
    # kernel rounds x.y like cuBLAS (fma chain) and |x|^2 like ATen's GPU reduction, (x0^2 + x2^2) + x1^2
    assert torch.equal(d, d_ref), int((d != d_ref).sum())
    out = geo.geometric_embedding(pts, dterm, w_d, b_d, w_a, b_a, 0.2, fa, k, red)
    # (1) given OUR indices, the oracle's embedding (cuBLAS SGEMM on this device; fp64 on the host as the judge)
    ref = G.embed_from_indices(d, a, dterm, w_d, b_d, w_a, b_a, red)
    assert (out - ref).abs().max() < 2e-5
    ref64 = G.embed_from_indices(d[:1, :8].double().cpu(), a[:1, :8].double().cpu(), dterm.double().cpu(), w_d.double().cpu(),
                                 b_d.double().cpu(), w_a.double().cpu(), b_a.double().cpu(), red)
    e_ours = (out[:1, :8].double().cpu() - ref64).abs().max()
    e_torch = (ref[:1, :8].double().cpu() - ref64).abs().max()
    assert e_ours < 2e-5 and e_ours < 4 * e_torch + 2e-6, (float(e_ours), float(e_torch))
    # (2) end to end against the oracle's own indices, away from the diagonal and from rows with another neighbour set
    full = G.embed_from_indices(d_ref, a_ref, dterm, w_d, b_d, w_a, b_a, red)
    ok = _offdiag(N, cuda).unsqueeze(0) & ~bad.unsqueeze(2)
    assert (out - full)[ok].abs().max() < 1e-4


def test_reference_golden_real_config(cuda):
    from unopose_b200.modules import GeometricStructureEmbedding

    z = np.load(os.path.join(GOLD, "geo_real.npz"))
    N, C, k = int(z["N"]), int(z["C"]), int(z["k"])
    pts, dterm, w_d, b_d, w_a, b_a = G.make_inputs(int(z["seed"]), 1, N, C, cuda)
    m = GeometricStructureEmbedding(dict(sigma_d=0.2, sigma_a=15, angle_k=k, reduction_a=str(z["red"]), hidden_dim=C))
    m.load_state_dict({"embedding.div_term": dterm, "proj_d.weight": w_d, "proj_d.bias": b_d, "proj_a.weight": w_a,
                       "proj_a.bias": b_a})
    m = m.to(cuda).eval()
    from unopose_b200 import _lib
    n0 = _lib.launch_count()
    with torch.no_grad():
        d, a = m.get_embedding_indices(pts)
        out = m(pts)
    assert _lib.launch_count() - n0 == 6          # indices; indices + 2 splits + 2 GEMM phases: the fused path ran
    d_ref, a_ref = torch.from_numpy(z["d_idx"]).to(cuda), torch.from_numpy(z["a_idx"]).to(cuda)
    bad = _check_indices(d, a, d_ref, a_ref, pts, k)
    rows = z["rows"].tolist()
    gold = torch.from_numpy(z["out_rows"]).to(cuda)            # (rows, N, C) of the reference module on CPU
    ok = torch.ones(len(rows), N, dtype=torch.bool, device=cuda)
    for r, i in enumerate(rows):
        ok[r, i] = False
        if bad[0, i]:
            ok[r] = False
    assert ok.float().mean() > 0.6
    assert (out[0, rows] - gold)[ok].abs().max() < 1e-4
    rs = out.double().sum(dim=(2, 3)).cpu().numpy()
    good = ~bad[0].cpu().numpy()
    assert np.allclose(rs[0][good], z["out_rowsum"][0][good], atol=0.05)


def test_module_golden_small_hidden(cuda):
    from unopose_b200.modules import GeometricStructureEmbedding

    g = torch.load(os.path.join(GOLD, "modules_small.pt"), weights_only=False)
    m = GeometricStructureEmbedding(dict(g["cfg_geo"]))
    m.load_state_dict(g["sd_geo"])
    m = m.to(cuda).eval()
    pts = torch.cat([torch.ones(2, 1, 3), g["sp1"]], 1).to(cuda)
    with torch.no_grad():
        out = m(pts)
        torch_path = m.proj_d(m.embedding(G.embedding_indices(pts, 0.2, 15, 3)[0]))   # sanity: module layers on device
    assert torch_path.shape == out.shape
    n = pts.shape[1]
    ok = _offdiag(n, cuda)
    assert (out - g["geo1"].to(cuda))[:, ok].abs().max() < 1e-4


def test_batch_independence_and_unsupported(cuda):
    from unopose_b200 import _lib
    from unopose_b200.modules import geo

    pts, dterm, w_d, b_d, w_a, b_a = G.make_inputs(7, 4, 60, 64, cuda)
    fa = 180.0 / (15 * math.pi)
    full = geo.geometric_embedding(pts, dterm, w_d, b_d, w_a, b_a, 0.2, fa, 3)
    one = geo.geometric_embedding(pts[2:3], dterm, w_d, b_d, w_a, b_a, 0.2, fa, 3)
    assert torch.equal(full[2:3], one)                          # bit-identical: tiles never mix pairs' arithmetic
    assert not geo.supported(48, 3) and not geo.supported(512, 3) and geo.supported(256, 3)
    with pytest.raises(_lib.UnoposeNativeError):
        geo.geometric_embedding(pts, dterm[:12], w_d[:24, :24], b_d[:24], w_a[:24, :24], b_a[:24], 0.2, fa, 3)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        geo.geometric_embedding(pts.cpu(), dterm, w_d, b_d, w_a, b_a, 0.2, fa, 3)
