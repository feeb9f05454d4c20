"""GPU parity tests of the pointnet2 kernels (through the C ABI, via the
drop-in ``_ext`` module) against (1) the C oracle, (2) the reference's own
extension compiled unmodified (oracle/_ref, when present on the box) and
(3) the committed golden vectors.  Bit-exact everywhere (integer/index work;
the float outputs are gathers or the same fused arithmetic)."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import pointnet2_oracle as O
from oracle import ref_ext
from util_clouds import batch_clouds

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ext():
    from unopose_b200.pointnet2 import _ext

    return _ext


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("n,m", [(1, 1), (2, 2), (3, 3), (31, 7), (196, 64), (255, 200), (256, 33), (500, 100),
                                 (777, 64), (1024, 128), (2048, 196), (2500, 300), (4096, 50), (5000, 2048),
                                 (7000, 40), (10000, 30), (14000, 20),
                                 # long chains (m >= 512) at cloud sizes around the per-thread point-count switches
                                 (2048, 2048), (2049, 512), (3000, 600), (4096, 1000), (5120, 700), (8000, 513),
                                 (8192, 600), (8193, 600)])
def test_fps_matches_oracle(cuda, n, m):
    xyz = batch_clouds(n * 7 + m, 3, n)
    got = _ext().furthest_point_sampling(T(xyz, cuda), m).cpu().numpy()
    assert got.dtype == np.int32
    assert np.array_equal(got, O.furthest_point_sampling(xyz, m))


@pytest.mark.parametrize("n", [196, 600, 2048, 5000])
def test_fps_ties_match_oracle(cuda, n):
    """Quantised coordinates + duplicated points + more samples than distinct points."""
    xyz = np.round(batch_clouds(n, 2, n) * 4) / 4
    xyz[:, n // 2:] = xyz[:, : n - n // 2]
    for m in (min(n, 400), min(n, 1500)):   # more samples than distinct points: every later pick is an exact tie
        got = _ext().furthest_point_sampling(T(xyz, cuda), m).cpu().numpy()
        assert np.array_equal(got, O.furthest_point_sampling(xyz, m))
    z = np.zeros((1, n, 3), np.float32)
    assert (_ext().furthest_point_sampling(T(z, cuda), 9).cpu().numpy() == 0).all()


@pytest.mark.parametrize("n,m,r,ns", [(2048, 2048, 0.1, 64), (2048, 2048, 0.2, 256), (300, 77, 0.25, 16),
                                      (5000, 196, 0.15, 32), (9000, 40, 0.3, 500), (50, 50, 10.0, 64),
                                      (33, 65, 0.01, 5), (4097, 31, 0.5, 1)])
def test_ball_query_matches_oracle(cuda, n, m, r, ns):
    xyz = batch_clouds(n + m, 2, n, "surface")
    q = xyz[:, :m] if m <= n else batch_clouds(5, 2, m)
    q = np.ascontiguousarray(q)
    got = _ext().ball_query(T(q, cuda), T(xyz, cuda), r, ns).cpu().numpy()
    assert np.array_equal(got, O.ball_query(q, xyz, r, ns))


def test_ball_query_no_hit_rows_zero(cuda):
    xyz = batch_clouds(3, 1, 100)
    q = xyz[:, :10] + 50.0
    got = _ext().ball_query(T(q, cuda), T(xyz, cuda), 0.1, 8)
    assert (got == 0).all()


@pytest.mark.parametrize("n,m,scales", [
    (2048, 2048, [(0.1, 64), (0.2, 256)]),      # the PositionalEncoding configuration
    (2048, 2048, [(0.2, 256), (0.1, 64)]),      # larger radius first
    (2048, 2048, [(0.2, 16), (0.1, 64)]),       # outer row saturates before the inner one
    (2048, 2048, [(0.15, 32)]),
    (300, 77, [(0.25, 16), (0.25, 3)]),         # equal radii
    (5000, 196, [(0.15, 32), (0.3, 100)]),      # three smem tiles
    (9000, 40, [(0.3, 500), (0.05, 7)]),
    (50, 50, [(10.0, 64), (0.01, 5)]),
    (33, 65, [(0.01, 5)]),
    (4097, 31, [(0.5, 1), (0.6, 2)]),
])
def test_ball_query_group_fused_matches_oracle(cuda, n, m, scales):
    xyz = batch_clouds(n + m + 1, 2, n, "surface")
    q = xyz[:, :m] if m <= n else batch_clouds(5, 2, m)
    q = np.ascontiguousarray(q)
    xyz_cf = np.ascontiguousarray(xyz.transpose(0, 2, 1))
    for group in (True, False):
        outs = _ext().ball_query_group(T(q, cuda), T(xyz, cuda), scales, group=group)
        assert len(outs) == len(scales)
        for (r, ns), (idx, g) in zip(scales, outs):
            exp = O.ball_query(q, xyz, r, ns)
            assert np.array_equal(idx.cpu().numpy(), exp), (r, ns)
            if group:
                assert np.array_equal(g.cpu().numpy(), O.group_points(xyz_cf, exp)), (r, ns)
            else:
                assert g is None


def test_ball_query_group_fused_no_hit_and_duplicates(cuda):
    xyz = batch_clouds(3, 1, 100)
    q = np.ascontiguousarray(xyz[:, :10] + 50.0)
    (i0, g0), (i1, g1) = _ext().ball_query_group(T(q, cuda), T(xyz, cuda), [(0.1, 8), (0.3, 4)])
    assert (i0 == 0).all() and (i1 == 0).all()
    p0 = torch.from_numpy(xyz[:, 0]).to(cuda)          # zero rows group point 0
    assert torch.equal(g0, p0[:, :, None, None].expand_as(g0)) and torch.equal(g1, p0[:, :, None, None].expand_as(g1))
    # duplicated points: every copy is a hit, ascending index order
    d = np.ascontiguousarray(np.repeat(batch_clouds(4, 2, 40), 3, axis=1))
    (idx, g), = _ext().ball_query_group(T(d, cuda), T(d, cuda), [(1e-3, 5)])
    assert np.array_equal(idx.cpu().numpy(), O.ball_query(d, d, 1e-3, 5))


@pytest.mark.parametrize("c,n,npoints,ns", [(3, 2048, 2048, 64), (3, 2048, 2048, 256), (5, 100, 7, 3), (256, 500, 64, 1),
                                            (1, 10, 1, 1)])
def test_group_and_grad_match_oracle(cuda, c, n, npoints, ns):
    rng = np.random.default_rng(c + n)
    pts = rng.standard_normal((2, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (2, npoints, ns)).astype(np.int32)
    got = _ext().group_points(T(pts, cuda), T(idx, cuda)).cpu().numpy()
    assert np.array_equal(got, O.group_points(pts, idx))
    go = rng.integers(-4, 5, (2, c, npoints, ns)).astype(np.float32)  # integers: atomics order-free
    gg = _ext().group_points_grad(T(go, cuda), T(idx, cuda), n).cpu().numpy()
    assert np.array_equal(gg, O.group_points_grad(go, idx, n))


@pytest.mark.parametrize("c,n,m", [(3, 2048, 196), (256, 2049, 196), (3, 5000, 2048), (256, 5000, 2048), (2, 9, 5)])
def test_gather_and_grad_match_oracle(cuda, c, n, m):
    rng = np.random.default_rng(c + n + m)
    pts = rng.standard_normal((2, c, n)).astype(np.float32)
    idx = rng.integers(0, n, (2, m)).astype(np.int32)
    got = _ext().gather_points(T(pts, cuda), T(idx, cuda)).cpu().numpy()
    assert np.array_equal(got, O.gather_points(pts, idx))
    go = rng.integers(-4, 5, (2, c, m)).astype(np.float32)
    gg = _ext().gather_points_grad(T(go, cuda), T(idx, cuda), n).cpu().numpy()
    assert np.array_equal(gg, O.gather_points_grad(go, idx, n))


def test_three_nn_interpolate_match_oracle(cuda):
    rng = np.random.default_rng(9)
    unknown = batch_clouds(1, 2, 333)
    known = batch_clouds(2, 2, 1500)
    d2, idx = _ext().three_nn(T(unknown, cuda), T(known, cuda))
    od2, oidx = O.three_nn(unknown, known)
    assert np.array_equal(idx.cpu().numpy(), oidx) and np.array_equal(d2.cpu().numpy(), od2)
    feats = rng.standard_normal((2, 6, 1500)).astype(np.float32)
    w = rng.random((2, 333, 3)).astype(np.float32)
    out = _ext().three_interpolate(T(feats, cuda), idx, T(w, cuda)).cpu().numpy()
    assert np.array_equal(out, O.three_interpolate(feats, oidx, w))
    go = rng.integers(-3, 4, (2, 6, 333)).astype(np.float32)
    wi = rng.integers(0, 3, (2, 333, 3)).astype(np.float32)
    gg = _ext().three_interpolate_grad(T(go, cuda), idx, T(wi, cuda), 1500).cpu().numpy()
    assert np.array_equal(gg, O.three_interpolate_grad(go, oidx, wi, 1500))


@pytest.mark.skipif(not ref_ext.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("n,m,kind", [(2048, 196, "ball"), (5000, 2048, "surface"), (196, 100, "ball"), (1000, 333, "surface")])
def test_against_reference_extension(cuda, n, m, kind):
    """The reference's own _ext (compiled unmodified for sm_100a) on the same inputs."""
    ref = ref_ext.load()
    xyz = T(batch_clouds(n, 4, n, kind), cuda)
    assert torch.equal(_ext().furthest_point_sampling(xyz, m), ref.furthest_point_sampling(xyz, m))
    for r, ns in ((0.1, 64), (0.2, 256)):
        mine = _ext().ball_query(xyz, xyz, r, ns)
        theirs = ref.ball_query(xyz, xyz, r, ns)
        assert torch.equal(mine, theirs)
        f = xyz.transpose(1, 2).contiguous()
        assert torch.equal(_ext().group_points(f, mine), ref.group_points(f, theirs))
    idx = ref.furthest_point_sampling(xyz, m)
    feats = torch.randn(4, 256, n, device=cuda)
    assert torch.equal(_ext().gather_points(feats, idx), ref.gather_points(feats, idx))
    q = xyz[:, : n // 2].contiguous()
    d2m, im = _ext().three_nn(q, xyz)
    d2r, ir = ref.three_nn(q, xyz)
    assert torch.equal(im, ir) and torch.equal(d2m, d2r)
    w = torch.rand(4, n // 2, 3, device=cuda)
    assert torch.equal(_ext().three_interpolate(feats, im, w), ref.three_interpolate(feats, ir, w))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pointnet2_*.npz"))))
def test_against_golden(cuda, path):
    g = np.load(path)
    xyz = T(g["xyz"], cuda)
    e = _ext()
    assert np.array_equal(e.furthest_point_sampling(xyz, g["fps_idx"].shape[1]).cpu().numpy(), g["fps_idx"])
    for tag in ("bq1", "bq2"):
        got = e.ball_query(xyz, xyz, float(g[tag + "_radius"]), g[tag + "_idx"].shape[2]).cpu().numpy()
        assert np.array_equal(got, g[tag + "_idx"])
    assert np.array_equal(e.group_points(T(g["feat"], cuda), T(g["bq1_idx"], cuda)).cpu().numpy(), g["grouped"])
    assert np.array_equal(e.gather_points(T(g["feat"], cuda), T(g["fps_idx"], cuda)).cpu().numpy(), g["gathered"])


def test_error_conventions(cuda):
    e = _ext()
    with pytest.raises(RuntimeError, match="CPU not supported"):
        e.furthest_point_sampling(torch.zeros(1, 8, 3), 2)
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        e.furthest_point_sampling(torch.zeros(1, 3, 8, device=cuda).transpose(1, 2), 2)
    with pytest.raises(RuntimeError, match="must be an int tensor"):
        e.gather_points(torch.zeros(1, 3, 8, device=cuda), torch.zeros(1, 2, dtype=torch.int64, device=cuda))
    with pytest.raises(RuntimeError, match="must be a float tensor"):
        e.ball_query(torch.zeros(1, 3, 3, device=cuda, dtype=torch.float64), torch.zeros(1, 3, 3, device=cuda), 0.1, 4)


def test_python_wrappers_autograd(cuda):
    from unopose_b200.pointnet2 import pointnet2_utils as P

    xyz = T(batch_clouds(4, 2, 512), cuda)
    idx = P.furthest_point_sample(xyz, 64)
    assert idx.dtype == torch.int32 and not idx.requires_grad
    feats = torch.randn(2, 8, 512, device=cuda, requires_grad=True)
    out = P.gather_operation(feats, idx)
    out.sum().backward()
    assert feats.grad is not None and feats.grad.sum().item() == pytest.approx(out.numel())
    bq = P.ball_query(0.3, 16, xyz, xyz)
    g = P.grouping_operation(feats, bq)
    assert g.shape == (2, 8, 512, 16)
    qg = P.QueryAndGroup(0.3, 16)
    assert qg(xyz, xyz).shape == (2, 3, 512, 16)
    ql = P.QueryAndLRFGroup(0.3, 16, use_xyz=True)
    assert ql(xyz, xyz, xyz.transpose(1, 2).contiguous()).shape == (2, 6, 512, 16)


def test_ball_query_group_randomised_shapes(cuda):
    """40 seeded random (n, m, radii, nsamples, cloud kind) cases of the fused kernel against the C oracle."""
    rng = np.random.default_rng(2024)
    for case in range(40):
        n = int(rng.integers(1, 5200))
        m = int(rng.integers(1, 700))
        kind = "surface" if case % 2 else "ball"
        nsc = int(rng.integers(1, 3))
        scales = [(float(rng.uniform(0.02, 0.6)), int(rng.integers(1, 300))) for _ in range(nsc)]
        xyz = batch_clouds(1000 + case, 2, n, kind)
        if case % 5 == 0:                                    # quantised coordinates: many exact distance ties at r
            xyz = (np.round(xyz * 8) / 8).astype(np.float32)
            scales = [(0.125 * int(rng.integers(1, 4)), ns) for _, ns in scales]
        q = np.ascontiguousarray(xyz[:, rng.integers(0, n, m)]) if case % 3 else np.ascontiguousarray(
            batch_clouds(7 + case, 2, m, "ball"))
        outs = _ext().ball_query_group(T(q, cuda), T(xyz, cuda), scales)
        xyz_cf = np.ascontiguousarray(xyz.transpose(0, 2, 1))
        for (r, ns), (idx, g) in zip(scales, outs):
            exp = O.ball_query(q, xyz, r, ns)
            assert np.array_equal(idx.cpu().numpy(), exp), (case, n, m, r, ns)
            assert np.array_equal(g.cpu().numpy(), O.group_points(xyz_cf, exp)), (case, n, m, r, ns)


@pytest.mark.gpu
def test_fps_thread_configurations_bit_identical(tmp_path):
    """UPK_FPS_CFG=5 / 6 run the FPS on 8 / 4 warps per instance with the general tie key (bit reversal of k mod BS
    evaluated for the winning point instead of per thread): the indices must equal the default configuration's bit for
    bit, including on quantised clouds with many exact distance ties (sampling_gpu.cu:64-70 tournament order)."""
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    child = os.path.join(root, "scripts", "dev", "fps_cfg_sweep.py")
    outs = {}
    for cfg in ("0", "5", "6"):
        p = str(tmp_path / ("fps_%s.pt" % cfg))
        r = subprocess.run([_sys.executable, child, "--child", p], env=dict(os.environ, UPK_FPS_CFG=cfg),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        outs[cfg] = torch.load(p)
    for cfg in ("5", "6"):
        for k, v in outs["0"].items():
            assert torch.equal(v, outs[cfg][k]), (cfg, k)
