"""CPU tests of the pointnet2 oracle (oracle/pointnet2_oracle.c): known-answer
invariants and the committed golden vectors that were produced
on a B200 by the reference's own extension compiled unmodified
(tests/golden/make_pointnet2_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import pointnet2_oracle as O
from util_clouds import batch_clouds

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fps_invariants():
    x = batch_clouds(0, 3, 500)
    idx = O.furthest_point_sampling(x, 64)
    assert idx.shape == (3, 64) and idx.dtype == np.int32
    assert (idx[:, 0] == 0).all()  # sampling_gpu.cu:90-91
    for b in range(3):
        assert len(set(idx[b].tolist())) == 64  # distinct points on a generic cloud
    # second pick is the farthest point from point 0
    d = ((x - x[:, :1]) ** 2).sum(-1)
    assert (idx[:, 1] == d.argmax(1)).all()


def test_fps_small_known_answer():
    # 4 collinear points: 0, 1, 3, 10 -> picks 0, then 10 (idx 3), then 3 (idx 2), then 1
    x = np.array([[[0, 0, 0], [1, 0, 0], [3, 0, 0], [10, 0, 0]]], np.float32)
    assert O.furthest_point_sampling(x, 4).tolist() == [[0, 3, 2, 1]]


def test_fps_tie_order_follows_tree():
    # all points identical -> every distance ties at 0: thread 0's first point wins (index 0)
    x = np.zeros((1, 300, 3), np.float32)
    assert (O.furthest_point_sampling(x, 5) == 0).all()
    # duplicates of the farthest point: indices 5 and 133 are the same far point, block size 256;
    # the tournament prefers bit-reversed-smaller thread id: bitrev8(5)=160, bitrev8(133)=161 -> 5
    x = np.zeros((1, 300, 3), np.float32)
    x[0, 5] = x[0, 133] = (1, 0, 0)
    assert O.furthest_point_sampling(x, 2)[0, 1] == 5
    # indices 6 (bitrev8=96) and 129 (bitrev8=129): 6 wins; 3 (192) vs 128 (1): 128 wins
    x = np.zeros((1, 300, 3), np.float32)
    x[0, 3] = x[0, 128] = (1, 0, 0)
    assert O.furthest_point_sampling(x, 2)[0, 1] == 128
    # same thread (k mod 256 equal): smaller k wins: 7 and 263
    x = np.zeros((1, 300, 3), np.float32)
    x[0, 7] = x[0, 263] = (1, 0, 0)
    assert O.furthest_point_sampling(x, 2)[0, 1] == 7


def test_ball_query_invariants():
    x = batch_clouds(1, 2, 300)
    idx = O.ball_query(x, x, 0.25, 16)
    assert idx.shape == (2, 300, 16)
    # rows are non-decreasing up to the padding, padded with the first hit; query itself is a hit
    for b in range(2):
        for j in range(300):
            row = idx[b, j]
            d2 = ((x[b] - x[b, j]) ** 2).sum(-1)
            hits = np.nonzero(d2 < np.float32(0.25) ** 2 - 1e-6)[0]
            k = min(len(hits), 16)
            assert j in row.tolist() or len(hits) > 16
            assert (np.diff(row[:k]) > 0).all()
            assert (row[k:] == row[0]).all()


def test_ball_query_no_hit_row_is_zero():
    xyz = np.array([[[5, 5, 5], [6, 6, 6]]], np.float32)
    q = np.array([[[0, 0, 0]]], np.float32)
    assert (O.ball_query(q, xyz, 0.1, 4) == 0).all()


def test_group_gather_roundtrip():
    rng = np.random.default_rng(2)
    pts = rng.standard_normal((2, 5, 40)).astype(np.float32)
    idx = rng.integers(0, 40, (2, 7, 3)).astype(np.int32)
    g = O.group_points(pts, idx)
    assert g.shape == (2, 5, 7, 3)
    assert g[1, 4, 6, 2] == pts[1, 4, idx[1, 6, 2]]
    ga = O.gather_points(pts, idx[:, :, 0].copy())
    assert np.array_equal(ga, g[..., 0])
    # gradient of a gather is a histogram-weighted scatter
    go = np.ones((2, 5, 7), np.float32)
    gg = O.gather_points_grad(go, idx[:, :, 0].copy(), 40)
    assert gg.sum() == go.sum()


def test_three_nn_and_interpolate():
    rng = np.random.default_rng(3)
    u = rng.standard_normal((1, 10, 3)).astype(np.float32)
    k = rng.standard_normal((1, 20, 3)).astype(np.float32)
    d2, idx = O.three_nn(u, k)
    full = ((u[0, :, None] - k[0, None]) ** 2).sum(-1)
    assert np.array_equal(idx[0], np.argsort(full, 1, kind="stable")[:, :3].astype(np.int32))
    np.testing.assert_allclose(d2[0], np.sort(full, 1)[:, :3], rtol=1e-5, atol=1e-6)
    feats = rng.standard_normal((1, 4, 20)).astype(np.float32)
    w = rng.random((1, 10, 3)).astype(np.float32)
    out = O.three_interpolate(feats, idx, w)
    ref = (feats[0][:, idx[0]] * w[0][None]).sum(-1)
    np.testing.assert_allclose(out[0], ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "pointnet2_*.npz"))))
def test_oracle_matches_reference_ext_golden(path):
    """Golden vectors = outputs of the reference's own _ext on a B200 (bit-exact)."""
    g = np.load(path)
    xyz = g["xyz"]
    assert np.array_equal(O.furthest_point_sampling(xyz, g["fps_idx"].shape[1]), g["fps_idx"])
    for tag in ("bq1", "bq2"):
        r, ns = float(g[tag + "_radius"]), int(g[tag + "_idx"].shape[2])
        assert np.array_equal(O.ball_query(xyz, xyz, r, ns), g[tag + "_idx"])
    assert np.array_equal(O.group_points(g["feat"], g["bq1_idx"]), g["grouped"])
    assert np.array_equal(O.gather_points(g["feat"], g["fps_idx"]), g["gathered"])
    d2, i3 = O.three_nn(g["unknown"], xyz)
    assert np.array_equal(i3, g["nn_idx"]) and np.array_equal(d2, g["nn_dist2"])
    assert np.array_equal(O.three_interpolate(g["feat"], g["nn_idx"], g["nn_w"]), g["interp"])
