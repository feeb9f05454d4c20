"""CPU: kernel family (3)'s 3x3 Procrustes solver, evaluated on the host from the same source
as the device code (upk_host_procrustes_rotation), against LAPACK SVD."""
import ctypes

import numpy as np

from unopose_b200 import _lib


def _solve(H):
    lib = _lib.load()
    H = np.ascontiguousarray(H, np.float64)
    R = np.empty_like(H)
    rc = lib.upk_host_procrustes_rotation(H.ctypes.data, H.shape[0], R.ctypes.data)
    assert rc == 0
    return R


def _svd_rotation(H):
    U, _, Vt = np.linalg.svd(H)
    V = Vt.transpose(0, 2, 1)
    d = np.sign(np.linalg.det(V @ U.transpose(0, 2, 1)))
    E = np.tile(np.eye(3), (H.shape[0], 1, 1))
    E[:, 2, 2] = d
    return V @ E @ U.transpose(0, 2, 1)


def _angle_deg(Ra, Rb):
    D = Ra.transpose(0, 2, 1) @ Rb
    s = 0.5 * np.sqrt((D[:, 2, 1] - D[:, 1, 2]) ** 2 + (D[:, 0, 2] - D[:, 2, 0]) ** 2 + (D[:, 1, 0] - D[:, 0, 1]) ** 2)
    c = 0.5 * (np.trace(D, axis1=1, axis2=2) - 1)
    return np.degrees(np.arctan2(s, c))


def test_full_rank_matches_svd():
    rng = np.random.default_rng(0)
    H = rng.standard_normal((2000, 3, 3))
    R = _solve(H)
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-12
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-12
    sv = np.linalg.svd(H, compute_uv=False)
    ok = (sv[:, 1] + np.sign(np.linalg.det(H)) * sv[:, 2]) > 1e-3 * sv[:, 0]   # well-conditioned
    assert _angle_deg(R[ok], _svd_rotation(H)[ok]).max() < 1e-6


def test_reflection_case():
    rng = np.random.default_rng(1)
    H = rng.standard_normal((500, 3, 3))
    H[np.linalg.det(H) > 0] *= -1  # force det(H) < 0: plain V U^T would be a reflection
    R = _solve(H)
    assert (np.linalg.det(R) > 0.999999).all()
    sv = np.linalg.svd(H, compute_uv=False)
    ok = (sv[:, 1] - sv[:, 2]) > 1e-2 * sv[:, 0]
    assert _angle_deg(R[ok], _svd_rotation(H)[ok]).max() < 1e-6


def test_rank2_triplets():
    """Every 3-point hypothesis gives a rank-2 H after centring."""
    rng = np.random.default_rng(2)
    n = 3000
    src = rng.standard_normal((n, 3, 3))
    q, r = np.linalg.qr(rng.standard_normal((n, 3, 3)))
    q = q * np.sign(np.linalg.det(q))[:, None, None]
    ref = src @ q.transpose(0, 2, 1) + rng.standard_normal((n, 1, 3))
    sc = src - src.mean(1, keepdims=True)
    rc = ref - ref.mean(1, keepdims=True)
    H = sc.transpose(0, 2, 1) @ rc
    R = _solve(H)
    sv = np.linalg.svd(H, compute_uv=False)
    ok = sv[:, 1] > 1e-3 * sv[:, 0]
    assert _angle_deg(R[ok], q[ok]).max() < 1e-6     # exact recovery of the planted rotation


def test_degenerate_inputs_give_valid_rotations():
    rng = np.random.default_rng(3)
    a = rng.standard_normal((100, 3, 1))
    b = rng.standard_normal((100, 1, 3))
    H = np.concatenate([a @ b, np.zeros((5, 3, 3)), 1e-30 * rng.standard_normal((5, 3, 3))])
    R = _solve(H)
    assert np.isfinite(R).all()
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-9
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-9
    # rank 1: the rotation still maximises tr(R H) = sigma_1
    tr = np.trace(R[:100] @ H[:100], axis1=1, axis2=2)
    assert np.allclose(tr, np.linalg.svd(H[:100], compute_uv=False)[:, 0], rtol=1e-9)


def test_lrf_z_axis_sign_matches_cusolver_recording():
    """The local-reference-frame kernels keep the RAW sign of the least-variance eigenvector when the +-1e-3 sign vote
    ties, like the reference (pointnet2_utils.py:451-456), whose raw sign is whatever torch.svd returns: cuSOLVER's
    batched Jacobi on a GPU.  tests/golden/lrf_svd_sign.npz holds (covariance, V) pairs recorded from torch.svd on a
    B200 (scripts/r2_parity_probe.py: random centres `_all` and centres whose vote ties, at both PositionalEncoding
    scales); the host build of the kernels' solver must return the same direction AND sign as V[:, -1] for every
    covariance that is not singular to fp32 precision."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "lrf_svd_sign.npz"))
    lib = _lib.load()
    checked = 0
    for key in sorted(k[:-4] for k in g.files if k.endswith("_cov")):
        cov, v = g[key + "_cov"].astype(np.float64), g[key + "_v"].astype(np.float64)
        c6 = np.ascontiguousarray(np.stack([cov[:, 0, 0], cov[:, 1, 1], cov[:, 2, 2], cov[:, 0, 1], cov[:, 0, 2],
                                            cov[:, 1, 2]], 1))
        z = np.empty((len(cov), 3))
        assert lib.upk_host_lrf_z_axis(c6.ctypes.data, len(cov), z.ctypes.data) == 0
        w = np.linalg.eigvalsh(cov)
        full = w[:, 0] > 1e-7 * w[:, 2]
        # well-separated smallest eigenvalue: direction defined to fp32 accuracy
        sep = full & ((w[:, 1] - w[:, 0]) > 1e-3 * w[:, 2])
        dot = (z * v[:, :, 2]).sum(1)
        assert (dot[full] > 0).all(), key
        assert (dot[sep] > 1 - 1e-4).all(), key
        checked += int(full.sum())
    assert checked > 1500
