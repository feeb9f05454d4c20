"""Seeded synthetic inputs of the reference's shapes for tests, smoke() and bench.py.

No datasets or checkpoints exist offline, so clouds are random geometry normalised like the
reference normalises templates (max ||p - mean|| = 1, oneref_feature_extraction.py:272-277) and
"features" are random vectors with a planted correspondence structure.
"""
import numpy as np


def unit_cloud(rng, n):
    """n points ~ uniform in a ball, centred, scaled so max||p - mean|| = 1."""
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    r = rng.random(n) ** (1.0 / 3.0)
    p = v * r[:, None]
    p -= p.mean(0, keepdims=True)
    p /= np.linalg.norm(p, axis=1).max()
    return p.astype(np.float32)


def surface_cloud(rng, n):
    """Points on a bumpy closed surface (closer to a depth-rendered object than a solid ball)."""
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    rad = 0.7 + 0.15 * np.sin(3 * v[:, 0]) * np.cos(2 * v[:, 1]) + 0.1 * v[:, 2] ** 2
    p = v * rad[:, None]
    p -= p.mean(0, keepdims=True)
    p /= np.linalg.norm(p, axis=1).max()
    return p.astype(np.float32)


def batch_clouds(seed, b, n, kind="ball"):
    rng = np.random.default_rng(seed)
    f = unit_cloud if kind == "ball" else surface_cloud
    return np.stack([f(rng, n) for _ in range(b)])


def random_rotation(rng):
    q, r = np.linalg.qr(rng.standard_normal((3, 3)))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def matching_instance(rng, n, c=256, noise=0.01, feat_noise=0.8, non_overlap=0.35, kind="surface"):
    """One query/reference pair with a planted rigid transform and planted feature matches.

    Returns dict: pts1 (n,3) query/observed, pts2 (n,3) reference, f1/f2 (n+1,c) features with the
    shared background token at row 0, score (2n,) in (0.5,1), R (3,3), t (3,) with
    pts1 ~= R pts2[perm] + t (the reference's convention: p_query = R p_ref + t)."""
    cloud = (unit_cloud if kind == "ball" else surface_cloud)(rng, n)
    R = random_rotation(rng).astype(np.float32)
    t = (rng.standard_normal(3) * 0.3).astype(np.float32)
    perm = rng.permutation(n)
    pts1 = cloud[perm] @ R.T + t + rng.standard_normal((n, 3)).astype(np.float32) * noise
    bg = rng.standard_normal(c).astype(np.float32)
    f2 = rng.standard_normal((n, c)).astype(np.float32)
    f1 = f2[perm] + feat_noise * rng.standard_normal((n, c)).astype(np.float32)
    lost = rng.random(n) < non_overlap
    f1[lost] = 0.6 * bg + feat_noise * rng.standard_normal((int(lost.sum()), c)).astype(np.float32)
    f1 = np.concatenate([bg[None], f1], 0)
    f2 = np.concatenate([bg[None], f2], 0)
    score = rng.uniform(0.5, 1.0, 2 * n).astype(np.float32)
    return dict(pts1=pts1.astype(np.float32), pts2=cloud, f1=f1, f2=f2, score=score, R=R, t=t,
                perm=perm, lost=lost)


def matching_batch(seed, b, n, c=256, **kw):
    rng = np.random.default_rng(seed)
    items = [matching_instance(rng, n, c, **kw) for _ in range(b)]
    return {k: np.stack([it[k] for it in items]) for k in items[0]}
