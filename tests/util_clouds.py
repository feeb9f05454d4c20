"""Seeded synthetic clouds shared by tests, smoke() and bench.py (SURVEY.md §8d generator)."""
import numpy as np


def unit_cloud(rng, n):
    """n points ~ uniform in a ball, centred and scaled so max||p - mean|| = 1
    (mirrors the reference's radius normalisation, oneref_feature_extraction.py:272-277)."""
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
