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


def _small_rotation(rng, max_deg):
    ax = rng.standard_normal(3)
    ax /= np.linalg.norm(ax)
    ang = np.deg2rad(max_deg) * rng.uniform(-1, 1)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def forward_batch(seed, b, n_query=2048, n_template=5000, img=224, radius_m=0.08, depth_m=0.8, noise_m=0.0005,
                  view_jitter_deg=3.0, splat=2):
    """Inputs of a full `UNOPose.forward` (BASELINE.json config 3; keys and shapes of the reference's test loader,
    SURVEY.md §3.1): YCB-V-sized objects (radius ~8 cm) seen at ~0.8 m, as RGB-D crops.

    Geometry.  The object is a bumpy closed surface.  The template cloud `tem1_pts` (object frame, metres) holds the
    `n_template` surface points that face the template camera; the observed cloud `pts` (camera frame, metres) is a
    rigidly moved, noisy subset of those that also face the query camera: `pts ~= R tem1_pts[sel] + t`.  The template
    view is within `view_jitter_deg` of the query view (UNOPose's setting: ONE reference view of the unseen object that
    shares visible surface with the query).

    Appearance.  Each point owns a pixel of its 224x224 crop (orthographic, nearest pixel, the crop spans the cloud's
    bounding box) and is painted as a (2*splat+1)^2 footprint whose colour is a smooth function of the OBJECT-space
    position, so the two crops show the same textured object from nearby views and corresponding points carry similar
    image content — what a feature extractor needs to produce matchable features.

    Returns numpy arrays: rgb / tem1_rgb (b,3,img,img) f32, rgb_choose (b,n_query) / tem1_choose (b,n_template) i64,
    pts (b,n_query,3), tem1_pts (b,n_template,3), R (b,3,3), t (b,3)."""
    rng = np.random.default_rng(seed)
    out = {k: [] for k in ("rgb", "rgb_choose", "pts", "tem1_rgb", "tem1_choose", "tem1_pts", "R", "t")}
    freq, phase = np.array([2.1, 2.7, 3.3]), np.array([0.3, 1.1, 2.0])

    def crop(points_view, points_obj):
        xy = points_view[:, :2]
        lo, hi = xy.min(0), xy.max(0)
        uv = np.clip(((xy - lo) / (hi - lo + 1e-9) * (img - 1 - 2 * splat)).round().astype(np.int64) + splat, 0, img - 1)
        # "mean/std-normalised" RGB (zero-centred), a multi-band texture fixed to the object's surface
        q = points_obj / radius_m
        col = (np.sin(q * freq + phase) * np.sin(q[:, [1, 2, 0]] * (3.1 * freq) + 2 * phase)
               + 0.5 * np.sin(q[:, [2, 0, 1]] * (5.3 * freq) + 3 * phase)).astype(np.float32)
        image = np.zeros((img, img, 3), np.float32)
        for dy in range(-splat, splat + 1):
            for dx in range(-splat, splat + 1):
                image[np.clip(uv[:, 1] + dy, 0, img - 1), np.clip(uv[:, 0] + dx, 0, img - 1)] = col
        return uv[:, 1] * img + uv[:, 0], image.transpose(2, 0, 1).copy()

    for _ in range(b):
        R = random_rotation(rng).astype(np.float64)
        V = _small_rotation(rng, view_jitter_deg) @ R                     # object -> template camera
        v = rng.standard_normal((6 * n_template, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        v = v[(v @ V.T)[:, 2] < 0.1][:n_template]                         # faces the template camera (looks down +z)
        assert len(v) == n_template
        rad = 0.7 + 0.15 * np.sin(3 * v[:, 0]) * np.cos(2 * v[:, 1]) + 0.1 * v[:, 2] ** 2
        tem = v * rad[:, None]
        tem *= radius_m / np.linalg.norm(tem - tem.mean(0), axis=1).max()
        t = np.array([rng.normal(0, 0.05), rng.normal(0, 0.05), depth_m + rng.normal(0, 0.05)])
        seen = np.flatnonzero((v @ R.T)[:, 2] < 0.05)                     # also faces the query camera
        sel = rng.choice(seen, n_query, replace=len(seen) < n_query)
        pts = tem[sel] @ R.T + t + rng.standard_normal((n_query, 3)) * noise_m
        tem_choose, tem_rgb = crop(tem @ V.T, tem)
        choose, rgb = crop(pts, tem[sel])
        for k, val in (("rgb", rgb), ("rgb_choose", choose), ("pts", pts.astype(np.float32)), ("tem1_rgb", tem_rgb),
                       ("tem1_choose", tem_choose), ("tem1_pts", tem.astype(np.float32)), ("R", R.astype(np.float32)),
                       ("t", t.astype(np.float32))):
            out[k].append(val)
    return {k: np.stack(v) for k, v in out.items()}
