"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

Plain-torch restatement of the reference's correspondence-and-pose path
(core/unopose/utils/model_utils.py).  It is written stage by stage (assignment ->
sampling -> Procrustes -> residual/top-K -> scoring) so the tests can feed each
CUDA stage the *same inputs* the oracle saw; every stage cites the reference
lines it follows and keeps the reference's operation order, so on CPU it is
pinned against golden vectors produced by importing the reference itself
(tests/golden/make_pose_golden.py), and on the B200 box it stands in for the
reference's GPU torch path (same ATen/cuBLAS/cuSOLVER kernels, same Philox draw).

Quirks deliberately preserved: the `score[:, N2:]` slice of
the coarse solver, eps=1e-5 weight normalisation, clamp of overflowing sampled
indices, expansion-form distances, the 1e-8 / 1e-6 guards.
"""
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------- a1 / a9
def feature_similarity(feat1, feat2, sim_type="cosine", temp=1.0, normalize_feat=True):
    """model_utils.py:260-282.  (B,N,C),(B,M,C) -> (B,N,M)."""
    if normalize_feat:
        feat1 = F.normalize(feat1, p=2, dim=2)
        feat2 = F.normalize(feat2, p=2, dim=2)
    if sim_type == "cosine":
        sim = feat1 @ feat2.transpose(1, 2)
    elif sim_type == "L2":
        sim = torch.sqrt(pairwise_sqdist(feat1, feat2, normalized=True))
    else:
        raise AssertionError(sim_type)
    return sim / temp


def pairwise_sqdist(x, y, normalized=False, channel_first=False):
    """model_utils.py:230-257: x^2 - 2xy + y^2 via matmul, clamped at 0 (squared distances)."""
    if channel_first:
        xy = torch.matmul(x.transpose(-1, -2), y)
        cdim = -2
    else:
        xy = torch.matmul(x, y.transpose(-1, -2))
        cdim = -1
    if normalized:
        d = 2.0 - 2.0 * xy
    else:
        x2 = torch.sum(x ** 2, dim=cdim).unsqueeze(-1)
        y2 = torch.sum(y ** 2, dim=cdim).unsqueeze(-2)
        d = x2 - 2 * xy + y2
    return d.clamp(min=0.0)


# --------------------------------------------------------------------------- a2
def dual_softmax_assignment(atten, score1=None, score2=None):
    """model_utils.py:443-457 (coarse) / :537-549 (fine).

    atten (B,N1+1,N2+1) with the background token at row/col 0; score1 (B,N1),
    score2 (B,N2) are the per-point overlap scores (None for the non-overlap
    variants :368-377 / :505-510).  Returns the masked foreground assignment
    (B,N1,N2) and the foreground masks w1 (B,N1), w2 (B,N2)."""
    B, n1p, n2p = atten.shape
    A = torch.softmax(atten, dim=2) * torch.softmax(atten, dim=1)
    if score1 is not None:
        one = torch.ones((B, 1), device=atten.device)
        s1 = torch.cat((one, score1), dim=1)[:, :, None].repeat(1, 1, n2p)
        s2 = torch.cat((one, score2), dim=1)[:, None, :].repeat(1, n1p, 1)
        A = A * s1 * s2
    label1 = torch.max(A[:, 1:, :], dim=2)[1]
    label2 = torch.max(A[:, :, 1:], dim=1)[1]
    w1 = (label1 > 0).float()
    w2 = (label2 > 0).float()
    A = A[:, 1:, 1:].contiguous()
    A = A * w1.unsqueeze(2) * w2.unsqueeze(1)
    return A, w1, w2


# --------------------------------------------------------------------------- a4
def weighted_procrustes(src, ref, weights=None, weight_thresh=0.0, eps=1e-5, src_centroid=None, ref_centroid=None):
    """model_utils.py:667-743 (batched form).  src, ref (B,N,3); weights (B,N)|None; optional precomputed centroids
    (B,3)|(B,1,3) replace the weighted means (:710-721).  Returns R (B,3,3), t (B,3) with  ref ~= R src + t."""
    B = src.shape[0]
    if weights is None:
        weights = torch.ones_like(src[:, :, 0])
    weights = torch.where(torch.lt(weights, weight_thresh), torch.zeros_like(weights), weights)
    weights = weights / (torch.sum(weights, dim=1, keepdim=True) + eps)
    w = weights.unsqueeze(2)
    c_src = torch.sum(src * w, dim=1, keepdim=True) if src_centroid is None else src_centroid.reshape(B, 1, 3)
    c_ref = torch.sum(ref * w, dim=1, keepdim=True) if ref_centroid is None else ref_centroid.reshape(B, 1, 3)
    src_c = src - c_src
    ref_c = ref - c_ref
    H = src_c.permute(0, 2, 1) @ (w * ref_c)
    U, _, V = torch.svd(H)
    Ut = U.transpose(1, 2)
    E = torch.eye(3).unsqueeze(0).repeat(B, 1, 1).to(src.device)
    E[:, -1, -1] = torch.sign(torch.det(V @ Ut))
    R = V @ E @ Ut
    t = c_ref.permute(0, 2, 1) - R @ c_src.permute(0, 2, 1)
    return R, t.squeeze(2)


# --------------------------------------------------------------------------- a3
def sampling_cdf(A_masked):
    """model_utils.py:457-461: P = A**1.5 flattened, cdf = cumsum / (last + 1e-8)."""
    B = A_masked.shape[0]
    P = A_masked.reshape(B, -1) ** 1.5
    cdf = torch.cumsum(P, dim=1)
    cdf = cdf / (cdf[:, -1].unsqueeze(1).contiguous() + 1e-8)
    return cdf


def sample_correspondences(cdf, u, N1, N2):
    """model_utils.py:462-465: searchsorted (left), split into (idx1, idx2), clamp."""
    idx = torch.searchsorted(cdf, u)
    idx1 = torch.clamp(idx.div(N2, rounding_mode="floor"), max=N1 - 1)
    idx2 = torch.clamp(idx % N2, max=N2 - 1)
    return idx1, idx2


# --------------------------------------------------------------------------- a7
def coarse_pose(atten, score, pts1, pts2, model_pts=None, n_proposal1=6000, n_proposal2=300, u=None,
                debug=False):
    """compute_coarse_Rt_overlap (model_utils.py:411-490) when `score` is given,
    compute_coarse_Rt (:336-408) when it is None.

    `u` (B, 3*n_proposal1) injects the uniform draws; by default they are drawn
    here with torch.rand at the reference's RNG consumption point (:462)."""
    B, N1, _ = pts1.shape
    N2 = pts2.shape[1]
    dev = pts1.device
    if model_pts is None:
        model_pts = pts2
    atten, pts1, pts2, model_pts = atten.float(), pts1.float(), pts2.float(), model_pts.float()
    if score is not None:
        s1 = score[:, :N1].float()
        s2 = score[:, N2:].float()  # sic (:440) — equals N1: only because N1 == N2
    else:
        s1 = s2 = None
    A, w1, w2 = dual_softmax_assignment(atten, s1, s2)
    cdf = sampling_cdf(A)
    if u is None:
        u = torch.rand(B, n_proposal1 * 3, device=dev)
    idx1, idx2 = sample_correspondences(cdf, u, N1, N2)
    H, K = n_proposal1, n_proposal2
    p1 = torch.gather(pts1, 1, idx1.unsqueeze(2).repeat(1, 1, 3)).reshape(B * H, 3, 3)
    p2 = torch.gather(pts2, 1, idx2.unsqueeze(2).repeat(1, 1, 3)).reshape(B * H, 3, 3)
    Rs, ts = weighted_procrustes(p2, p1, None, weight_thresh=0.5)  # WeightedProcrustes() default :747
    Rs = Rs.reshape(B, H, 3, 3)
    ts = ts.reshape(B, H, 1, 3)
    p1 = p1.reshape(B, H, 3, 3)
    p2 = p2.reshape(B, H, 3, 3)
    resid = torch.norm((p1 - ts) @ Rs - p2, dim=3).mean(2)  # :475
    top = torch.topk(resid, K, dim=1, largest=False)[1]  # :476
    Rk = torch.gather(Rs, 1, top.reshape(B, K, 1, 1).repeat(1, 1, 3, 3))
    tk = torch.gather(ts, 1, top.reshape(B, K, 1, 1).repeat(1, 1, 1, 3))
    X = ((pts1.unsqueeze(1) - tk) @ Rk).reshape(B * K, -1, 3)  # :481
    M = model_pts.unsqueeze(1).repeat(1, K, 1, 1).reshape(B * K, -1, 3)
    nn = torch.sqrt(pairwise_sqdist(X, M)).min(2)[0].reshape(B, K, -1)  # :483-484
    scores = w1.unsqueeze(1).sum(2) / ((nn * w1.unsqueeze(1)).sum(2) + 1e-8)  # :485
    best_score, best = scores.max(1)
    R = torch.gather(Rk, 1, best.reshape(B, 1, 1, 1).repeat(1, 1, 3, 3)).squeeze(1)
    t = torch.gather(tk, 1, best.reshape(B, 1, 1, 1).repeat(1, 1, 1, 3)).squeeze(2).squeeze(1)
    if not debug:
        return R, t, best_score
    pool = torch.gather(top, 1, best.unsqueeze(1)).squeeze(1)  # SURVEY.md A.8 pool index
    return R, t, best_score, dict(A=A, w1=w1, w2=w2, cdf=cdf, u=u, idx1=idx1, idx2=idx2, Rs=Rs,
                                  ts=ts.squeeze(2), resid=resid, top=top, scores=scores, pool=pool)


# --------------------------------------------------------------------------- a8
def fine_pose(atten, score, pts1, pts2, model_pts=None, dis_thres=0.15, debug=False):
    """compute_fine_Rt_overlap (model_utils.py:527-566) when `score` is given
    (weight_thresh 0.001), compute_fine_Rt (:493-524) when it is None (thresh 0.0)."""
    if model_pts is None:
        model_pts = pts2
    atten, pts1, pts2, model_pts = atten.float(), pts1.float(), pts2.float(), model_pts.float()
    N1 = pts1.shape[1]
    if score is not None:
        s1, s2, thr = score[:, :N1], score[:, N1:], 0.001
    else:
        s1 = s2 = None
        thr = 0.0
    A, w1, _ = dual_softmax_assignment(atten, s1, s2)
    rowsum = A.sum(2)
    soft = (A / (A.sum(2, keepdim=True) + 1e-6)) @ pts2  # :552-553
    R, t = weighted_procrustes(soft, pts1, rowsum, weight_thresh=thr)
    X = (pts1 - t.unsqueeze(1)) @ R
    nn = torch.sqrt(pairwise_sqdist(X, model_pts)).min(2)[0]
    inl = (nn < dis_thres).float()
    s = (inl * w1).sum(1) / (w1.sum(1) + 1e-8)
    s = s * w1.mean(1)
    if not debug:
        return R, t, s
    return R, t, s, dict(A=A, w1=w1, rowsum=rowsum, soft=soft, nn=nn)


# --------------------------------------------------------------------------- metrics
def rotation_geodesic_deg(Ra, Rb):
    """Geodesic angle (degrees) between batches of rotations, computed in float64."""
    Ra, Rb = Ra.double(), Rb.double()
    D = Ra.transpose(-1, -2) @ Rb
    # robust for tiny angles: use the skew part (sin) together with the trace (cos)
    tr = D.diagonal(dim1=-2, dim2=-1).sum(-1)
    skew = 0.5 * (D - D.transpose(-1, -2))
    s = torch.sqrt(skew[..., 2, 1] ** 2 + skew[..., 0, 2] ** 2 + skew[..., 1, 0] ** 2)
    c = 0.5 * (tr - 1.0)
    return torch.rad2deg(torch.atan2(s, c))


def relative_translation_error(ta, tb):
    ta, tb = ta.double(), tb.double()
    return (ta - tb).norm(dim=-1) / tb.norm(dim=-1).clamp_min(1e-12)
