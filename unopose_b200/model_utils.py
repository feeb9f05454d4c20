"""Drop-in for the hot-path functions of ``core.unopose.utils.model_utils`` (reference file).

Same names, argument meaning, return values and RNG consumption as the reference;
the work runs in the sm_100a kernels of libunopose_b200.so through the C ABI
(include/unopose_b200.h).  CUDA tensors only — there is no CPU or plain-torch
fallback for the kernel families (1)-(5).

Pose convention (reference): ``p_ref ~= (p_query - t) @ R``  <=>  ``p_query = R p_ref + t``.
"""
import ctypes

import torch
import torch.nn as nn

from . import _lib as L
from .pointnet2.lrf import LRF  # noqa: F401  (reference model_utils.py:766-823)
from .pointnet2.pointnet2_utils import furthest_point_sample, gather_operation  # noqa: F401


def _f32c(x):
    return x.float().contiguous()


def _need_cuda(x, what):
    if not x.is_cuda:
        raise RuntimeError("%s: CPU not supported (CUDA tensors only; no fallback path)" % what)


def _no_grad_path(what, *tensors):
    """The kernel paths carry no autograd edge: refuse to silently detach."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError("%s: the kernel path has no autograd edge (evaluation only); run under torch.no_grad() "
                           "or detach the inputs" % what)


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


# --------------------------------------------------------------------------- a1
class SimilarityStats:
    """Exponent sums emitted by the similarity GEMM's epilogue for `compute_fine_Rt[_overlap](..., stats=)`:
    valid only for the atten tensor they were produced with (same values, same temp)."""

    __slots__ = ("buf", "temp", "shape")

    def __init__(self, buf, temp, shape):
        self.buf, self.temp, self.shape = buf, float(temp), tuple(shape)


LARGE_GEOMETRY = 512 * 512      # (N1+1)*(N2+1) above which the fine solver runs its streaming passes (assign_geom)


def _padded_atten(b, n, m, device):
    """Storage for a large-geometry `atten`: rows of `ld` floats with 3 pad floats in front, so that element (i, 1) of
    every row is 16-byte aligned and ld % 4 == 0 — the GEMM then stores its tiles with TMA and the assignment passes
    read them with 128-bit loads.  Returns (view (b,n,m) with strides (n*ld, ld, 1), ld); same values, shape and dtype
    as a contiguous tensor, `.contiguous()` gives one."""
    ld = (m + 3 + 3) // 4 * 4
    storage = torch.empty((b, n, ld), dtype=torch.float32, device=device)
    return storage[:, :, 3:3 + m], ld


def _atten_pitch(atten):
    """(tensor, row pitch in floats) the kernels can address: a pitched fp32 view of a large geometry as it is, anything
    else as a contiguous copy."""
    if (atten.dtype == torch.float32 and atten.dim() == 3 and atten.shape[0] > 0 and atten.stride(2) == 1
            and atten.stride(1) >= atten.shape[2] and atten.stride(0) == atten.shape[1] * atten.stride(1)
            and (atten.is_contiguous() or atten.shape[1] * atten.shape[2] > LARGE_GEOMETRY)):
        return atten, atten.stride(1)
    a = atten.float().contiguous()
    return a, a.shape[2]


def compute_feature_similarity(feat1, feat2, type="cosine", temp=1.0, normalize_feat=True, return_stats=False):
    """(B,N,C),(B,M,C) -> (B,N,M).  Reference: model_utils.py:260-282.

    For the fine shape (N*M > 512^2, normalised cosine logits, tensor-core path) the result is a PITCHED view (row
    pitch M+3 rounded up to a multiple of 4 floats, see _padded_atten): same shape, dtype and values as the reference's
    tensor, not `is_contiguous()`.

    return_stats=True (extension): -> (atten, stats); for normalised cosine logits on the tensor-core path the
    GEMM epilogue also emits the exponent sums of the dual-softmax assignment, which
    `compute_fine_Rt[_overlap](atten, ..., stats=stats)` consumes instead of a first pass over `atten`
    (stats is None when the fused path does not apply)."""
    if type not in ("cosine", "L2"):
        raise AssertionError(type)
    _need_cuda(feat1, "compute_feature_similarity")
    _no_grad_path("compute_feature_similarity", feat1, feat2)
    f1, f2 = _f32c(feat1), _f32c(feat2)
    b, n, c = f1.shape
    m = f2.shape[1]
    lib = L.load()
    nbytes = lib.upk_feature_similarity_workspace_bytes(b, n, m, c, int(bool(normalize_feat)))
    ws = _workspace(nbytes, f1.device)
    if type == "cosine" and normalize_feat and b > 0 and n * m > LARGE_GEOMETRY:
        out, ld = _padded_atten(b, n, m, f1.device)
        buf = None
        if return_stats:
            sbytes = lib.upk_similarity_stats_bytes(b, n, m)
            buf = torch.empty(max(int(sbytes) // 4, 1), dtype=torch.float32, device=f1.device)
        for sb in ((buf, None) if buf is not None else (None,)):
            with torch.cuda.device(f1.device):
                rc = lib.upk_feature_similarity_stats_ld(L.ptr(f1), L.ptr(f2), b, n, m, c, float(temp), L.ptr(ws), ws.numel(),
                                                         out.data_ptr(), ld, L.ptr(sb), sb.numel() * 4 if sb is not None else 0,
                                                         L.stream_ptr(f1))
            if rc == 0:
                stats = SimilarityStats(sb, temp, (b, n, m)) if sb is not None else None
                return (out, stats) if return_stats else out
            if rc != -2:                       # anything but "unsupported here": a real failure
                L.check(rc, "feature_similarity_stats_ld")
    out = torch.empty((b, n, m), dtype=torch.float32, device=f1.device)
    with torch.cuda.device(f1.device):
        L.check(lib.upk_feature_similarity(L.ptr(f1), L.ptr(f2), b, n, m, c, float(temp), int(bool(normalize_feat)),
                                           0 if type == "cosine" else 1, L.ptr(ws), ws.numel(), L.ptr(out),
                                           L.stream_ptr(f1)), "feature_similarity")
    return (out, None) if return_stats else out


def pairwise_distance(x, y, normalized=False, channel_first=False):
    """Squared pairwise distances, expansion form, clamped at 0.  Reference: model_utils.py:230-257.

    Host-side torch glue: inside the pose solvers this arithmetic is fused into the
    scoring kernels (families (4)); as a standalone function it only serves the
    geometric embedding (a "next" row), so it stays a plain torch expression."""
    if channel_first:
        xy = torch.matmul(x.transpose(-1, -2), y)
        cdim = -2
    else:
        xy = torch.matmul(x, y.transpose(-1, -2))
        cdim = -1
    if normalized:
        d = 2.0 - 2.0 * xy
    else:
        d = torch.sum(x ** 2, dim=cdim).unsqueeze(-1) - 2 * xy + torch.sum(y ** 2, dim=cdim).unsqueeze(-2)
    return d.clamp(min=0.0)


# --------------------------------------------------------------------------- a7
def _coarse(atten, score, pts1, pts2, model_pts, n_proposal1, n_proposal2, u=None, return_debug=False):
    _need_cuda(pts1, "compute_coarse_Rt")
    B, N1, _ = pts1.shape
    N2 = pts2.shape[1]
    dev = pts1.device
    atten, pts1, pts2 = _f32c(atten), _f32c(pts1), _f32c(pts2)
    n_model = N2
    if model_pts is not None:
        model_pts = _f32c(model_pts)
        n_model = model_pts.shape[1]
    H, K = int(n_proposal1), int(n_proposal2)
    s1 = s2 = None
    ld = 0
    if score is not None:
        score = _f32c(score)
        ld = score.shape[1]
        s1 = score  # score[:, :N1]
        # reference quirk (model_utils.py:440): score2 = score[:, N2:]  — only N2 long when N1 == N2
        if score.shape[1] - N2 != N2:
            raise RuntimeError("compute_coarse_Rt_overlap: score[:, N2:] must have N2 columns "
                               "(the reference slices score[:, N2:], which requires N1 == N2)")
        s2 = score[:, N2:]
    # RNG: exactly one torch.rand(B, 3H) from the default generator, at the reference's position (:462)
    if u is None:
        u = torch.rand(B, H * 3, device=dev)
    u = _f32c(u)
    lib = L.load()
    ws = _workspace(lib.upk_coarse_pose_workspace_bytes(B, N1, N2, H, K), dev)
    R = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    t = torch.empty((B, 3), dtype=torch.float32, device=dev)
    sc = torch.empty((B,), dtype=torch.float32, device=dev)
    pool = torch.empty((B,), dtype=torch.int32, device=dev)
    dbg = None
    dbg_t = None
    if return_debug:
        dbg_t = dict(
            w1=torch.empty((B, N1), device=dev), w2=torch.empty((B, N2), device=dev),
            cdf=torch.empty((B, N1 * N2), device=dev),
            idx1=torch.empty((B, H, 3), dtype=torch.int32, device=dev),
            idx2=torch.empty((B, H, 3), dtype=torch.int32, device=dev),
            Rs=torch.empty((B, H, 3, 3), device=dev), ts=torch.empty((B, H, 3), device=dev),
            resid=torch.empty((B, H), device=dev), top=torch.empty((B, K), dtype=torch.int32, device=dev),
            scores=torch.empty((B, K), device=dev))
        dbg = L.CoarseDebug(**{k: v.data_ptr() for k, v in dbg_t.items()})
    with torch.cuda.device(dev):
        L.check(lib.upk_coarse_pose(
            L.ptr(atten), L.ptr(s1), ld, (s2.data_ptr() if s2 is not None else None), ld,
            L.ptr(pts1), L.ptr(pts2), L.ptr(model_pts), n_model, L.ptr(u), B, N1, N2, H, K,
            L.ptr(ws), ws.numel(), L.ptr(R), L.ptr(t), L.ptr(sc), L.ptr(pool),
            ctypes.addressof(dbg) if dbg is not None else None, L.stream_ptr(pts1)), "coarse_pose")
    if return_debug:
        dbg_t["pool"] = pool
        dbg_t["u"] = u
        return R, t, sc, dbg_t
    return R, t, sc


def compute_coarse_Rt(atten, pts1, pts2, model_pts=None, n_proposal1=6000, n_proposal2=300):
    """Reference: model_utils.py:336-408.  -> (R (B,3,3), t (B,3), pose_score (B,))."""
    return _coarse(atten, None, pts1, pts2, model_pts, n_proposal1, n_proposal2)


def compute_coarse_Rt_overlap(atten, score, pts1, pts2, model_pts=None, n_proposal1=6000, n_proposal2=300):
    """Reference: model_utils.py:411-490 (the variant the model calls, coarse module :99)."""
    return _coarse(atten, score, pts1, pts2, model_pts, n_proposal1, n_proposal2)


# --------------------------------------------------------------------------- a8
def _fine(atten, score, pts1, pts2, model_pts, dis_thres, weight_thresh, return_debug=False, stats=None):
    _need_cuda(pts1, "compute_fine_Rt")
    B, N1 = pts1.shape[:2]
    N2 = pts2.shape[1]
    dev = pts1.device
    pts1, pts2 = _f32c(pts1), _f32c(pts2)
    atten, a_ld = _atten_pitch(atten)
    n_model = N2
    if model_pts is not None:
        model_pts = _f32c(model_pts)
        n_model = model_pts.shape[1]
    s1 = s2 = None
    ld = 0
    if score is not None:
        score = _f32c(score)
        ld = score.shape[1]
        s1 = score
        s2 = score[:, N1:]  # model_utils.py:538
        if s2.shape[1] != N2:
            raise RuntimeError("compute_fine_Rt_overlap: score must be (B, N1 + N2)")
    lib = L.load()
    ws = _workspace(lib.upk_fine_pose_workspace_bytes(B, N1, N2), dev)
    R = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    t = torch.empty((B, 3), dtype=torch.float32, device=dev)
    sc = torch.empty((B,), dtype=torch.float32, device=dev)
    dbg = None
    dbg_t = None
    if return_debug:
        dbg_t = dict(w1=torch.empty((B, N1), device=dev), w2=torch.empty((B, N2), device=dev),
                     soft=torch.empty((B, N1, 3), device=dev), asum=torch.empty((B, N1), device=dev),
                     nn=torch.empty((B, N1), device=dev))
        dbg = L.FineDebug(**{k: v.data_ptr() for k, v in dbg_t.items()})
    tail = (L.ptr(s1), ld, (s2.data_ptr() if s2 is not None else None), ld,
            L.ptr(pts1), L.ptr(pts2), L.ptr(model_pts), n_model, B, N1, N2, float(dis_thres),
            float(weight_thresh), L.ptr(ws), ws.numel(), L.ptr(R), L.ptr(t), L.ptr(sc),
            ctypes.addressof(dbg) if dbg is not None else None, L.stream_ptr(pts1))
    if tuple(atten.shape) != (B, N1 + 1, N2 + 1):
        raise RuntimeError("compute_fine_Rt: atten must be (B, N1 + 1, N2 + 1)")
    with torch.cuda.device(dev):
        if stats is not None:
            if stats.shape != (B, N1 + 1, N2 + 1):
                raise RuntimeError("compute_fine_Rt: stats belong to a different atten tensor")
            L.check(lib.upk_fine_pose_ld(atten.data_ptr(), a_ld, L.ptr(stats.buf), stats.buf.numel() * 4, stats.temp, *tail),
                    "fine_pose_ld")
        else:
            L.check(lib.upk_fine_pose_ld(atten.data_ptr(), a_ld, None, 0, 0.0, *tail), "fine_pose_ld")
    if return_debug:
        return R, t, sc, dbg_t
    return R, t, sc


def compute_fine_Rt(atten, pts1, pts2, model_pts=None, dis_thres=0.15, stats=None):
    """Reference: model_utils.py:493-524 (weight_thresh 0.0).  `stats`: see compute_feature_similarity."""
    return _fine(atten, None, pts1, pts2, model_pts, dis_thres, 0.0, stats=stats)


def compute_fine_Rt_overlap(atten, score, pts1, pts2, model_pts=None, dis_thres=0.15, stats=None):
    """Reference: model_utils.py:527-566 (weight_thresh 0.001; the variant the model calls, fine module :120).
    `stats` (extension): the SimilarityStats returned with THIS atten by compute_feature_similarity(...,
    return_stats=True); the solve then reads atten twice instead of three times."""
    return _fine(atten, score, pts1, pts2, model_pts, dis_thres, 0.001, stats=stats)


def transform_points(pts, R, t):
    """(pts (B,N,3) - t (B,3)) @ R (B,3,3): a cloud moved by a pose, as the fine module does with the coarse pose
    (`p1_ = (p1 - init_t.unsqueeze(1)) @ init_R`, oneref_predator_fine_point_matching.py:65-72).  One kernel
    instead of a broadcast subtract + a K=3 GEMM.  Evaluation path (no autograd edge)."""
    _need_cuda(pts, "transform_points")
    _no_grad_path("transform_points", pts, R, t)
    pts, R, t = _f32c(pts), _f32c(R), _f32c(t)
    B, N = pts.shape[:2]
    out = torch.empty_like(pts)
    with torch.cuda.device(pts.device):
        L.check(L.load().upk_transform_points(L.ptr(pts), L.ptr(R), L.ptr(t), B, N, L.ptr(out), L.stream_ptr(pts)),
                "transform_points")
    return out


# --------------------------------------------------------------------------- a4
def weighted_procrustes(src_points, ref_points, weights=None, weight_thresh=0.0, eps=1e-5,
                        return_transform=False, src_centroid=None, ref_centroid=None):
    """Reference: model_utils.py:667-743.  (N,3)/(B,N,3) inputs; returns (R, t) or a 4x4 transform.
    ``ref ~= R src + t``.  Precomputed centroids (B,3)/(B,1,3) replace the weighted means like in the reference
    (:710-721): the 3x3 covariance is then formed by torch glue exactly like the reference forms it (a given centroid is
    subtracted from its cloud, the weighted mean from the other), the rotation comes from the kernel family's solver
    (`_rotation_from_H`) and ``t = c_ref - R c_src`` uses the given centroids."""
    _need_cuda(src_points, "weighted_procrustes")
    _no_grad_path("weighted_procrustes", src_points, ref_points, weights)
    squeeze = src_points.ndim == 2
    if squeeze:
        src_points = src_points.unsqueeze(0)
        ref_points = ref_points.unsqueeze(0)
        if weights is not None:
            weights = weights.unsqueeze(0)
    src, ref = _f32c(src_points), _f32c(ref_points)
    w = _f32c(weights) if weights is not None else None
    B, N, _ = src.shape
    dev = src.device
    given = src_centroid is not None or ref_centroid is not None
    if given:
        # the reference subtracts the GIVEN centroid from the respective cloud and its own weighted mean from the
        # other: build H exactly like that (torch glue, a 3x3 per instance; no caller in the repository passes
        # centroids) and run the kernel family's solver on it
        wt = w if w is not None else torch.ones_like(src[:, :, 0])
        wt = torch.where(wt < weight_thresh, torch.zeros_like(wt), wt)
        wt = (wt / (wt.sum(1, keepdim=True) + eps)).unsqueeze(2)
        cs = (src * wt).sum(1, keepdim=True) if src_centroid is None else _f32c(src_centroid).reshape(B, 1, 3)
        cr = (ref * wt).sum(1, keepdim=True) if ref_centroid is None else _f32c(ref_centroid).reshape(B, 1, 3)
        H = (src - cs).transpose(1, 2) @ (wt * (ref - cr))
        R = _rotation_from_H(H)
        t = (cr.transpose(1, 2) - R @ cs.transpose(1, 2)).squeeze(2)
    else:
        R = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
        t = torch.empty((B, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if N == 3 and w is None and weight_thresh <= 1.0 and eps == 1e-5:
                # triplet fast path == WeightedProcrustes()(src, ref, None): ref plays "p1", src plays "p2"
                L.check(L.load().upk_kabsch_triplets(L.ptr(ref), L.ptr(src), B, L.ptr(R), L.ptr(t), None,
                                                     L.stream_ptr(src)), "kabsch_triplets")
            else:
                L.check(L.load().upk_weighted_procrustes(L.ptr(src), L.ptr(ref), L.ptr(w), B, N, float(weight_thresh),
                                                         float(eps), L.ptr(R), L.ptr(t), L.stream_ptr(src)),
                        "weighted_procrustes")
    if return_transform:
        T = torch.eye(4, device=dev).unsqueeze(0).repeat(B, 1, 1)
        T[:, :3, :3] = R
        T[:, :3, 3] = t
        return T.squeeze(0) if squeeze else T
    if squeeze:
        return R.squeeze(0), t.squeeze(0)
    return R, t


def _rotation_from_H(H):
    """R = V diag(1, 1, sign det(V U^T)) U^T for H (B,3,3) = U S V^T on the device, with the kernel family's solver:
    the six-point problem src = (+-e_k), ref = (+-H[k, :]) has zero centroids and covariance 2 H / (6 + eps), i.e. the
    same rotation, so upk_weighted_procrustes solves it without any centring effect."""
    B = H.shape[0]
    eye = torch.eye(3, device=H.device, dtype=torch.float32).unsqueeze(0).expand(B, 3, 3)
    src = torch.cat([eye, -eye], 1).contiguous()
    ref = torch.cat([H, -H], 1).float().contiguous()
    R = torch.empty((B, 3, 3), dtype=torch.float32, device=H.device)
    t = torch.empty((B, 3), dtype=torch.float32, device=H.device)
    with torch.cuda.device(H.device):
        L.check(L.load().upk_weighted_procrustes(L.ptr(src), L.ptr(ref), None, B, 6, 0.0, 1e-5, L.ptr(R), L.ptr(t),
                                                 L.stream_ptr(H)), "weighted_procrustes")
    return R


class WeightedProcrustes(nn.Module):
    """Reference: model_utils.py:746-763."""

    def __init__(self, weight_thresh=0.5, eps=1e-5, return_transform=False):
        super().__init__()
        self.weight_thresh = weight_thresh
        self.eps = eps
        self.return_transform = return_transform

    def forward(self, src_points, tgt_points, weights=None, src_centroid=None, ref_centroid=None):
        return weighted_procrustes(src_points, tgt_points, weights=weights, weight_thresh=self.weight_thresh,
                                   eps=self.eps, return_transform=self.return_transform,
                                   src_centroid=src_centroid, ref_centroid=ref_centroid)


# --------------------------------------------------------------------------- a15
class _GatherRows(torch.autograd.Function):
    """x (B,N,C), idx (B,m) int32 -> (B,m,C): one coalesced row copy per sample.  Same values as the
    reference's transpose -> gather_operation -> transpose (model_utils.py:146-149), without the two
    full-tensor transposes."""

    @staticmethod
    def forward(ctx, x, idx):
        L.check_cuda(x, "x")
        L.check_float(x, "x")          # the reference's gather_operation raises "must be a float tensor" too
        L.check_int(idx, "idx")
        x = x.contiguous()
        idx = idx.contiguous()
        b, n, c = x.shape
        m = idx.shape[1]
        ctx.save_for_backward(idx)
        ctx.n_src = n
        out = torch.empty((b, m, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            L.check(L.load().upk_gather_rows(L.ptr(x), L.ptr(idx), b, n, m, c, L.ptr(out), L.stream_ptr(x)),
                    "gather_rows")
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        L.check_float(grad_out, "grad_out")
        g = grad_out.contiguous()
        b, m, c = g.shape
        gx = torch.empty((b, ctx.n_src, c), dtype=torch.float32, device=g.device)
        with torch.cuda.device(g.device):
            L.check(L.load().upk_gather_rows_grad(L.ptr(g), L.ptr(idx), b, ctx.n_src, m, c, L.ptr(gx),
                                                  L.stream_ptr(g)), "gather_rows_grad")
        return gx, None


def gather_rows_pinned(x_host, idx):
    """Zero-copy row gather: x_host is a PINNED host tensor (B,N,C) fp32, idx (B,m) int32 on the GPU ->
    (B,m,C) on the GPU.  The kernel reads the selected rows straight out of the page-locked host memory
    (unified virtual addressing: the host pointer is valid on the device), so only m of the N rows cross
    PCIe.  The host buffer must stay unchanged until the stream has passed this point."""
    if x_host.is_cuda or not x_host.is_pinned():
        raise RuntimeError("gather_rows_pinned: x must be a pinned host tensor")
    if x_host.dtype != torch.float32 or not x_host.is_contiguous():
        raise RuntimeError("gather_rows_pinned: x must be a contiguous float tensor")
    L.check_int(idx, "idx")
    L.check_cuda(idx, "idx")
    idx = idx.contiguous()
    b, n, c = x_host.shape
    m = idx.shape[1]
    out = torch.empty((b, m, c), dtype=torch.float32, device=idx.device)
    with torch.cuda.device(idx.device):
        L.check(L.load().upk_gather_rows(x_host.data_ptr(), L.ptr(idx), b, n, m, c, L.ptr(out), L.stream_ptr(idx)),
                "gather_rows(pinned host source)")
    return out


def _gather_rows(x, idx):
    if not x.is_cuda and x.is_pinned():      # host-fed front end: features stay in pinned host memory
        return gather_rows_pinned(x, idx)
    return _GatherRows.apply(x, idx)


def sample_pts_feats(pts, feats, npoint=2048, return_index=False):
    """FPS + gather of points and features.  Reference: model_utils.py:137-153."""
    with torch.autocast(device_type="cuda", enabled=False):
        pts = pts.to(dtype=torch.float32)
        feats = feats.to(dtype=torch.float32)
        idx = furthest_point_sample(pts.contiguous(), npoint)
        pts, feats = _gather_rows(pts, idx), _gather_rows(feats, idx)
    return (pts, feats, idx) if return_index else (pts, feats)


def sample_pts_feats_wlrf(pts, pts_lrf, feats, npoint=2048, return_index=False):
    """Reference: model_utils.py:156-177."""
    with torch.autocast(device_type="cuda", enabled=False):
        pts = pts.to(dtype=torch.float32)
        pts_lrf = pts_lrf.to(dtype=torch.float32)
        feats = feats.to(dtype=torch.float32)
        idx = furthest_point_sample(pts.contiguous(), npoint)
        pts, pts_lrf, feats = _gather_rows(pts, idx), _gather_rows(pts_lrf, idx), _gather_rows(feats, idx)
    return (pts, pts_lrf, feats, idx) if return_index else (pts, pts_lrf, feats)


def gather_pts_feats(sample_idx, pts, feats):
    """Reference: model_utils.py:180-193."""
    with torch.autocast(device_type="cuda", enabled=False):
        return _gather_rows(pts.float(), sample_idx), _gather_rows(feats.float(), sample_idx)


def gather_pts_feats_wlrf(sample_idx, pts, pts_lrf, feats):
    """Reference: model_utils.py:196-212."""
    with torch.autocast(device_type="cuda", enabled=False):
        return (_gather_rows(pts.float(), sample_idx), _gather_rows(pts_lrf.float(), sample_idx),
                _gather_rows(feats.float(), sample_idx))
