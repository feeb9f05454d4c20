"""Drop-in for ``core.unopose.model.pointnet2.pointnet2_utils`` (reference file of the same name).

Public names, call signatures, dtypes and autograd behaviour follow the
reference (file:line cited per item); the work is done by the sm_100a kernels
behind ``unopose_b200.pointnet2._ext``.
"""
import os

import torch
import torch.nn as nn
from torch.autograd import Function

from . import _ext
from .lrf import LRF_batch  # noqa: F401  (re-exported, reference pointnet2_utils.py:429-481)


class FurthestPointSampling(Function):
    """xyz (B,N,3) f32, npoint -> (B,npoint) int32; start index 0.  Ref: pointnet2_utils.py:51-80."""

    @staticmethod
    def forward(ctx, xyz, npoint):
        inds = _ext.furthest_point_sampling(xyz, npoint)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, grad=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    """features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint).  Ref: pointnet2_utils.py:83-117."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n_src = features.size(2)
        return _ext.gather_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.gather_points_grad(grad_out.contiguous(), idx, ctx.n_src), None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) L2 (sqrt), idx (B,n,3) int32).
    Ref: pointnet2_utils.py:120-148."""

    @staticmethod
    def forward(ctx, unknown, known):
        dist2, idx = _ext.three_nn(unknown, known)
        ctx.mark_non_differentiable(idx)
        return torch.sqrt(dist2), idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    """features (B,c,m), idx (B,n,3), weight (B,n,3) -> (B,c,n).  Ref: pointnet2_utils.py:151-204."""

    @staticmethod
    def forward(ctx, features, idx, weight):
        ctx.save_for_backward(idx, weight)
        ctx.m_src = features.size(2)
        return _ext.three_interpolate(features, idx, weight)

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        g = _ext.three_interpolate_grad(grad_out.contiguous(), idx, weight, ctx.m_src)
        return g, None, None


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    """features (B,C,N), idx (B,npoint,nsample) int32 -> (B,C,npoint,nsample).
    Ref: pointnet2_utils.py:207-255."""

    @staticmethod
    def forward(ctx, features, idx):
        ctx.save_for_backward(idx)
        ctx.n_src = features.size(2)
        return _ext.group_points(features, idx)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _ext.group_points_grad(grad_out.contiguous(), idx, ctx.n_src), None


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    """(radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> (B,npoint,nsample) int32.
    Ref: pointnet2_utils.py:258-289 (note: the native call takes new_xyz first)."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        inds = _ext.ball_query(new_xyz, xyz, radius, nsample)
        ctx.mark_non_differentiable(inds)
        return inds

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None, None


ball_query = BallQuery.apply


def _fusable(xyz, sample_uniformly):
    # the fused kernel has no autograd edge to xyz; the unfused sequence keeps the reference's
    # GroupingOperation.backward for callers that differentiate through the grouped coordinates
    return xyz.is_cuda and not sample_uniformly and not (torch.is_grad_enabled() and xyz.requires_grad)


def ball_query_and_group(xyz, new_xyz, scales):
    """[(radius, nsample)] x1 or x2 -> [(idx (B,npoint,nsample) int32, grouped_xyz (B,3,npoint,nsample))]:
    ball_query + grouping_operation(xyz^T, idx) per scale in one fused scan (_ext.ball_query_group).
    Non-differentiable (evaluation path)."""
    with torch.no_grad():
        return _ext.ball_query_group(new_xyz.contiguous(), xyz.contiguous(), list(scales), group=True)


# UPK_LRF_SVD=torch keeps torch.svd inside the groupers (bit-for-bit the reference's GPU behaviour where the sign vote
# of a frame ties and the raw SVD sign decides); default: the fused kernel with its deterministic tie rule
_LRF_KERNEL = os.environ.get("UPK_LRF_SVD", "kernel") != "torch"


def lrf_group(xyz, new_xyz, grouped_xyz, radius, eps=1e-10, use_xyz=True, normalize_xyz=False):
    """LRF_batch + the feature assembly of QueryAndLRFGroup.forward (pointnet2_utils.py:429-481, :556-571) in one kernel.
    xyz, new_xyz (B,n,3); grouped_xyz (B,3,n,ns) ABSOLUTE neighbour coordinates -> (B,6 or 3,n,ns)."""
    from .. import _lib as L

    L.check_cuda(xyz, "xyz")
    c = xyz.float().contiguous()
    q = new_xyz.float().contiguous()
    g = grouped_xyz.float().contiguous()
    B, _, n, ns = g.shape
    out = torch.empty((B, 6 if use_xyz else 3, n, ns), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device):
        L.check(L.load().upk_lrf_group(L.ptr(c), L.ptr(q), L.ptr(g), B, n, ns, float(radius), float(eps),
                                       1 if use_xyz else 0, 1 if normalize_xyz else 0, L.ptr(out), L.stream_ptr(g)),
                "lrf_group")
    return out


def _uniform_resample_(idx, nsample):
    """sample_uniformly branch of the reference groupers (pointnet2_utils.py:342-351,
    :542-551): per ball, unique indices padded with random re-draws.  Host loop,
    as in the reference; not on the UNOPose path (sample_uniformly=False)."""
    unique_cnt = torch.zeros((idx.shape[0], idx.shape[1]))
    for ib in range(idx.shape[0]):
        for ir in range(idx.shape[1]):
            uniq = torch.unique(idx[ib, ir, :])
            k = uniq.shape[0]
            unique_cnt[ib, ir] = k
            pick = torch.randint(0, k, (nsample - k,), dtype=torch.long)
            idx[ib, ir, :] = torch.cat((uniq, uniq[pick]))
    return unique_cnt


class QueryAndGroup(nn.Module):
    """Ball query + grouping.  Ref: pointnet2_utils.py:292-378.
    forward(xyz (B,N,3), new_xyz (B,npoint,3), features (B,C,N)|None)
      -> (B, 3+C, npoint, nsample) [, grouped_xyz] [, unique_cnt]"""

    def __init__(self, radius, nsample, use_xyz=True, ret_grouped_xyz=False, normalize_xyz=False,
                 sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        if self.ret_unique_cnt:
            assert self.sample_uniformly

    def forward(self, xyz, new_xyz, features=None, pre=None):
        """`pre` = (idx, grouped_xyz) from ball_query_and_group (a caller that shares one scan between scales)."""
        unique_cnt = None
        if pre is None and _fusable(xyz, self.sample_uniformly):
            pre = ball_query_and_group(xyz, new_xyz, [(self.radius, self.nsample)])[0]
        if pre is not None:
            idx, grouped_xyz = pre
        else:
            idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
            unique_cnt = _uniform_resample_(idx, self.nsample) if self.sample_uniformly else None
            grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            new_features = torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = grouped_xyz
        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)


class GroupAll(nn.Module):
    """Group everything into one ball.  Ref: pointnet2_utils.py:381-426."""

    def __init__(self, use_xyz=True, ret_grouped_xyz=False):
        super().__init__()
        self.use_xyz = use_xyz
        self.ret_grouped_xyz = ret_grouped_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            gf = features.unsqueeze(2)
            new_features = torch.cat([grouped_xyz, gf], dim=1) if self.use_xyz else gf
        else:
            new_features = grouped_xyz
        return (new_features, grouped_xyz) if self.ret_grouped_xyz else new_features


class QueryAndLRFGroup(nn.Module):
    """Ball query + grouping + per-centre local reference frame.  Ref: pointnet2_utils.py:484-584.

    The reference additionally runs ``grouping_operation(features, idx)`` even
    when ``use_feature=False`` and throws the result away (:565 vs :566-571,
    SURVEY.md A.12); that dead launch is skipped here — outputs are identical.
    """

    def __init__(self, radius, nsample, use_xyz=False, use_feature=False, ret_grouped_xyz=False,
                 normalize_xyz=False, sample_uniformly=False, ret_unique_cnt=False):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz
        self.use_feature = use_feature
        self.ret_grouped_xyz = ret_grouped_xyz
        self.normalize_xyz = normalize_xyz
        self.sample_uniformly = sample_uniformly
        self.ret_unique_cnt = ret_unique_cnt
        self.lrf = LRF_batch(r_lrf=self.radius)
        if self.ret_unique_cnt:
            assert self.sample_uniformly

    def forward(self, xyz, new_xyz, features=None, pre=None):
        """`pre` = (idx, grouped_xyz) from ball_query_and_group (a caller that shares one scan between scales)."""
        unique_cnt = None
        if pre is None and _fusable(xyz, self.sample_uniformly):
            pre = ball_query_and_group(xyz, new_xyz, [(self.radius, self.nsample)])[0]
        if pre is not None:
            idx, grouped_xyz = pre
        else:
            idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
            unique_cnt = _uniform_resample_(idx, self.nsample) if self.sample_uniformly else None
            grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,nsample)
        if (pre is not None and _LRF_KERNEL and not torch.is_grad_enabled() and xyz.shape[1] == new_xyz.shape[1]
                and not (features is not None and self.use_feature) and not self.ret_grouped_xyz
                and not self.ret_unique_cnt and (features is not None or self.use_xyz)):
            # frames + feature assembly in ONE kernel (upk_lrf_group) instead of a cuSOLVER SVD and ~15 passes
            return lrf_group(xyz, new_xyz, grouped_xyz, self.radius, self.lrf.eps,
                             use_xyz=(features is not None and self.use_xyz), normalize_xyz=self.normalize_xyz)
        lrf_features = self.lrf(xyz, grouped_xyz.transpose(1, 2))  # (B,npoint,3,nsample)
        lrf_features = lrf_features.transpose(1, 2).contiguous()
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if self.normalize_xyz:
            grouped_xyz = grouped_xyz / self.radius
        if features is not None:
            new_features = torch.cat([grouped_xyz, lrf_features], dim=1) if self.use_xyz else lrf_features
            if self.use_feature:
                grouped_features = grouping_operation(features, idx)
                new_features = torch.cat([grouped_features, new_features], dim=1)
        else:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            new_features = lrf_features
        ret = [new_features]
        if self.ret_grouped_xyz:
            ret.append(grouped_xyz)
        if self.ret_unique_cnt:
            ret.append(unique_cnt)
        return ret[0] if len(ret) == 1 else tuple(ret)
