"""Coarse / fine point-matching modules: same constructor config keys, forward
signature, `end_points` keys and parameter names as the reference
(core/unopose/model/oneref_predator_coarse_point_matching.py, oneref_predator_fine_point_matching.py),
so released checkpoints load with `load_state_dict` unchanged.  Evaluation path only: the training
branches need the reference's loss functions (loss_utils.py, out of scope) and raise.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..model_utils import (compute_coarse_Rt_overlap, compute_feature_similarity, compute_fine_Rt_overlap,
                           transform_points)
from ..pointnet2.pointnet2_utils import QueryAndGroup, QueryAndLRFGroup, ball_query_and_group
from .layers import Conv1d, SharedMLP
from .linear import linear
from .transformer import GeometricTransformer, SparseToDenseTransformer


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    if hasattr(cfg, "get"):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def _overlap_scores(head_out, n1):
    """scores (B, n1+1+n2+1, 1) from the score head over cat(f1, f2) -> sigmoid'ed (B, n1+n2) without the
    two background tokens (coarse module :68-72, fine module :91-95)."""
    s1, s2 = head_out[:, 1:(n1 + 1)], head_out[:, (n1 + 2):]
    return torch.clamp(torch.sigmoid(torch.cat((s1, s2), dim=1).squeeze(-1)), min=0, max=1)


class CoarsePointMatchingOneRef(nn.Module):
    """forward(p1 (B,n1,3), f1 (B,n1,C), geo1 (B,n1+1,n1+1,C), p2, f2, geo2, radius (B,), end_points) ->
    end_points with init_R (B,3,3), init_t (B,3), init_pose_score (B,)."""

    def __init__(self, cfg, return_feat=False):
        super().__init__()
        self.cfg = cfg
        self.return_feat = return_feat
        self.nblock = _get(cfg, "nblock")
        hidden = _get(cfg, "hidden_dim")
        self.in_proj = nn.Linear(_get(cfg, "input_dim"), hidden)
        self.out_proj = nn.Linear(hidden, _get(cfg, "out_dim"))
        self.bg_token = nn.Parameter(torch.randn(1, 1, hidden) * 0.02)
        self.score_heads = nn.ModuleList([nn.Linear(hidden, 1) for _ in range(self.nblock)])
        self.transformers = nn.ModuleList(
            [GeometricTransformer(["self", "cross"], hidden, num_heads=4) for _ in range(self.nblock)])

    def matching_features(self, f1, geo1, f2, geo2):
        """The dense (cuBLAS) part: returns out_proj(f1), out_proj(f2) (B,n+1,C) and the overlap scores."""
        B, n1 = f1.size(0), f1.size(1)
        bg = self.bg_token.repeat(B, 1, 1)
        f1 = torch.cat([bg, linear(self.in_proj, f1)], dim=1)
        f2 = torch.cat([bg, linear(self.in_proj, f2)], dim=1)
        for i in range(self.nblock):
            f1, f2 = self.transformers[i](f1, geo1, f2, geo2)
        score = _overlap_scores(self.score_heads[self.nblock - 1](torch.cat((f1, f2), dim=1)), n1)
        return linear(self.out_proj, f1), linear(self.out_proj, f2), score

    def forward(self, p1, f1, geo1, p2, f2, geo2, radius, end_points):
        if self.training:
            raise NotImplementedError("training branch (losses) is out of scope; use the reference module to train")
        g1, g2, score = self.matching_features(f1, geo1, f2, geo2)
        with torch.no_grad():   # the pose solve carries no gradient in the reference either (coarse module :98)
            atten = compute_feature_similarity(g1, g2, _get(self.cfg, "sim_type"), _get(self.cfg, "temp"),
                                               _get(self.cfg, "normalize_feat"))
            init_R, init_t, init_score = compute_coarse_Rt_overlap(atten, score, p1, p2, None, _get(self.cfg, "nproposal1"),
                                                                   _get(self.cfg, "nproposal2"))
        end_points["init_pose_score"] = init_score
        end_points["init_R"] = init_R
        end_points["init_t"] = init_t
        if self.return_feat:
            return end_points, g1, g2
        return end_points


class PositionalEncoding(nn.Module):
    """Two-scale local geometry encoding: ball query -> (LRF) grouping -> SharedMLP -> max over the ball
    (fine module :138-178).  forward(pts (B,N,3)) -> (B,N,out_dim)."""

    def __init__(self, out_dim, r1=0.1, r2=0.2, nsample1=32, nsample2=64, use_lrf=True, use_xyz=False, use_feature=False,
                 bn=True):
        super().__init__()
        if use_lrf:
            self.group1 = QueryAndLRFGroup(r1, nsample1, use_xyz=use_xyz, use_feature=use_feature)
            self.group2 = QueryAndLRFGroup(r2, nsample2, use_xyz=use_xyz, use_feature=use_feature)
        else:
            self.group1 = QueryAndGroup(r1, nsample1, use_xyz=use_xyz)
            self.group2 = QueryAndGroup(r2, nsample2, use_xyz=use_xyz)
        input_dim = 3 + (3 if use_xyz else 0) + (3 if use_feature else 0)
        self.mlp1 = SharedMLP([input_dim, 32, 64, 128], bn=bn)
        self.mlp2 = SharedMLP([input_dim, 32, 64, 128], bn=bn)
        self.mlp3 = Conv1d(256, out_dim, activation=None, bn=None)

    @staticmethod
    def _mlp_max(mlp, x):
        """mlp(x).max over the ball: CUDA inference in eval mode runs the fused tcgen05 kernel (modules/pe.py: the
        three layers of a 128-sample tile stay on chip); otherwise the torch layers (cuDNN/cuBLAS)."""
        if x.is_cuda and not mlp.training and not torch.is_grad_enabled():
            from . import pe
            if pe.supported(mlp, x):
                return pe.shared_mlp_max(x, mlp)
        return mlp(x).max(dim=3)[0]

    def forward(self, pts1, pts2=None):
        if pts2 is None:
            pts2 = pts1
        pts1 = pts1.to(dtype=torch.float32).contiguous()
        pts2 = pts2.to(dtype=torch.float32).contiguous()
        with torch.autocast(device_type="cuda", enabled=False):
            feats = pts1.transpose(1, 2).contiguous()
            pre1 = pre2 = None
            g1, g2 = self.group1, self.group2
            if pts1.is_cuda and not (g1.sample_uniformly or g2.sample_uniformly) and \
                    not (torch.is_grad_enabled() and pts1.requires_grad):
                # both scales from ONE scan of the cloud (fused ball query + grouping kernel)
                pre1, pre2 = ball_query_and_group(pts1, pts2, [(g1.radius, g1.nsample), (g2.radius, g2.nsample)])
            f1 = self._mlp_max(self.mlp1, g1(pts1, pts2, feats, pre=pre1))
            f2 = self._mlp_max(self.mlp2, g2(pts1, pts2, feats, pre=pre2))
            return self.mlp3(torch.cat([f1, f2], dim=1)).transpose(1, 2)


class FinePointMatchingOneRef(nn.Module):
    """forward(p1 (B,n1,3), f1 (B,n1,C), geo1, fps_idx1 (B,m) int32, p2, f2, geo2, fps_idx2, radius, end_points)
    -> end_points with pred_R, pred_t (de-normalised by radius), pred_pose_score."""

    def __init__(self, cfg, return_feat=False):
        super().__init__()
        self.cfg = cfg
        self.return_feat = return_feat
        self.nblock = _get(cfg, "nblock")
        hidden = _get(cfg, "hidden_dim")
        self.in_proj = nn.Linear(_get(cfg, "input_dim"), hidden)
        self.out_proj = nn.Linear(hidden, _get(cfg, "out_dim"))
        self.dis_proj = nn.Linear(2 * hidden, 3)  # present in the reference's state dict, unused in forward (:23)
        self.bg_token = nn.Parameter(torch.randn(1, 1, hidden) * 0.02)
        self.PE = PositionalEncoding(hidden, r1=_get(cfg, "pe_radius1"), r2=_get(cfg, "pe_radius2"),
                                     nsample1=_get(cfg, "nsample1", 32), nsample2=_get(cfg, "nsample2", 64),
                                     use_lrf=_get(cfg, "use_lrf"), use_xyz=_get(cfg, "use_xyz"),
                                     use_feature=_get(cfg, "use_feature", False))
        self.score_heads = nn.ModuleList([nn.Linear(hidden, 1) for _ in range(self.nblock)])
        self.transformers = nn.ModuleList([
            SparseToDenseTransformer(hidden, num_heads=4, sparse_blocks=["self", "cross"],
                                     focusing_factor=_get(cfg, "focusing_factor"), with_bg_token=True,
                                     replace_bg_token=True) for _ in range(self.nblock)])

    def matching_features(self, p1, f1, geo1, fps_idx1, p2, f2, geo2, fps_idx2, end_points):
        B, n1 = p1.size(0), p1.size(1)
        if "init_R" in end_points and "init_t" in end_points:
            if p1.is_cuda and not (torch.is_grad_enabled() and p1.requires_grad):
                p1_ = transform_points(p1, end_points["init_R"], end_points["init_t"])
            else:
                p1_ = (p1 - end_points["init_t"].unsqueeze(1)) @ end_points["init_R"]
        else:
            p1_ = p1
        bg = self.bg_token.repeat(B, 1, 1)
        f1 = torch.cat([bg, linear(self.in_proj, f1) + self.PE(p1_)], dim=1)
        f2 = torch.cat([bg, linear(self.in_proj, f2) + self.PE(p2)], dim=1)
        for i in range(self.nblock):
            f1, f2 = self.transformers[i](f1, geo1, fps_idx1, f2, geo2, fps_idx2)
        score = _overlap_scores(self.score_heads[self.nblock - 1](torch.cat((f1, f2), dim=1)), n1)
        return linear(self.out_proj, f1), linear(self.out_proj, f2), score

    def forward(self, p1, f1, geo1, fps_idx1, p2, f2, geo2, fps_idx2, radius, end_points):
        if self.training:
            raise NotImplementedError("training branch (losses) is out of scope; use the reference module to train")
        g1, g2, score = self.matching_features(p1, f1, geo1, fps_idx1, p2, f2, geo2, fps_idx2, end_points)
        with torch.no_grad():   # evaluation branch: pose only
            atten, stats = compute_feature_similarity(g1, g2, _get(self.cfg, "sim_type"), _get(self.cfg, "temp"),
                                                      _get(self.cfg, "normalize_feat"), return_stats=True)
            pred_R, pred_t, pred_score = compute_fine_Rt_overlap(atten, score, p1, p2, None, stats=stats)
        end_points["pred_R"] = pred_R
        end_points["pred_t"] = pred_t * (radius.reshape(-1, 1) + 1e-6)
        end_points["pred_pose_score"] = pred_score
        if self.return_feat:
            return end_points, g1, g2
        return end_points
