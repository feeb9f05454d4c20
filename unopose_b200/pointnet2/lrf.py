"""Per-centre local reference frames: the torch implementation (CPU tensors, autograd, `UPK_LRF_SVD=torch`); CUDA
inference goes through `pointnet2_utils.lrf_group` (one kernel).

``LRF_batch`` follows the maths of the reference class of the same name
(core/unopose/model/pointnet2/pointnet2_utils.py:429-481): z axis = covariance
eigenvector of the smallest eigenvalue with majority-vote sign, x axis =
distance/height-weighted projection, y = x × z; neighbours are expressed in
that frame, scaled by 1/r.
"""
import torch
import torch.nn as nn


class LRF_batch(nn.Module):
    def __init__(self, eps=1e-10, r_lrf=0.1):
        super().__init__()
        self.eps = eps
        self.r_lrf = r_lrf

    def forward(self, xyz, xyz_group):
        """xyz (B,N,3) centres, xyz_group (B,N,3,M) neighbours -> (B,N,3,M)."""
        B, N, _, M = xyz_group.shape
        centre = xyz.unsqueeze(3)
        to_centre = centre - xyz_group  # p - p_i, (B,N,3,M)
        cov = torch.einsum("bnim,bnjm->bnij", to_centre, to_centre) / M
        _, _, v = torch.svd(cov)
        z_raw = v[..., -1]  # (B,N,3) direction of least variance
        with torch.no_grad():
            h = torch.einsum("bni,bnim->bnm", z_raw, to_centre)
            vote = (h > 1e-3).sum(-1) - (h < -1e-3).sum(-1)
            sign = 1.0 - 2.0 * (vote < 0).to(xyz_group.dtype)
        zp = sign.unsqueeze(-1) * z_raw  # (B,N,3)

        rel = -to_centre  # p_i - p
        height = torch.einsum("bni,bnim->bnm", zp, rel)  # (B,N,M)
        in_plane = rel - height.unsqueeze(2) * zp.unsqueeze(3)
        dist = torch.sqrt((rel ** 2).sum(dim=2))  # (B,N,M)
        alpha = (self.r_lrf - dist) ** 2
        beta = height * height
        x_dir = ((alpha * beta).unsqueeze(2) * in_plane).sum(3)  # (B,N,3)
        xp = x_dir / (torch.sqrt((x_dir ** 2).sum(2, keepdim=True)) + self.eps)
        yp = torch.cross(xp, zp, dim=2)
        frame = torch.stack((xp, yp, zp), dim=3)  # columns x,y,z  (B,N,3,3)
        local = (xyz_group - centre) / self.r_lrf
        return torch.einsum("bnij,bnim->bnjm", frame, local)


class LRF(nn.Module):
    """Global (per-cloud) reference frame, anchored at a given centre — the reference class of the same
    name (core/unopose/utils/model_utils.py:766-823).  Host-side torch code.

    forward(xyz (B,3,1) centre, xyz_group (B,3,N) cloud) -> (B,3,N) coordinates in the frame, scaled by 1/r_lrf
    with r_lrf a (B,) tensor."""

    def __init__(self, r_lrf, eps=1e-10):
        super().__init__()
        self.eps = eps
        self.r_lrf = r_lrf

    def forward(self, xyz, xyz_group):
        B, _, N = xyz_group.shape
        r = self.r_lrf[:, None, None]
        to_centre = xyz - xyz_group                                   # p - p_i  (B,3,N)
        cov = torch.bmm(to_centre, to_centre.transpose(1, 2)) / N
        _, _, v = torch.svd(cov)
        z_raw = v[..., -1]                                            # (B,3)
        with torch.no_grad():
            h = torch.einsum("bi,bin->bn", z_raw, to_centre)
            vote = (h > 1e-3).sum(-1) - (h < -1e-3).sum(-1)
            sign = 1.0 - 2.0 * (vote < 0).to(xyz_group.dtype)
        zp = sign.unsqueeze(-1) * z_raw                               # (B,3)
        rel = -to_centre
        height = torch.einsum("bi,bin->bn", zp, rel)                  # (B,N)
        in_plane = rel - zp.unsqueeze(2) * height.unsqueeze(1)
        dist = torch.sqrt((rel ** 2).sum(dim=1))                      # (B,N)
        alpha = (self.r_lrf[:, None] - dist) ** 2
        x_dir = ((alpha * height * height).unsqueeze(1) * in_plane).sum(2)
        xp = x_dir / (torch.sqrt((x_dir ** 2).sum(1, keepdim=True)) + self.eps)
        yp = torch.cross(xp, zp, dim=1)
        frame = torch.stack((xp, yp, zp), dim=2)                      # columns x,y,z
        return torch.einsum("bij,bin->bjn", frame, (xyz_group - xyz) / r)


def get_batch_lrf(pts, use_ref_rad=False, radius=None, eps=1e-10, return_frame=False):
    """pts (B,N,3) -> (B,N,3): coordinates in the cloud's global reference frame, the reference models'
    ``get_batch_lrf`` (oneref_grf_predator_pose_estimation_model.py:78-93: centroid, radius, ``LRF``).
    CUDA tensors run ONE kernel (upk_global_lrf); CPU tensors / tensors that require grad take the torch module
    above.  `radius` (B,) overrides the radius choice."""
    if radius is None and use_ref_rad:
        radius = torch.ones(pts.shape[0], device=pts.device)
    if not pts.is_cuda or (torch.is_grad_enabled() and pts.requires_grad):
        centroids = torch.mean(pts, 1, True)
        r = radius if radius is not None else torch.norm(pts - centroids, dim=2).max(1)[0]
        out = LRF(r, eps)(centroids.transpose(1, 2), pts.transpose(1, 2)).transpose(1, 2).contiguous()
        return (out, None) if return_frame else out
    from .. import _lib as L

    p = pts.float().contiguous()
    B, N = p.shape[:2]
    out = torch.empty_like(p)
    frame = torch.empty((B, 13), dtype=torch.float32, device=p.device) if return_frame else None
    rad = radius.float().contiguous() if radius is not None else None
    with torch.cuda.device(p.device):
        L.check(L.load().upk_global_lrf(L.ptr(p), L.ptr(rad), B, N, float(eps), L.ptr(out), L.ptr(frame),
                                        L.stream_ptr(p)), "global_lrf")
    return (out, frame) if return_frame else out
