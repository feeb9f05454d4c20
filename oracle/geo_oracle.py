"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path.

Plain-torch restatement of the reference's GeometricStructureEmbedding
(core/unopose/model/transformer.py:287-350) and SinusoidalPositionalEmbedding (:261-284), same operation
order, so on CPU it is pinned against golden vectors produced by importing the reference itself
(tests/golden/make_geo_golden.py -> tests/golden/geo_*.npz, and `geo1`/`geo2` of modules_small.pt), and on
the B200 box (CUDA tensors) it stands in for the reference's GPU torch path (same ATen / cuBLAS kernels).
"""
import math

import torch

from .pose_oracle import pairwise_sqdist


def div_term(d_model):
    """transformer.py:266-268."""
    return torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))


def sinusoid(idx, dterm):
    """transformer.py:271-284: (*,) -> (*, d_model), [sin, cos] interleaved per frequency."""
    shape = idx.shape
    om = idx.reshape(-1, 1, 1) * dterm.view(1, -1, 1)
    emb = torch.cat([torch.sin(om), torch.cos(om)], dim=2)
    return emb.view(*shape, 2 * dterm.numel())


def embedding_indices(points, sigma_d, sigma_a, angle_k):
    """transformer.py:303-336."""
    B, N, _ = points.shape
    dist_map = torch.sqrt(pairwise_sqdist(points, points))
    d_indices = dist_map / sigma_d
    k = angle_k
    knn = dist_map.topk(k=k + 1, dim=2, largest=False)[1][:, :, 1:]
    knn = knn.unsqueeze(3).expand(B, N, k, 3)
    expanded = points.unsqueeze(1).expand(B, N, N, 3)
    knn_points = torch.gather(expanded, dim=2, index=knn)
    ref = knn_points - points.unsqueeze(2)
    anc = points.unsqueeze(1) - points.unsqueeze(2)
    ref = ref.unsqueeze(2).expand(B, N, N, k, 3)
    anc = anc.unsqueeze(3).expand(B, N, N, k, 3)
    sin_values = torch.linalg.norm(torch.cross(ref, anc, dim=-1), dim=-1)
    cos_values = torch.sum(ref * anc, dim=-1)
    angles = torch.atan2(sin_values, cos_values)
    factor_a = 180.0 / (sigma_a * math.pi)
    return d_indices, angles * factor_a


def embed_from_indices(d_indices, a_indices, dterm, w_d, b_d, w_a, b_a, reduction="max"):
    """transformer.py:338-350 given the indices."""
    d = torch.nn.functional.linear(sinusoid(d_indices, dterm), w_d, b_d)
    a = torch.nn.functional.linear(sinusoid(a_indices, dterm), w_a, b_a)
    a = a.max(dim=3)[0] if reduction == "max" else a.mean(dim=3)
    return d + a


def geometric_embedding(points, dterm, w_d, b_d, w_a, b_a, sigma_d, sigma_a, angle_k, reduction="max"):
    d_idx, a_idx = embedding_indices(points, sigma_d, sigma_a, angle_k)
    return embed_from_indices(d_idx, a_idx, dterm, w_d, b_d, w_a, b_a, reduction)


def make_inputs(seed, B, N, C, dev="cpu"):
    """Deterministic test inputs (CPU generator): a radius-normalised cloud behind the background point (1,1,1)
    like the callers build it (oneref_grf_predator_pose_estimation_model.py:33-38), and Linear-initialised weights."""
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(B, N - 1, 3, generator=g)
    p = p / p.norm(dim=2).max(dim=1)[0].view(B, 1, 1)
    pts = torch.cat([torch.ones(B, 1, 3), p], 1)
    bound = 1.0 / math.sqrt(C)
    u = lambda *s: (torch.rand(*s, generator=g) * 2 - 1) * bound  # noqa: E731
    w_d, b_d, w_a, b_a = u(C, C), u(C), u(C, C), u(C)
    return [t.to(dev) for t in (pts, div_term(C), w_d, b_d, w_a, b_a)]
