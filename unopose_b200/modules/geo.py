"""Fused GeometricStructureEmbedding: host wrappers of `upk_geometric_embedding[_indices]`.

Reference: core/unopose/model/transformer.py:287-350 (`GeometricStructureEmbedding`) and :261-284
(`SinusoidalPositionalEmbedding`).  CUDA tensors only — there is no CPU implementation behind these calls.
"""
import torch

from .. import _lib as L


MAX_POINTS = 4096   # the index kernel stages the cloud and one distance row per warp in shared memory (48 B per point)


def supported(hidden_dim, angle_k, n_points=None):
    if n_points is not None and not (angle_k < n_points <= MAX_POINTS):
        return False
    return bool(L.load().upk_geometric_embedding_supported(int(hidden_dim), int(angle_k)))


def _points(points):
    L.check_cuda(points, "points")
    if points.dim() != 3 or points.shape[2] != 3:
        raise RuntimeError("points must be (B, N, 3)")
    return points.float().contiguous()


def geometric_embedding_indices(points, sigma_d, factor_a, angle_k):
    """points (B,N,3) -> d_indices (B,N,N), a_indices (B,N,N,k)  (`get_embedding_indices`, transformer.py:303-336)."""
    p = _points(points)
    B, N = p.shape[:2]
    d = torch.empty((B, N, N), dtype=torch.float32, device=p.device)
    a = torch.empty((B, N, N, angle_k), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        L.check(L.load().upk_geometric_embedding_indices(L.ptr(p), B, N, int(angle_k), float(sigma_d), float(factor_a),
                                                         L.ptr(d), L.ptr(a), L.stream_ptr(p)), "geometric_embedding_indices")
    return d, a


def geometric_embedding(points, div_term, w_d, b_d, w_a, b_a, sigma_d, factor_a, angle_k, reduction="max"):
    """points (B,N,3) -> (B,N,N,C) = proj_d(emb(d)) + red_k proj_a(emb(a_k))  (`forward`, transformer.py:338-350).

    One index kernel + two tcgen05 kernels; the (B,N,N,[k,]C) sinusoid tensors of the reference exist only as
    shared-memory operand tiles."""
    if reduction not in ("max", "mean"):
        raise ValueError(f"Unsupported reduction mode: {reduction}.")
    p = _points(points)
    B, N = p.shape[:2]
    C = w_d.shape[0]
    if not supported(C, angle_k):
        raise L.UnoposeNativeError("geometric_embedding: hidden_dim %d / angle_k %d not supported by the fused kernel"
                                   % (C, angle_k))
    t = [x.detach().float().contiguous() for x in (div_term, w_d, b_d, w_a, b_a)]
    for x in t:
        L.check_cuda(x, "parameter")
    lib = L.load()
    ws_bytes = lib.upk_geometric_embedding_workspace_bytes(B, N, C, int(angle_k))
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=p.device)
    out = torch.empty((B, N, N, C), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        L.check(lib.upk_geometric_embedding(L.ptr(p), B, N, C, int(angle_k), float(sigma_d), float(factor_a),
                                            L.ptr(t[0]), L.ptr(t[1]), L.ptr(t[2]), L.ptr(t[3]), L.ptr(t[4]),
                                            1 if reduction == "mean" else 0, L.ptr(ws), ws_bytes, L.ptr(out),
                                            L.stream_ptr(p)), "geometric_embedding")
    return out
