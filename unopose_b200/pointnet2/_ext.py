"""Drop-in for the reference's compiled module ``core.unopose.model.pointnet2._ext``.

Same 9 functions, same positional signatures and the same error behaviour as
``_ext_src/src/bindings.cpp:11-24`` — but each call goes through the C ABI of
libunopose_b200.so (include/unopose_b200.h) into sm_100a kernels.  Outputs are
allocated here (callee-allocates, like the reference's torch::zeros) with
torch.empty: the kernels overwrite every element.
"""
import torch

from .. import _lib as L


def _dev(x):
    return torch.cuda.device(x.device)


def furthest_point_sampling(points, nsamples):
    """points (B,N,3) f32 -> (B,nsamples) int32.  Ref: sampling.cpp:70-91."""
    L.check_contiguous(points, "points")
    L.check_float(points, "points")
    L.check_cuda(points, "points")
    b, n, _ = points.shape
    out = torch.empty((b, int(nsamples)), dtype=torch.int32, device=points.device)
    with _dev(points):
        L.check(L.load().upk_furthest_point_sampling(L.ptr(points), b, n, int(nsamples), L.ptr(out),
                                                     L.stream_ptr(points)), "furthest_point_sampling")
    return out


def gather_points(points, idx):
    """points (B,C,N) f32, idx (B,M) int32 -> (B,C,M).  Ref: sampling.cpp:20-43."""
    L.check_contiguous(points, "points")
    L.check_contiguous(idx, "idx")
    L.check_float(points, "points")
    L.check_int(idx, "idx")
    L.check_cuda(points, "points")
    if not idx.is_cuda:
        raise RuntimeError("idx must be a CUDA tensor")
    b, c, n = points.shape
    m = idx.shape[1]
    out = torch.empty((b, c, m), dtype=torch.float32, device=points.device)
    with _dev(points):
        L.check(L.load().upk_gather_points(L.ptr(points), L.ptr(idx), b, c, n, m, L.ptr(out),
                                           L.stream_ptr(points)), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """grad_out (B,C,M), idx (B,M) -> (B,C,n).  Ref: sampling.cpp:45-68."""
    L.check_contiguous(grad_out, "grad_out")
    L.check_contiguous(idx, "idx")
    L.check_float(grad_out, "grad_out")
    L.check_int(idx, "idx")
    L.check_cuda(grad_out, "grad_out")
    if not idx.is_cuda:
        raise RuntimeError("idx must be a CUDA tensor")
    b, c, m = grad_out.shape
    out = torch.empty((b, c, int(n)), dtype=torch.float32, device=grad_out.device)
    with _dev(grad_out):
        L.check(L.load().upk_gather_points_grad(L.ptr(grad_out), L.ptr(idx), b, c, int(n), m,
                                                L.ptr(out), L.stream_ptr(grad_out)), "gather_points_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """new_xyz (B,M,3), xyz (B,N,3) -> (B,M,nsample) int32.  Ref: ball_query.cpp:13-37.
    NOTE the argument order (new_xyz first) — it differs from the Python-level
    BallQuery.forward(radius, nsample, xyz, new_xyz)."""
    L.check_contiguous(new_xyz, "new_xyz")
    L.check_contiguous(xyz, "xyz")
    L.check_float(new_xyz, "new_xyz")
    L.check_float(xyz, "xyz")
    L.check_cuda(new_xyz, "new_xyz")
    if not xyz.is_cuda:
        raise RuntimeError("xyz must be a CUDA tensor")
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = torch.empty((b, m, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    with _dev(new_xyz):
        L.check(L.load().upk_ball_query(L.ptr(new_xyz), L.ptr(xyz), b, n, m, float(radius), int(nsample),
                                        L.ptr(out), L.stream_ptr(new_xyz)), "ball_query")
    return out


def ball_query_group(new_xyz, xyz, scales, group=True):
    """Fused ball_query (+ grouping of the xyz channels) for one or two (radius, nsample) scales in one
    scan of the cloud (upk_ball_query_group).  new_xyz (B,M,3), xyz (B,N,3), scales = [(r, ns)] or
    [(r0, ns0), (r1, ns1)] -> [(idx (B,M,ns) int32, grouped_xyz (B,3,M,ns) | None), ...] with
    idx == ball_query(new_xyz, xyz, r, ns) and grouped_xyz == group_points(xyz^T, idx), bit-exact.
    Not part of the reference's _ext; it replaces the ball_query -> transpose -> grouping_operation
    sequences of QueryAndGroup / QueryAndLRFGroup (pointnet2_utils.py:292-378, :484-584)."""
    L.check_contiguous(new_xyz, "new_xyz")
    L.check_contiguous(xyz, "xyz")
    L.check_float(new_xyz, "new_xyz")
    L.check_float(xyz, "xyz")
    L.check_cuda(new_xyz, "new_xyz")
    if not xyz.is_cuda:
        raise RuntimeError("xyz must be a CUDA tensor")
    if not 1 <= len(scales) <= 2:
        raise ValueError("ball_query_group takes one or two (radius, nsample) scales")
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    outs, args = [], []
    for r, ns in scales:
        idx = torch.empty((b, m, int(ns)), dtype=torch.int32, device=new_xyz.device)
        g = torch.empty((b, 3, m, int(ns)), dtype=torch.float32, device=new_xyz.device) if group else None
        outs.append((idx, g))
        args += [float(r), int(ns), L.ptr(idx), L.ptr(g) if g is not None else None]
    if len(scales) == 1:
        args += [0.0, 0, None, None]
    with _dev(new_xyz):
        L.check(L.load().upk_ball_query_group(L.ptr(new_xyz), L.ptr(xyz), b, n, m, *args,
                                              L.stream_ptr(new_xyz)), "ball_query_group")
    return outs


def group_points(points, idx):
    """points (B,C,N), idx (B,npoints,nsample) -> (B,C,npoints,nsample).  Ref: group_points.cpp:17-40."""
    L.check_contiguous(points, "points")
    L.check_contiguous(idx, "idx")
    L.check_float(points, "points")
    L.check_int(idx, "idx")
    L.check_cuda(points, "points")
    if not idx.is_cuda:
        raise RuntimeError("idx must be a CUDA tensor")
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = torch.empty((b, c, npoints, nsample), dtype=torch.float32, device=points.device)
    with _dev(points):
        L.check(L.load().upk_group_points(L.ptr(points), L.ptr(idx), b, c, n, npoints, nsample, L.ptr(out),
                                          L.stream_ptr(points)), "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """grad_out (B,C,npoints,nsample), idx -> (B,C,n).  Ref: group_points.cpp:42-65."""
    L.check_contiguous(grad_out, "grad_out")
    L.check_contiguous(idx, "idx")
    L.check_float(grad_out, "grad_out")
    L.check_int(idx, "idx")
    L.check_cuda(grad_out, "grad_out")
    if not idx.is_cuda:
        raise RuntimeError("idx must be a CUDA tensor")
    b, c, npoints, nsample = grad_out.shape
    out = torch.empty((b, c, int(n)), dtype=torch.float32, device=grad_out.device)
    with _dev(grad_out):
        L.check(L.load().upk_group_points_grad(L.ptr(grad_out), L.ptr(idx), b, c, int(n), npoints, nsample,
                                               L.ptr(out), L.stream_ptr(grad_out)), "group_points_grad")
    return out


def three_nn(unknowns, knows):
    """unknowns (B,n,3), knows (B,m,3) -> [dist2 (B,n,3) f32, idx (B,n,3) int32].  Ref: interpolate.cpp:19-45."""
    L.check_contiguous(unknowns, "unknowns")
    L.check_contiguous(knows, "knows")
    L.check_float(unknowns, "unknowns")
    L.check_float(knows, "knows")
    L.check_cuda(unknowns, "unknowns")
    if not knows.is_cuda:
        raise RuntimeError("knows must be a CUDA tensor")
    b, n, _ = unknowns.shape
    m = knows.shape[1]
    dist2 = torch.empty((b, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((b, n, 3), dtype=torch.int32, device=unknowns.device)
    with _dev(unknowns):
        L.check(L.load().upk_three_nn(L.ptr(unknowns), L.ptr(knows), b, n, m, L.ptr(dist2), L.ptr(idx),
                                      L.stream_ptr(unknowns)), "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """points (B,c,m), idx (B,n,3), weight (B,n,3) -> (B,c,n).  Ref: interpolate.cpp:47-74."""
    L.check_contiguous(points, "points")
    L.check_contiguous(idx, "idx")
    L.check_contiguous(weight, "weight")
    L.check_float(points, "points")
    L.check_int(idx, "idx")
    L.check_float(weight, "weight")
    L.check_cuda(points, "points")
    if not idx.is_cuda:
        raise RuntimeError("idx must be a CUDA tensor")
    if not weight.is_cuda:
        raise RuntimeError("weight must be a CUDA tensor")
    b, c, m = points.shape
    n = idx.shape[1]
    out = torch.empty((b, c, n), dtype=torch.float32, device=points.device)
    with _dev(points):
        L.check(L.load().upk_three_interpolate(L.ptr(points), L.ptr(idx), L.ptr(weight), b, c, m, n, L.ptr(out),
                                               L.stream_ptr(points)), "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """grad_out (B,c,n), idx, weight -> (B,c,m).  Ref: interpolate.cpp:76-104."""
    L.check_contiguous(grad_out, "grad_out")
    L.check_contiguous(idx, "idx")
    L.check_contiguous(weight, "weight")
    L.check_float(grad_out, "grad_out")
    L.check_int(idx, "idx")
    L.check_float(weight, "weight")
    L.check_cuda(grad_out, "grad_out")
    if not idx.is_cuda:
        raise RuntimeError("idx must be a CUDA tensor")
    if not weight.is_cuda:
        raise RuntimeError("weight must be a CUDA tensor")
    b, c, n = grad_out.shape
    out = torch.empty((b, c, int(m)), dtype=torch.float32, device=grad_out.device)
    with _dev(grad_out):
        L.check(L.load().upk_three_interpolate_grad(L.ptr(grad_out), L.ptr(idx), L.ptr(weight), b, c, n, int(m),
                                                    L.ptr(out), L.stream_ptr(grad_out)), "three_interpolate_grad")
    return out
