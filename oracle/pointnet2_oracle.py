"""ORACLE — test infrastructure only.  numpy/ctypes front-end of pointnet2_oracle.c.

Each function mirrors one reference extension function
(core/unopose/model/pointnet2/_ext_src/src/bindings.cpp:11-24) on CPU numpy arrays.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpointnet2_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def build():
    src = os.path.join(_HERE, "pointnet2_oracle.c")
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def furthest_point_sampling(xyz, m):
    xyz = _f32(xyz)
    b, n, _ = xyz.shape
    out = np.zeros((b, m), np.int32)
    lib().oracle_furthest_point_sampling(_fp(xyz), b, n, m, _ip(out))
    return out


def gather_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, c, n = points.shape
    m = idx.shape[1]
    out = np.zeros((b, c, m), np.float32)
    lib().oracle_gather_points(_fp(points), _ip(idx), b, c, n, m, _fp(out))
    return out


def gather_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    b, c, m = grad_out.shape
    out = np.zeros((b, c, n), np.float32)
    lib().oracle_gather_points_grad(_fp(grad_out), _ip(idx), b, c, n, m, _fp(out))
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    new_xyz, xyz = _f32(new_xyz), _f32(xyz)
    b, m, _ = new_xyz.shape
    n = xyz.shape[1]
    out = np.zeros((b, m, nsample), np.int32)
    lib().oracle_ball_query(_fp(new_xyz), _fp(xyz), b, n, m, ctypes.c_float(radius), nsample, _ip(out))
    return out


def group_points(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, c, n = points.shape
    _, npoints, nsample = idx.shape
    out = np.zeros((b, c, npoints, nsample), np.float32)
    lib().oracle_group_points(_fp(points), _ip(idx), b, c, n, npoints, nsample, _fp(out))
    return out


def group_points_grad(grad_out, idx, n):
    grad_out, idx = _f32(grad_out), _i32(idx)
    b, c, npoints, nsample = grad_out.shape
    out = np.zeros((b, c, n), np.float32)
    lib().oracle_group_points_grad(_fp(grad_out), _ip(idx), b, c, n, npoints, nsample, _fp(out))
    return out


def three_nn(unknown, known):
    unknown, known = _f32(unknown), _f32(known)
    b, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.zeros((b, n, 3), np.float32)
    idx = np.zeros((b, n, 3), np.int32)
    lib().oracle_three_nn(_fp(unknown), _fp(known), b, n, m, _fp(d2), _ip(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, c, m = points.shape
    n = idx.shape[1]
    out = np.zeros((b, c, n), np.float32)
    lib().oracle_three_interpolate(_fp(points), _ip(idx), _fp(weight), b, c, m, n, _fp(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, idx, weight = _f32(grad_out), _i32(idx), _f32(weight)
    b, c, n = grad_out.shape
    out = np.zeros((b, c, m), np.float32)
    lib().oracle_three_interpolate_grad(_fp(grad_out), _ip(idx), _fp(weight), b, c, n, m, _fp(out))
    return out
