"""ctypes binding of libunopose_b200.so (the C ABI declared in include/unopose_b200.h).

There is NO fallback: if the shared library is missing or a symbol cannot be
resolved, importing/calling raises.  The product path never routes through the
CPU oracle or through plain torch ops.
"""
import ctypes
import os

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libunopose_b200.so")

_lib = None

c_f = ctypes.c_void_p  # device pointers are passed as raw addresses
c_i = ctypes.c_int
c_fl = ctypes.c_float
c_st = ctypes.c_void_p
c_sz = ctypes.c_size_t

# name -> argtypes (restype is always int unless noted)
_SIGNATURES = {
    "upk_abi_version": [],
    "upk_built_sm": [],
    "upk_furthest_point_sampling": [c_f, c_i, c_i, c_i, c_f, c_st],
    "upk_gather_points": [c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_gather_points_grad": [c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_ball_query": [c_f, c_f, c_i, c_i, c_i, c_fl, c_i, c_f, c_st],
    "upk_ball_query_group": [c_f, c_f, c_i, c_i, c_i, c_fl, c_i, c_f, c_f, c_fl, c_i, c_f, c_f, c_st],
    "upk_group_points": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_group_points_grad": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_gather_rows": [c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_gather_rows_grad": [c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_three_nn": [c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_st],
    "upk_three_interpolate": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_three_interpolate_grad": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_set_similarity_mode": [c_i],
    "upk_feature_similarity": [c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_i, c_i, c_f, c_sz, c_f, c_st],
    "upk_coarse_pose": [c_f, c_f, c_i, c_f, c_i, c_f, c_f, c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_i,
                        c_f, c_sz, c_f, c_f, c_f, c_f, c_f, c_st],
    "upk_coarse_assignment": [c_f, c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_f, c_sz, c_f, c_f, c_f, c_st],
    "upk_coarse_assignment_profile": [c_f, c_f, c_i, c_f, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_st],
    "upk_sample_hypotheses": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_st],
    "upk_kabsch_triplets": [c_f, c_f, c_i, c_f, c_f, c_f, c_st],
    "upk_topk_smallest": [c_f, c_i, c_i, c_i, c_f, c_st],
    "upk_score_hypotheses": [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_select_best": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_st],
    "upk_fill_f32": [c_f, c_sz, c_fl, c_st],
    "upk_pack_candidates": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_unpack_candidates": [c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_st],
    # peer exchange (the upk_peer_t* is passed with ctypes.byref(peer.PeerStruct))
    "upk_pack_candidates_peer": [c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, ctypes.c_void_p, ctypes.c_size_t,
                                 ctypes.c_size_t, c_i, c_st],
    "upk_unpack_candidates_peer": [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_st],
    "upk_score_hypotheses_peer": [c_f, c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, ctypes.c_void_p,
                                  ctypes.c_size_t, ctypes.c_size_t, c_i, c_st],
    "upk_select_best_peer": [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, c_i, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_f,
                             c_f, c_f, c_f, c_st],
    "upk_unpack_candidates_compact_peer": [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, c_i, c_i, c_i, c_f, c_f, c_f, c_f,
                                           c_st],
    "upk_unpack_candidates_compact": [c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_st],
    "upk_topk_smallest_ld": [c_f, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_select_best_map": [c_f, c_f, c_f, c_f, c_f, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_st],
    "upk_rpe_scores": [c_f, c_f, c_i, c_i, c_i, c_i, c_i, c_f, c_st],
    "upk_linear": [c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_f, c_sz, c_f, c_st],
    "upk_peer_all_gather": [c_f, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, c_i, c_st],
    "upk_peer_wait": [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t, c_i, c_f, ctypes.c_size_t, c_st],
    "upk_fine_pose": [c_f, c_f, c_i, c_f, c_i, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_fl,
                      c_f, c_sz, c_f, c_f, c_f, c_f, c_st],
    "upk_feature_similarity_stats": [c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_f, c_sz, c_f, c_f, c_sz, c_st],
    "upk_fine_pose_stats": [c_f, c_f, c_sz, c_fl, c_f, c_i, c_f, c_i, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_fl,
                            c_f, c_sz, c_f, c_f, c_f, c_f, c_st],
    "upk_feature_similarity_stats_ld": [c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_f, c_sz, c_f, c_i, c_f, c_sz, c_st],
    "upk_fine_pose_ld": [c_f, c_i, c_f, c_sz, c_fl, c_f, c_i, c_f, c_i, c_f, c_f, c_f, c_i, c_i, c_i, c_i, c_fl, c_fl,
                         c_f, c_sz, c_f, c_f, c_f, c_f, c_st],
    "upk_weighted_procrustes": [c_f, c_f, c_f, c_i, c_i, c_fl, c_fl, c_f, c_f, c_st],
    "upk_global_lrf": [c_f, c_f, c_i, c_i, c_fl, c_f, c_f, c_st],
    "upk_geometric_embedding_supported": [c_i, c_i],
    "upk_geometric_embedding_indices": [c_f, c_i, c_i, c_i, c_fl, c_fl, c_f, c_f, c_st],
    "upk_geometric_embedding": [c_f, c_i, c_i, c_i, c_i, c_fl, c_fl, c_f, c_f, c_f, c_f, c_f, c_i, c_f, c_sz, c_f, c_st],
    "upk_shared_mlp_max_supported": [c_i, c_i, c_i, c_i, c_i, c_i],
    "upk_shared_mlp_max": [c_f, c_i, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_f, c_f, c_f, c_f, c_f, c_f, c_st],
    "upk_lrf_group": [c_f, c_f, c_f, c_i, c_i, c_i, c_fl, c_fl, c_i, c_i, c_f, c_st],
    "upk_transform_points": [c_f, c_f, c_f, c_i, c_i, c_f, c_st],
    "upk_host_procrustes_rotation": [c_f, c_i, c_f],
    "upk_host_lrf_z_axis": [c_f, c_i, c_f],
}

# size_t-returning workspace queries
_SIZE_FUNCS = {
    "upk_feature_similarity_workspace_bytes": [c_i, c_i, c_i, c_i, c_i],
    "upk_coarse_pose_workspace_bytes": [c_i, c_i, c_i, c_i, c_i],
    "upk_coarse_assignment_workspace_bytes": [c_i, c_i, c_i],
    "upk_fine_pose_workspace_bytes": [c_i, c_i, c_i],
    "upk_similarity_stats_bytes": [c_i, c_i, c_i],
    "upk_geometric_embedding_workspace_bytes": [c_i, c_i, c_i, c_i],
    "upk_linear_workspace_bytes": [c_i, c_i, c_i],
}


class CoarseDebug(ctypes.Structure):
    """struct upk_coarse_debug (include/unopose_b200.h)"""
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("w1", "w2", "cdf", "idx1", "idx2", "Rs", "ts", "resid", "top", "scores")]


class FineDebug(ctypes.Structure):
    """struct upk_fine_debug (include/unopose_b200.h)"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("w1", "w2", "soft", "asum", "nn")]


class UnoposeNativeError(RuntimeError):
    pass


def load():
    """Load the library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise UnoposeNativeError(
            "libunopose_b200.so not found at %s — build it with `python -m unopose_b200.build` "
            "(there is no CPU/torch fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = ctypes.c_int
    for name, argtypes in _SIZE_FUNCS.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = ctypes.c_size_t
    lib.upk_launch_count.argtypes = []
    lib.upk_launch_count.restype = ctypes.c_ulonglong
    _lib = lib
    return lib


def exported_symbols():
    return sorted(list(_SIGNATURES) + list(_SIZE_FUNCS) + ["upk_launch_count"])


def launch_count():
    return int(load().upk_launch_count())


def check(rc, what):
    if rc == 0:
        return
    if rc < 0:
        raise UnoposeNativeError("%s: invalid/unsupported arguments (code %d)" % (what, rc))
    raise UnoposeNativeError("%s: CUDA error %d" % (what, rc))


def stream_ptr(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def ptr(t):
    return t.data_ptr() if t is not None else None


# --- argument checks with the reference's error text (_ext_src/include/utils.h:10-30) ---
def check_cuda(x, name):
    if not x.is_cuda:
        raise RuntimeError("CPU not supported")  # sampling.cpp:39,65,87


def check_contiguous(x, name):
    if not x.is_contiguous():
        raise RuntimeError("%s must be a contiguous tensor" % name)


def check_float(x, name):
    if x.dtype != torch.float32:
        raise RuntimeError("%s must be a float tensor" % name)


def check_int(x, name):
    if x.dtype != torch.int32:
        raise RuntimeError("%s must be an int tensor" % name)
