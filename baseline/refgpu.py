"""Loader for the staged, UNMODIFIED reference (baseline/_ref/, see stage_reference.py) — test / bench infrastructure.

`load()` imports the reference's own `core.unopose.utils.model_utils`, `core.unopose.model.pointnet2.pointnet2_utils`,
`transformer` and the two matching modules from baseline/_ref with

  * `core.unopose.model.pointnet2._ext` = the reference's extension compiled unmodified for sm_100a
    (oracle/_ref/ref_pointnet2_ext.so, oracle/build_ref_ext.py) — its stock GPU code path;
  * `detectron2.utils.logger` stubbed (two no-op log helpers pulled in by loss_utils.py:4; detectron2 is not installed).

`patched(ns)` is a context manager that swaps in unopose_b200 exactly the way INTEGRATION.md §1-2 tells a maintainer
to (the `_ext` module and the pose-function names), leaving every other line of the reference's Python as it is, and
restores the stock functions on exit.  Never imported by the product (unopose_b200/).
"""
import contextlib
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref")
ROOT = os.path.dirname(HERE)

POSE_NAMES = ("compute_feature_similarity", "compute_coarse_Rt", "compute_coarse_Rt_overlap", "compute_fine_Rt",
              "compute_fine_Rt_overlap", "weighted_procrustes", "WeightedProcrustes", "sample_pts_feats",
              "sample_pts_feats_wlrf", "gather_pts_feats", "gather_pts_feats_wlrf")

_ns = None


def available():
    return os.path.exists(os.path.join(REF_ROOT, "core", "unopose", "utils", "model_utils.py"))


def _stub_detectron2():
    if "detectron2.utils.logger" in sys.modules:
        return
    d2l = types.ModuleType("detectron2.utils.logger")
    d2l.log_first_n = d2l.log_every_n = lambda *a, **k: None
    sys.modules.setdefault("detectron2", types.ModuleType("detectron2"))
    sys.modules.setdefault("detectron2.utils", types.ModuleType("detectron2.utils"))
    sys.modules["detectron2.utils.logger"] = d2l


def load(need_ext=True):
    """-> namespace(model_utils, pointnet2_utils, transformer, coarse_mod, fine_mod, ext) of the stock reference."""
    global _ns
    if _ns is not None:
        if need_ext and _ns.ext is None:
            raise RuntimeError("the reference was loaded without its extension earlier in this process")
        return _ns
    if not available():
        raise RuntimeError("reference not staged: run `python baseline/stage_reference.py` where /root/reference exists")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _stub_detectron2()
    ext = None
    from oracle import ref_ext

    if need_ext or ref_ext.available():      # the .so loads without a GPU; it only EXECUTES on one
        ext = ref_ext.load()
        if ext is None:
            raise RuntimeError("oracle/_ref/ref_pointnet2_ext.so missing (python oracle/build_ref_ext.py)")
        sys.modules["core.unopose.model.pointnet2._ext"] = ext
    else:
        import builtins

        builtins.__POINTNET2_SETUP__ = True
    pu = importlib.import_module("core.unopose.model.pointnet2.pointnet2_utils")
    if ext is not None:
        pu._ext = ext
    mu = importlib.import_module("core.unopose.utils.model_utils")
    tr = importlib.import_module("core.unopose.model.transformer")
    cm = importlib.import_module("core.unopose.model.oneref_predator_coarse_point_matching")
    fm = importlib.import_module("core.unopose.model.oneref_predator_fine_point_matching")
    assert os.path.realpath(mu.__file__).startswith(os.path.realpath(REF_ROOT)), mu.__file__
    _ns = types.SimpleNamespace(model_utils=mu, pointnet2_utils=pu, transformer=tr, coarse_mod=cm, fine_mod=fm, ext=ext)
    return _ns


@contextlib.contextmanager
def patched(ns=None, record=None):
    """INTEGRATION.md §1-2 applied to the loaded reference: `_ext` -> unopose_b200.pointnet2._ext, pose functions ->
    unopose_b200.model_utils (in model_utils itself and in the module globals that bound the names at import time).
    `record` (a list) receives (name, output) of every `_ext` call made while patched."""
    ns = ns or load()
    import unopose_b200.model_utils as new_mu
    import unopose_b200.pointnet2._ext as new_ext

    ext = new_ext if record is None else recording_ext(new_ext, record)
    saved = []

    def swap(mod, name, val):
        saved.append((mod, name, getattr(mod, name)))
        setattr(mod, name, val)

    swap(ns.pointnet2_utils, "_ext", ext)
    mods = [ns.model_utils, ns.coarse_mod, ns.fine_mod, ns.transformer]
    mods += [m for m in (getattr(ns, "model_mod", None), getattr(ns, "featx_mod", None)) if m is not None]
    for mod in mods:
        for name in POSE_NAMES:
            if hasattr(mod, name):
                swap(mod, name, getattr(new_mu, name))
    try:
        yield ns
    finally:
        for mod, name, val in reversed(saved):
            setattr(mod, name, val)


def _stub_timm():
    """`timm` is not installed here; the reference's feature extractor subclasses
    `timm.models.vision_transformer.VisionTransformer` (oneref_feature_extraction.py:12,25).  The stub provides that one
    class on top of unopose_b200.model._ViT (the same token pipeline and parameter names: patch_embed, _pos_embed,
    norm_pre, blocks, norm), so the reference's OWN `ViT.forward`, `ViT_AE`, `ViTEncoderOneRef` and `UNOPose` run
    unmodified.  The ViT is outside the parity scope — it only has to be the same network in
    the stock, patched and product runs."""
    if "timm.models.vision_transformer" in sys.modules:
        return
    from unopose_b200.model import _ViT

    class VisionTransformer(_ViT):
        def __init__(self, patch_size=16, embed_dim=768, depth=12, num_heads=12, mlp_ratio=4, qkv_bias=True,
                     init_values=None, reg_tokens=0, no_embed_class=False, norm_layer=None, **kw):
            if reg_tokens and not no_embed_class:
                raise NotImplementedError("timm stub: register tokens need no_embed_class=True")
            if not no_embed_class:
                raise NotImplementedError("timm stub: only the no_embed_class=True (DINOv2 reg4) layouts")
            super().__init__(patch_size, embed_dim, depth, num_heads, reg_tokens, init_values or 1.0, 224)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    vt = types.ModuleType("timm.models.vision_transformer")
    vt.VisionTransformer = VisionTransformer
    timm.models, models.vision_transformer = models, vt
    sys.modules.update({"timm": timm, "timm.models": models, "timm.models.vision_transformer": vt})


def load_model():
    """-> (namespace of load(), the reference's own `UNOPose` class, its module) with timm stubbed (see _stub_timm).
    The model / feature-extraction modules are added to the namespace as `model_mod` / `featx_mod` and `patched()`
    swaps the pose-function names they bound at import time, too."""
    ns = load()
    if getattr(ns, "model_mod", None) is None:
        _stub_timm()
        ns.featx_mod = importlib.import_module("core.unopose.model.oneref_feature_extraction")
        ns.model_mod = importlib.import_module("core.unopose.model.oneref_grf_predator_pose_estimation_model")
        assert os.path.realpath(ns.model_mod.__file__).startswith(os.path.realpath(REF_ROOT))
    return ns, ns.model_mod.UNOPose, ns.model_mod


def real_model_cfg(test_coarse_only=False):
    """The `model` node of configs/main_cfg.py:119-181 (values as data), DINOv2 ViT-B/14-reg4 feature extractor."""
    coarse, fine, geo = real_cfgs()
    fx = Cfg(vit_type="vit_base_patch14_reg4_dinov2", up_type="linear", embed_dim=768, out_dim=256, use_pyramid_feat=True,
             pretrained=False, vit_ckpt="")
    return Cfg(coarse_npoint=196, fine_npoint=2048, use_ref_rad=False, test_coarse_only=test_coarse_only,
               feature_extraction=fx, geo_embedding=geo, coarse_point_matching=coarse, fine_point_matching=fine)


def recording_ext(ext, record):
    """Proxy of an `_ext` module that appends (function name, output) of every call to `record`."""
    proxy = types.SimpleNamespace()
    for name in ("gather_points", "gather_points_grad", "furthest_point_sampling", "three_nn", "three_interpolate",
                 "three_interpolate_grad", "ball_query", "group_points", "group_points_grad"):
        fn = getattr(ext, name)

        def wrap(*a, _fn=fn, _name=name):
            out = _fn(*a)
            record.append((_name, out))
            return out

        setattr(proxy, name, wrap)
    return proxy


@contextlib.contextmanager
def recording(ns, record):
    """Record the stock reference's `_ext` calls (same format as patched(record=...))."""
    saved = ns.pointnet2_utils._ext
    ns.pointnet2_utils._ext = recording_ext(saved, record)
    try:
        yield ns
    finally:
        ns.pointnet2_utils._ext = saved


class Cfg(dict):
    """dict with attribute access and .get, like the OmegaConf nodes the reference passes to its modules."""
    __getattr__ = dict.__getitem__


def real_cfgs():
    """Module configs of configs/main_cfg.py:128-181 (values copied as data: shapes of the hot path)."""
    coarse = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, temp=0.1, sim_type="cosine", normalize_feat=True,
                 loss_predator_thres=0.15, loss_dis_thres=0.3, nproposal1=6000, nproposal2=300)
    fine = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, pe_radius1=0.1, pe_radius2=0.2, focusing_factor=3,
               temp=0.1, sim_type="cosine", normalize_feat=True, loss_predator_thres=0.15, loss_dis_thres=0.3,
               use_lrf=True, use_xyz=True, nsample1=64, nsample2=256)
    geo = Cfg(sigma_d=0.2, sigma_a=15, angle_k=3, reduction_a="max", hidden_dim=256)
    return coarse, fine, geo
