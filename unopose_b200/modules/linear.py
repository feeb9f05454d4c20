"""`nn.Linear` forward on the 3xTF32 tcgen05 GEMM (`upk_linear`) for the transformer blocks of the matching modules.

The reference runs every `Linear` of core/unopose/model/transformer.py as an fp32 SIMT SGEMM (TF32 is switched off at
its entry point, main_unopose.py:139-141); at the real config these are 65 536 x 256 x 256/512 GEMMs per dense layer
call and ~150 small ones per forward — 40 % of the matching forward's device time once the hot path is on kernels
(profiles/r2_config3_forward.json).  `linear(layer, x)` keeps the module structure and parameter names (the layer IS
the `nn.Linear`) and only replaces the arithmetic: fp32-level accuracy (3xTF32 split, ~1e-6 relative), bias and the
optional ReLU applied in the GEMM's epilogue.  Evaluation path on CUDA; anything else (autograd, CPU tensors, shapes the
kernel does not take) goes to `F.linear`, the same layer computed by torch.
"""
import torch
import torch.nn.functional as F

from .. import _lib as L

MIN_ROWS = 1024       # below this a cuBLAS SGEMM launch is as fast as split + GEMM


def _eligible(layer, x):
    return (x.is_cuda and x.dtype == torch.float32 and not (torch.is_grad_enabled() and (x.requires_grad or layer.weight.requires_grad))
            and layer.in_features % 16 == 0 and layer.out_features >= 16 and x.numel() // layer.in_features >= MIN_ROWS
            and layer.weight.dtype == torch.float32 and layer.weight.is_cuda)


def linear(layer, x, relu=False):
    """relu?(x @ layer.weight.T + layer.bias) with x (..., in_features)."""
    if not _eligible(layer, x):
        y = F.linear(x, layer.weight, layer.bias)
        return F.relu(y) if relu else y
    lib = L.load()
    x2 = x.reshape(-1, layer.in_features)
    if not x2.is_contiguous() or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    w = layer.weight if layer.weight.is_contiguous() else layer.weight.contiguous()
    rows = x2.shape[0]
    y = torch.empty((rows, layer.out_features), dtype=torch.float32, device=x.device)
    ws = torch.empty(lib.upk_linear_workspace_bytes(rows, layer.in_features, layer.out_features), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.upk_linear(L.ptr(x2), L.ptr(w), L.ptr(layer.bias), rows, layer.in_features, layer.out_features, int(relu),
                            L.ptr(ws), ws.numel(), L.ptr(y), L.stream_ptr(x))
    if rc == -2:          # UPK_ERR_UNSUPPORTED: the GEMM does not take this call
        y = F.linear(x, layer.weight, layer.bias)
        return F.relu(y) if relu else y
    L.check(rc, "linear")
    return y.view(*x.shape[:-1], layer.out_features)


def _cat_params(layers):
    """Concatenated (weight, bias) of several `nn.Linear`s with the same input, cached on the first layer and rebuilt
    whenever one of the parameters changes (load_state_dict, optimiser step, .to())."""
    key = tuple((l.weight.data_ptr(), l.weight._version, l.bias.data_ptr(), l.bias._version) for l in layers)
    cache = getattr(layers[0], "_upk_cat_cache", None)
    if cache is None or cache[0] != key:
        with torch.no_grad():
            cache = (key, torch.cat([l.weight for l in layers], 0).contiguous(), torch.cat([l.bias for l in layers], 0).contiguous())
        layers[0]._upk_cat_cache = cache
    return cache[1], cache[2]


def linear_multi(layers, x):
    """[layer(x) for layer in layers] for `nn.Linear`s that share their input: ONE operand split and ONE GEMM over the
    concatenated weights (q / k / v projections of an attention block) instead of one of each per layer.  Every output
    element goes through the same MMA sequence as in `linear(layer, x)`, so the values are bit-identical; the outputs
    are column slices of one (rows, sum out_features) tensor."""
    layers = list(layers)
    if len(layers) == 1:
        return [linear(layers[0], x)]
    same = all(l.in_features == layers[0].in_features and l.bias is not None for l in layers)
    if not same or not all(_eligible(l, x) for l in layers):
        return [linear(l, x) for l in layers]
    lib = L.load()
    inf = layers[0].in_features
    outs = [l.out_features for l in layers]
    w, bias = _cat_params(layers)
    x2 = x.reshape(-1, inf)
    if not x2.is_contiguous() or x2.data_ptr() % 16:
        x2 = x2.contiguous()
    rows, total = x2.shape[0], sum(outs)
    y = torch.empty((rows, total), dtype=torch.float32, device=x.device)
    ws = torch.empty(lib.upk_linear_workspace_bytes(rows, inf, total), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.upk_linear(L.ptr(x2), L.ptr(w), L.ptr(bias), rows, inf, total, 0, L.ptr(ws), ws.numel(), L.ptr(y), L.stream_ptr(x))
    if rc == -2:
        return [linear(l, x) for l in layers]
    L.check(rc, "linear")
    y = y.view(*x.shape[:-1], total)
    res, o = [], 0
    for n in outs:
        res.append(y[..., o:o + n])
        o += n
    return res
