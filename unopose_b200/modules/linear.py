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
