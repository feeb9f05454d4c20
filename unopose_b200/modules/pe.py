"""Fused SharedMLP + max over the ball (the MLP half of `PositionalEncoding`): host wrapper of
`upk_shared_mlp_max`.  Reference: oneref_predator_fine_point_matching.py:167-176 (`self.mlp1(group(...)).max(dim=3)[0]`)
with SharedMLP of pointnet2/pytorch_utils.py:25-48.  CUDA tensors only.
"""
import torch

from .. import _lib as L


def fold_shared_mlp(mlp):
    """[(W', b')] per layer of an eval-mode SharedMLP: the batch norm folded into the 1x1 conv
    (y = relu(bn(conv(x))) = relu(W' x + b'))."""
    out = []
    for layer in mlp:
        conv = layer.conv
        w = conv.weight.detach().reshape(conv.weight.shape[0], -1).float()
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
        if hasattr(layer, "normlayer"):
            bn = layer.normlayer.bn
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)
            w = w * scale[:, None]
            b = (b - bn.running_mean.float()) * scale + bn.bias.detach().float()
        if not hasattr(layer, "activation"):
            raise L.UnoposeNativeError("shared_mlp_max: every layer must end in ReLU")
        out.append((w.contiguous(), b.contiguous()))
    return out


def supported(mlp, x):
    if len(mlp) != 3 or x.dim() != 4:
        return False
    c = [mlp[0].conv.weight.shape[1]] + [layer.conv.weight.shape[0] for layer in mlp]
    return bool(L.load().upk_shared_mlp_max_supported(c[0], c[1], c[2], c[3], x.shape[2], x.shape[3]))


def shared_mlp_max(x, mlp):
    """x (B,cin,m,ns) -> (B,C3,m) = mlp(x).max(dim=3)[0] for an eval-mode 3-layer SharedMLP."""
    L.check_cuda(x, "x")
    if mlp.training:
        raise L.UnoposeNativeError("shared_mlp_max: the batch norm must be in eval mode (running statistics)")
    if not supported(mlp, x):
        raise L.UnoposeNativeError("shared_mlp_max: geometry not supported by the fused kernel")
    x = x.float().contiguous()
    (w1, b1), (w2, b2), (w3, b3) = fold_shared_mlp(mlp)
    B, cin, m, ns = x.shape
    out = torch.empty((B, w3.shape[0], m), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        L.check(L.load().upk_shared_mlp_max(L.ptr(x), B, cin, m, ns, w1.shape[0], w2.shape[0], w3.shape[0], L.ptr(w1),
                                            L.ptr(b1), L.ptr(w2), L.ptr(b2), L.ptr(w3), L.ptr(b3), L.ptr(out),
                                            L.stream_ptr(x)), "shared_mlp_max")
    return out
