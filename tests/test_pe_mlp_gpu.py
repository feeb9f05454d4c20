"""GPU: the fused SharedMLP + max kernel (unopose_b200/csrc/pe_mlp.cu) against the torch layers it
replaces (1x1 Conv2d + eval BatchNorm2d + ReLU, x3, max over the ball) — the same cuDNN/cuBLAS ops the reference runs —
and the PositionalEncoding module against the reference's output (golden `pe_p2`).

Tolerance: 3xTF32 products accumulate in fp32 like SGEMM; activations are O(1..10) after three layers with
kaiming-initialised weights: 2e-5 relative to the largest activation, and no further from an fp64 evaluation than
the fp32 torch path is (x4 + 1e-6 relative).
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_torch_layers():
    """The torch layers are the fp32 comparison: cuDNN would otherwise run the 1x1 convs in TF32 (torch's default
    `cudnn.allow_tf32 = True`, which the reference never changes: its GPU path has 5e-4 relative error here; ours keeps
    fp32-level accuracy with 3xTF32, i.e. it is the more exact of the two and matches the reference's CPU path)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _mlp(cin, dev, seed):
    from unopose_b200.modules.layers import SharedMLP

    torch.manual_seed(seed)
    mlp = SharedMLP([cin, 32, 64, 128], bn=True)
    for layer in mlp:                       # non-trivial batch-norm statistics and affine parameters
        bn = layer.normlayer.bn
        bn.running_mean.normal_(0, 0.3)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.data.uniform_(0.5, 1.5)
        bn.bias.data.normal_(0, 0.2)
    return mlp.to(dev).eval()


@pytest.mark.parametrize("B,cin,m,ns", [(2, 6, 64, 64), (1, 6, 2048, 256), (3, 3, 40, 32), (2, 9, 17, 128),
                                        (1, 6, 5, 384), (2, 6, 2048, 64)])
def test_fused_mlp_max_vs_torch(cuda, B, cin, m, ns):
    from unopose_b200 import _lib
    from unopose_b200.modules import pe

    mlp = _mlp(cin, cuda, 10 + ns)
    x = torch.randn(B, cin, m, ns, device=cuda)
    x[:, :, :, ns // 2:] = x[:, :, :, :1]          # padded balls repeat the first hit, like ball_query
    n0 = _lib.launch_count()
    with torch.no_grad():
        got = pe.shared_mlp_max(x, mlp)
        assert _lib.launch_count() - n0 == 1
        ref = mlp(x).max(dim=3)[0]
        ref64 = mlp.double()(x.double()).max(dim=3)[0]
        mlp.float()
    scale = ref64.abs().max().item()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 2e-5 * scale
    torch.backends.cudnn.allow_tf32 = True
    with torch.no_grad():
        ref_tf32 = mlp.float()(x).max(dim=3)[0]     # what the reference's GPU path computes with torch's defaults
    torch.backends.cudnn.allow_tf32 = False
    assert (got - ref_tf32).abs().max().item() < 5e-3 * scale
    e_ours, e_torch = (got.double() - ref64).abs().max().item(), (ref.double() - ref64).abs().max().item()
    assert e_ours < 4 * e_torch + 1e-6 * scale, (e_ours, e_torch, scale)


def test_unsupported_geometries_and_modes(cuda):
    from unopose_b200 import _lib
    from unopose_b200.modules import pe

    mlp = _mlp(6, cuda, 1)
    assert not pe.supported(mlp, torch.zeros(1, 6, 8, 48, device=cuda))      # 48 neither divides nor is a multiple of 128
    assert not pe.supported(mlp, torch.zeros(1, 6, 3, 32, device=cuda))      # 96 samples: not a whole tile
    assert pe.supported(mlp, torch.zeros(1, 6, 4, 32, device=cuda))
    with pytest.raises(_lib.UnoposeNativeError):
        pe.shared_mlp_max(torch.zeros(1, 6, 8, 48, device=cuda), mlp)
    with pytest.raises(_lib.UnoposeNativeError):
        pe.shared_mlp_max(torch.zeros(1, 6, 4, 32, device=cuda), mlp.train())
    with pytest.raises(RuntimeError, match="CPU not supported"):
        pe.shared_mlp_max(torch.zeros(1, 6, 4, 32), mlp.eval())


def test_positional_encoding_module_uses_fused_kernel(cuda):
    """PositionalEncoding with the reference's weights: the fused branch (eval, no grad) against the torch branch of
    the same module on the same grouped inputs."""
    from unopose_b200.modules import FinePointMatchingOneRef

    class Cfg(dict):
        __getattr__ = dict.__getitem__

    g = torch.load(os.path.join(GOLD, "modules_small.pt"), weights_only=False)
    m = FinePointMatchingOneRef(Cfg(g["cfg_fine"])).to(cuda).eval()
    m.load_state_dict({k: v.to(cuda) for k, v in g["sd_fine"].items()})
    p2 = g["p2"].to(cuda).contiguous()
    with torch.no_grad():
        fused = m.PE(p2)
    with torch.enable_grad():                      # the torch branch (autograd on)
        plain = m.PE(p2).detach()
    assert fused.shape == plain.shape == (2, p2.shape[1], 32)
    assert (fused - plain).abs().max().item() < 2e-5 * max(1.0, plain.abs().max().item())
