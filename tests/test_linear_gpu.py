"""GPU: the transformer-block arithmetic of the matching modules that runs on our kernels — `upk_linear` (3xTF32 tcgen05
GEMM with bias / ReLU epilogue, modules/linear.py) and `upk_rpe_scores` (modules/transformer.py::_rpe_scores) — against
fp64 evaluations and the fp32 torch ops they replace (reference: core/unopose/model/transformer.py:94-201, 392-395)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fp32_matmul():
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.parametrize("rows,cin,cout,relu", [(6304, 256, 256, False), (65536, 256, 512, True), (65536, 512, 256, False),
                                                (32 * 2049, 256, 256, False), (1500, 64, 48, True), (4096, 256, 16, False)])
def test_tc_linear_fp32_level_accuracy(cuda, fp32_matmul, rows, cin, cout, relu):
    from unopose_b200 import _lib
    from unopose_b200.modules.linear import linear

    g = torch.Generator().manual_seed(rows + cin)
    layer = nn.Linear(cin, cout).to(cuda).eval()
    x = (torch.randn(rows, cin, generator=g) * 1.7).to(cuda)
    n0 = _lib.launch_count()
    with torch.no_grad():
        y = linear(layer, x, relu=relu)
        ref32 = F.linear(x, layer.weight, layer.bias)
        ref64 = F.linear(x.double(), layer.weight.double(), layer.bias.double())
        if relu:
            ref32, ref64 = F.relu(ref32), F.relu(ref64)
    assert _lib.launch_count() - n0 == 2            # operand split + GEMM: the kernel path, not F.linear
    assert y.shape == ref32.shape and y.dtype == torch.float32
    scale = float(ref64.abs().max())
    e_ours, e_cublas = float((y.double() - ref64).abs().max()), float((ref32.double() - ref64).abs().max())
    print("linear %dx%d->%d: max err ours %.2e, cuBLAS fp32 %.2e (scale %.2f)" % (rows, cin, cout, e_ours, e_cublas, scale))
    assert e_ours <= 2e-6 * scale + 4 * e_cublas


def test_tc_linear_dispatch_and_shapes(cuda):
    from unopose_b200 import _lib
    from unopose_b200.modules.linear import linear

    layer = nn.Linear(256, 256).to(cuda).eval()
    x = torch.randn(4, 2049, 256, device=cuda)
    with torch.no_grad():
        y = linear(layer, x)
        assert y.shape == (4, 2049, 256) and torch.allclose(y, layer(x), atol=2e-5)
        n0 = _lib.launch_count()
        small = linear(layer, x[:, :10])                     # below MIN_ROWS: torch
        assert _lib.launch_count() == n0 and torch.equal(small, layer(x[:, :10]))
        xt = x.transpose(0, 1)                               # non-contiguous input
        assert torch.allclose(linear(layer, xt), layer(xt), atol=2e-5)
    x.requires_grad_(True)
    n0 = _lib.launch_count()
    out = linear(layer, x)                                   # autograd: torch's own layer, differentiable
    assert _lib.launch_count() == n0 and out.requires_grad
    head = nn.Linear(256, 1).to(cuda).eval()                 # score heads: out_features < 16 -> torch
    with torch.no_grad():
        assert torch.equal(linear(head, x.detach()), head(x.detach()))


@pytest.mark.parametrize("B,N,C", [(3, 197, 256), (2, 50, 128)])
def test_rpe_scores_kernel(cuda, fp32_matmul, B, N, C):
    from unopose_b200 import _lib
    from unopose_b200.modules.transformer import _rpe_scores

    g = torch.Generator().manual_seed(N)
    embed = torch.randn(B, N, N, C, generator=g).to(cuda)
    q2 = torch.randn(B, N, C, 4, generator=g).to(cuda)
    n0 = _lib.launch_count()
    with torch.no_grad():
        got = _rpe_scores(embed, q2)
    assert _lib.launch_count() - n0 == 1
    ref64 = torch.matmul(embed.double(), q2.double()).permute(0, 3, 1, 2)
    ref32 = torch.matmul(embed, q2).permute(0, 3, 1, 2)
    e_ours, e_torch = float((got.double() - ref64).abs().max()), float((ref32.double() - ref64).abs().max())
    print("rpe scores: max err ours %.2e, torch fp32 %.2e" % (e_ours, e_torch))
    assert got.shape == (B, 4, N, N) and e_ours <= 2 * e_torch + 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("rows", [6304, 65536])
def test_linear_multi_bit_identical_to_single_layers(cuda, rows):
    """`linear_multi` (q / k / v projections of an attention block as ONE operand split + ONE GEMM over the
    concatenated weights, transformer.py:116-118 / :371-373) returns exactly what `linear(layer, x)` returns per layer;
    the cached concatenation follows in-place parameter updates."""
    import torch.nn as nn
    from unopose_b200.modules.linear import linear, linear_multi
    torch.manual_seed(3)
    layers = [nn.Linear(256, 256).to(cuda) for _ in range(3)]
    x = torch.randn(rows // 197 if rows == 6304 else 32, 197 if rows == 6304 else 2048, 256, device=cuda)
    with torch.no_grad():
        for n in (3, 2):
            outs = linear_multi(layers[:n], x)
            for l, o in zip(layers[:n], outs):
                assert o.shape == x.shape[:-1] + (256,)
                assert torch.equal(o, linear(l, x))
        layers[1].weight.mul_(0.5)                       # bumps the version counter: the cache must be rebuilt
        outs = linear_multi(layers, x)
        assert torch.equal(outs[1], linear(layers[1], x)) and torch.equal(outs[2], linear(layers[2], x))
