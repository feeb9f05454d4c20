"""CPU experiment behind DESIGN.md §9 item 0: error of split-precision GEMM schemes on the similarity logits
(normalised 256-d features, temp 0.1) and on the geometric embedding operands, against fp64.

    3xTF32               a_hi b_hi + a_hi b_lo + a_lo b_hi, tf32 operands           (what the kernels run today)
    3xFP16, scale 2^12   the same three products with fp16 operands of x * 4096      (kind::f16: 2x the MMA rate, half the bytes)
"FTZ" = fp16 subnormals flushed to zero (worst case for hardware that does not honour them).
"""
import torch

torch.manual_seed(0)
n, c = 2048, 256
a = torch.nn.functional.normalize(torch.randn(n, c), dim=1)
b = torch.nn.functional.normalize(torch.randn(n, c) + 0.5 * torch.randn(1, c), dim=1)
ex = (a.double() @ b.double().T) / 0.1


def tf32(x):
    return ((x.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


ah, bh = tf32(a.clone()), tf32(b.clone())
al, bl = tf32((a - ah).clone()), tf32((b - bh).clone())
print("fp32 matmul           %.3g" % ((a @ b.T) / 0.1).double().sub(ex).abs().max())
print("1xTF32                %.3g" % ((ah @ bh.T) / 0.1).double().sub(ex).abs().max())
print("3xTF32                %.3g" % ((ah @ bh.T + ah @ bl.T + al @ bh.T) / 0.1).double().sub(ex).abs().max())
ftz = lambda x: torch.where(x.abs() < 6.103515625e-05, torch.zeros_like(x), x)  # noqa: E731
for sc in (1.0, 256.0, 4096.0):
    A, B = a * sc, b * sc
    h1, h2 = A.half().float(), B.half().float()
    l1, l2 = (A - h1).half().float(), (B - h2).half().float()
    d = (h1 @ h2.T + h1 @ l2.T + l1 @ h2.T) / (sc * sc) / 0.1
    df = (ftz(h1) @ ftz(h2).T + ftz(h1) @ ftz(l2).T + ftz(l1) @ ftz(h2).T) / (sc * sc) / 0.1
    print("3xFP16 scale %-6g   %.3g   (FTZ: %.3g)" % (sc, d.double().sub(ex).abs().max(), df.double().sub(ex).abs().max()))
