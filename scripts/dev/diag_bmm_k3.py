"""Diagnostic: how does torch.matmul (cuBLAS) on the GPU round a K = 3 dot product?  Dumps operands and results for an
offline search over candidate FMA orders (gpurun_out/bmm_k3.npz)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
out = {}
for name, B, n, m in (("geo", 2, 197, 197), ("score", 600, 196, 196), ("fine", 16, 2048, 2048), ("one", 1, 196, 196)):
    x = torch.randn(B, n, 3, generator=g).to(dev)
    y = torch.randn(B, m, 3, generator=g).to(dev) if name != "geo" else x
    xy = torch.matmul(x, y.transpose(-1, -2))
    x2 = torch.sum(x ** 2, dim=-1)
    sl = slice(0, 4)
    out[name + "_x"] = x[sl].cpu().numpy(); out[name + "_y"] = y[sl].cpu().numpy()
    out[name + "_xy"] = xy[sl].cpu().numpy(); out[name + "_x2"] = x2[sl].cpu().numpy()
    # (pts - t) @ R as in the solvers
R = torch.linalg.qr(torch.randn(600, 3, 3, generator=g))[0].to(dev)
p = torch.randn(600, 196, 3, generator=g).to(dev)
out["rot_p"] = p[:4].cpu().numpy(); out["rot_R"] = R[:4].cpu().numpy(); out["rot_out"] = (p @ R)[:4].cpu().numpy()
R16 = R[:16].contiguous(); p16 = torch.randn(16, 2048, 3, generator=g).to(dev)
out["rot16_p"] = p16[:2].cpu().numpy(); out["rot16_R"] = R16[:2].cpu().numpy(); out["rot16_out"] = (p16 @ R16)[:2].cpu().numpy()
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "bmm_k3.npz"), **out)
print("saved")
