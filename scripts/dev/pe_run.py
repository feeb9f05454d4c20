"""Dev tool: one PositionalEncoding forward at the real config (for ncu captures of k_lrf_group / k_shared_mlp_max)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from unopose_b200.modules.matching import PositionalEncoding  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
torch.manual_seed(0)
pe = PositionalEncoding(256, r1=0.1, r2=0.2, nsample1=64, nsample2=256, use_lrf=True, use_xyz=True).to(dev).eval()
p = torch.randn(B, 2048, 3, device=dev)
p = p / p.norm(dim=2).max(dim=1)[0].view(B, 1, 1)
with torch.no_grad():
    for _ in range(2):
        out = pe(p)
torch.cuda.synchronize()
print(out.shape, float(out.abs().mean()))
