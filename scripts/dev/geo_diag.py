import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import geo_oracle as G
from unopose_b200.modules import geo
dev = torch.device("cuda:0")
pts, dterm, w_d, b_d, w_a, b_a = G.make_inputs(297, 1, 197, 256, dev)
fa = 180.0 / (15 * math.pi)
d, a = geo.geometric_embedding_indices(pts, 0.2, fa, 3)
d_ref, a_ref = G.embedding_indices(pts, 0.2, 15, 3)
diff = (a - a_ref).abs()
print("frac elems > 2e-4", (diff > 2e-4).float().mean().item(), "max", diff.max().item())
bad = (diff > 2e-4).flatten(2).any(2)[0]
i = int(bad.nonzero()[0])
print("row", i, "bad cols", (diff[0, i] > 2e-4).any(1).nonzero().flatten()[:20].tolist())
knn_ref = (d_ref).topk(4, dim=2, largest=False)[1][0, i]
knn_my = (d).topk(4, dim=2, largest=False)[1][0, i]
print("knn ref", knn_ref.tolist(), "knn mine(from d)", knn_my.tolist(), d_ref[0, i, knn_ref].tolist())
j = int((diff[0, i] > 2e-4).any(1).nonzero()[0])
print("j", j, "mine", a[0, i, j].tolist(), "ref", a_ref[0, i, j].tolist())
# which permutation?
for r in range(3):
    print("col", r, "mine vs ref cols:", [(a[0, i, :, r] - a_ref[0, i, :, c]).abs().max().item() for c in range(3)])
# cpu reference
d_c, a_c = G.embedding_indices(pts.cpu(), 0.2, 15, 3)
print("cpu vs gpu oracle: a max", (a_c - a_ref.cpu()).abs().max().item(), "mine vs cpu", (a.cpu() - a_c).abs().max().item(),
      "frac", ((a.cpu() - a_c).abs() > 2e-4).float().mean().item())
