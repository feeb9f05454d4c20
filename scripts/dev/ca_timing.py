"""Phase timing of k_coarse_assign_exact (SM cycles between its cluster barriers)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from unopose_b200 import _lib as L
from unopose_b200.pipeline import HotPathConfig, synthetic_inputs
from unopose_b200 import model_utils as MU
lib = L.load(); dev = torch.device("cuda:0"); cfg = HotPathConfig(); B = 16
inp = synthetic_inputs(0, B, cfg, device=dev)
att = MU.compute_feature_similarity(inp["c_f1"], inp["c_f2"], "cosine", 0.1, True)
n = 196
w1 = torch.empty(B, n, device=dev); w2 = torch.empty(B, n, device=dev); cdf = torch.empty(B, n * n, device=dev)
st = torch.zeros(16, dtype=torch.int64, device=dev)
sc = inp["c_score"]
for it in range(3):
    rc = lib.upk_coarse_assignment_profile(L.ptr(att), L.ptr(sc), sc.shape[1], sc[:, n:].data_ptr(), sc.shape[1], B, n, n, L.ptr(w1), L.ptr(w2), L.ptr(cdf), L.ptr(st), L.stream_ptr(att))
    torch.cuda.synchronize()
    s = st.cpu().tolist()
    print("rc", rc, "phases (cycles):", [s[i + 1] - s[i] for i in range(8)], "total", s[8] - s[0])
names = ["load+rowstats+colmax", "cmax+expcol", "colsum gather+chain", "A+labels", "w2+P->global", "scan carry-free", "carry chain", "finalize"]
print(names)
