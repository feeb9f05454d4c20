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
# event timing of the kernel alone (stage-wise entry), 50 launches back to back
ws = torch.empty(max(lib.upk_coarse_assignment_workspace_bytes(B, n, n), 256), dtype=torch.uint8, device=dev)
def run():
    L.check(lib.upk_coarse_assignment(L.ptr(att), L.ptr(sc), sc.shape[1], sc[:, n:].data_ptr(), sc.shape[1], B, n, n, L.ptr(ws), ws.numel(),
                                      L.ptr(w1), L.ptr(w2), L.ptr(cdf), L.stream_ptr(att)), "ca")
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for Bx in (16,):
    e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    print("k_coarse_assign_exact: %.1f us per launch (B=%d, events over 50 launches)" % (e0.elapsed_time(e1) / 50 * 1e3, Bx))
import ctypes
try:
    cudart = ctypes.CDLL("libcudart.so")
except OSError:
    cudart = None
