"""Dev tool: accuracy + timing of the tcgen05 similarity GEMM against fp64."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unopose_b200 import _lib, model_utils as MU

dev = torch.device("cuda:0")
lib = _lib.load()
torch.manual_seed(0)
for (B, n, m, c) in ((1, 2049, 2049, 256), (2, 600, 900, 64), (16, 2049, 2049, 256)):
    f1 = torch.randn(B, n, c, device=dev)
    f2 = torch.randn(B, m, c, device=dev)
    ref = (torch.nn.functional.normalize(f1.double(), dim=2) @ torch.nn.functional.normalize(f2.double(), dim=2).transpose(1, 2)) / 0.1
    for mode in (0, 3, 1):
        lib.upk_set_similarity_mode(mode)
        out = MU.compute_feature_similarity(f1, f2, "cosine", 0.1, True)
        torch.cuda.synchronize()
        err = (out.double() - ref).abs().max().item()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5):
            MU.compute_feature_similarity(f1, f2, "cosine", 0.1, True)
        t1.record(); torch.cuda.synchronize()
        print("B=%d n=%d m=%d c=%d mode=%d  max|err|=%.3e  %.1f us" % (B, n, m, c, mode, err, t0.elapsed_time(t1) / 5 * 1e3), flush=True)
lib.upk_set_similarity_mode(3)
