import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from unopose_b200.modules import FinePointMatchingOneRef, CoarsePointMatchingOneRef
from unopose_b200.modules.matching import PositionalEncoding
class Cfg(dict):
    __getattr__ = dict.__getitem__
B=16; dev=torch.device("cuda:0"); torch.manual_seed(0)
cf = Cfg(nblock=3, input_dim=256, hidden_dim=256, out_dim=256, pe_radius1=0.1, pe_radius2=0.2, focusing_factor=3,
         temp=0.1, sim_type="cosine", normalize_feat=True, use_lrf=True, use_xyz=True, nsample1=64, nsample2=256)
fine = FinePointMatchingOneRef(cf).to(dev).eval()
p = torch.randn(B, 2048, 3, device=dev); p = p / p.norm(dim=2).max(dim=1)[0].view(B,1,1)
f = torch.randn(B, 2048, 256, device=dev)
geo = torch.randn(B, 197, 197, 256, device=dev)
idx = torch.stack([torch.randperm(2048, device=dev)[:196] for _ in range(B)]).int()
radius = torch.ones(B, device=dev)
def run():
    with torch.no_grad():
        return fine(p, f, geo, idx, p.clone(), f.clone(), geo, idx, radius, {})
run(); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity, record_function
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    run(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
# PE alone
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/it
with torch.no_grad():
    print("PE ms", t(lambda: fine.PE(p)))
    g1 = fine.PE.group1
    feats = p.transpose(1,2).contiguous()
    print("group1 ms", t(lambda: g1(p, p, feats)), "group2 ms", t(lambda: fine.PE.group2(p, p, feats)))
    x2 = fine.PE.group2(p, p, feats)
    print("mlp2 ms", t(lambda: fine.PE.mlp2(x2).max(dim=3)[0]), x2.shape)
