import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unopose_b200.modules import FinePointMatchingOneRef
class Cfg(dict):
    __getattr__ = dict.__getitem__
dev = torch.device("cuda:0")
g = torch.load(os.path.join(ROOT, "tests/golden/modules_small.pt"), weights_only=False)
g = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in g.items()}
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
m = FinePointMatchingOneRef(Cfg(g["cfg_fine"]), return_feat=True).to(dev).eval()
m.load_state_dict(g["sd_fine"])
with torch.no_grad():
    grp = m.PE.group1(g["p2"].contiguous(), g["p2"].contiguous(), g["p2"].transpose(1, 2).contiguous())
    d = (grp - g["grp_p2"]).abs()
    print("group1 diff max", d.max().item(), "frac>1e-4", (d > 1e-4).float().mean().item())
    # which channels
    print("per-channel max", d.amax(dim=(0, 2, 3)).tolist())
    pe = m.PE(g["p2"])
    print("PE diff max", (pe - g["pe_p2"]).abs().max().item())
    ep0 = {"init_R": g["init_R"], "init_t": g["init_t"]}
    ep, g1, g2 = m(g["p1"], g["f1"], g["geo1"], g["fps1"], g["p2"], g["f2"], g["geo2"], g["fps2"], g["radius"], ep0)
    print("g1 diff", (g1 - g["fine_g1"]).abs().max().item(), "g2 diff", (g2 - g["fine_g2"]).abs().max().item())
