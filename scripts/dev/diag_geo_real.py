"""Diagnostic: reference GeometricStructureEmbedding (GPU / CPU) vs the fused kernel on FPS-sampled surface points."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from baseline import refgpu
from unopose_b200.modules import geo as G
from unopose_b200.synthetic import matching_batch
from util_state import keyed_state_dict
ref = refgpu.load()
dev = torch.device("cuda:0")
cc, cf, cg = refgpu.real_cfgs()
d = matching_batch(62, 2, 2048, 256)
p1 = torch.from_numpy(d["pts1"]).to(dev); f1 = torch.from_numpy(d["f1"][:, 1:].copy()).to(dev)
sp1, sf1, i1 = ref.model_utils.sample_pts_feats(p1, f1, 196, True)
pts = torch.cat([torch.ones(2, 1, 3, device=dev), sp1], 1)
m = ref.transformer.GeometricStructureEmbedding(cg).eval()
m.load_state_dict(keyed_state_dict(m.state_dict(), 5))
mc = m
mg = ref.transformer.GeometricStructureEmbedding(cg).eval(); mg.load_state_dict(m.state_dict()); mg = mg.to(dev)
print("tf32 flags", torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.get_float32_matmul_precision())
with torch.no_grad():
    dg, ag = mg.get_embedding_indices(pts)
    dc, ac = mc.get_embedding_indices(pts.cpu())
    do, ao = G.geometric_embedding_indices(pts, cg.sigma_d, mg.factor_a, cg.angle_k)
    eg = mg(pts); ec = mc(pts.cpu())
    eo = G.geometric_embedding(pts, mg.embedding.div_term, mg.proj_d.weight, mg.proj_d.bias, mg.proj_a.weight, mg.proj_a.bias, cg.sigma_d, mg.factor_a, cg.angle_k, "max")
    # the GPU module's own pieces
    de = mg.embedding(dg); pd = mg.proj_d(de)
    pd64 = torch.nn.functional.linear(de.double(), mg.proj_d.weight.double(), mg.proj_d.bias.double())
off = ~torch.eye(197, dtype=torch.bool, device=dev)
print("d idx: ref-gpu vs ref-cpu %.3e | ours vs ref-gpu %.3e | ours vs ref-cpu %.3e (off-diagonal)" % (
    (dg - dc.to(dev))[:, off].abs().max(), (do - dg)[:, off].abs().max(), (do - dc.to(dev))[:, off].abs().max()))
print("a idx: ref-gpu vs ref-cpu %.3e | ours vs ref-gpu %.3e | ours vs ref-cpu %.3e" % (
    (ag - ac.to(dev))[:, off].abs().max(), (ao - ag)[:, off].abs().max(), (ao - ac.to(dev))[:, off].abs().max()))
print("emb: ref-gpu vs ref-cpu %.3e | ours vs ref-gpu %.3e | ours vs ref-cpu %.3e" % (
    (eg - ec.to(dev)).abs().max(), (eo - eg).abs().max(), (eo - ec.to(dev)).abs().max()))
print("proj_d on GPU vs fp64: %.3e" % (pd.double() - pd64).abs().max())
rows = (eo - eg).abs().amax(dim=(2, 3))
print("rows off >1e-4 ours-vs-refgpu:", int((rows > 1e-4).sum()), " ours-vs-refcpu:", int(((eo - ec.to(dev)).abs().amax(dim=(2, 3)) > 1e-4).sum()),
      " refgpu-vs-refcpu:", int(((eg - ec.to(dev)).abs().amax(dim=(2, 3)) > 1e-4).sum()))
b, i = [int(v) for v in torch.nonzero(rows > 1e-4)[0]]
j = int((eo - eg)[b, i].abs().amax(dim=1).argmax())
print("example row", b, i, "col", j, "d", float(dg[b, i, j]), float(do[b, i, j]), "a", ag[b, i, j].tolist(), ao[b, i, j].tolist())
print("diag a ref-gpu", ag[b, i, i].tolist(), "ours", ao[b, i, i].tolist(), "cpu", ac[b, i, i].tolist())
