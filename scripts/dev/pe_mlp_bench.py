"""Dev: device time of upk_shared_mlp_max alone at the PositionalEncoding shapes (B x 2048 centres, ns = 64 / 256,
SharedMLP 6 -> 32 -> 64 -> 128).  UPK_PE_MLP_SLOTS=2 selects the round-1 two-slot kernel.
    python scripts/dev/pe_mlp_bench.py [B] [--once]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from unopose_b200 import _lib as L  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else 16
once = "--once" in sys.argv
dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = L.load()
res = {"B": B, "slots": os.environ.get("UPK_PE_MLP_SLOTS", "4")}
ws = [torch.randn(32, 6, device=dev) * 0.4, torch.randn(64, 32, device=dev) * 0.2, torch.randn(128, 64, device=dev) * 0.15]
bs = [torch.randn(32, device=dev) * 0.1, torch.randn(64, device=dev) * 0.1, torch.randn(128, device=dev) * 0.1]
for ns in (64, 256):
    x = torch.randn(B, 6, 2048, ns, device=dev)
    out = torch.empty(B, 128, 2048, device=dev)

    def run():
        L.check(lib.upk_shared_mlp_max(x.data_ptr(), B, 6, 2048, ns, 32, 64, 128, ws[0].data_ptr(), bs[0].data_ptr(),
                                       ws[1].data_ptr(), bs[1].data_ptr(), ws[2].data_ptr(), bs[2].data_ptr(), out.data_ptr(),
                                       torch.cuda.current_stream().cuda_stream), "mlp")
    run()
    torch.cuda.synchronize()
    if once:
        continue
    # fp32 torch layers on a slice (numerics sanity)
    xs = x[:1, :, :64]
    h = xs
    for w, b in zip(ws, bs):
        h = torch.relu(torch.einsum("oc,bcms->boms", w.double(), h.double()) + b.double().view(1, -1, 1, 1))
    refmax = h.max(dim=3)[0].float()
    err = float((out[:1, :, :64] - refmax).abs().max())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    tiles = B * 2048 * ns // 128
    flop = 2.0 * B * 2048 * ns * (6 * 32 + 32 * 64 + 64 * 128)
    res["ns%d" % ns] = {"ms": ms, "us_per_tile_per_sm": ms * 1e3 / (tiles / 148), "algorithmic_tflops": flop / ms / 1e9,
                        "max_abs_err_vs_fp64": err}
print(json.dumps(res))
