"""Dev tool: time the fused GeometricStructureEmbedding against the reference torch-op sequence on this GPU.
    python scripts/geo_bench.py [B] [N] [C]
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unopose_b200.modules import GeometricStructureEmbedding, geo  # noqa: E402


def timeit(fn, it=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 197
    C = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    mod = GeometricStructureEmbedding(dict(sigma_d=0.2, sigma_a=15, angle_k=3, reduction_a="max", hidden_dim=C)).to(dev).eval()
    p = torch.randn(B, N - 1, 3, device=dev)
    pts = torch.cat([torch.ones(B, 1, 3, device=dev), p / p.norm(dim=2).max(dim=1)[0].view(B, 1, 1)], 1)
    dterm, w_d, b_d, w_a, b_a = (mod.embedding.div_term, mod.proj_d.weight.detach(), mod.proj_d.bias.detach(),
                                 mod.proj_a.weight.detach(), mod.proj_a.bias.detach())
    fa = 180.0 / (15 * math.pi)
    torch.backends.cuda.matmul.allow_tf32 = False
    t_f = timeit(lambda: geo.geometric_embedding(pts, dterm, w_d, b_d, w_a, b_a, 0.2, fa, 3))
    t_i = timeit(lambda: geo.geometric_embedding_indices(pts, 0.2, fa, 3))

    def torch_sequence():     # the reference's op sequence = the module's non-fused branch (taken when autograd is on)
        with torch.enable_grad():
            return mod(pts).detach()
    t_r = timeit(torch_sequence, it=3, warm=1)
    # algorithmic work: per pair 1 distance row + k angle rows, each a C x C projection
    k = 3
    flops = 2.0 * B * N * N * (1 + k) * C * C
    pairs = B * N * N
    tiles = -(-pairs // 128) + -(-pairs // (4 * (32 // k)))          # 128-row MMA tiles actually issued
    issued = 3.0 * tiles * 128 * 2 * C * C                           # 3xTF32
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    try:
        peaks = json.load(open(os.path.join(root, "MEASURED_PEAKS.json")))
        peak, src = float(peaks["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json, sustained dense bf16)"
    except Exception:
        peak, src = 1405.0, "fallback"
    ach = flops / t_f * 1e-9
    print(json.dumps({
        "metric": "geometric_embeddings_per_s", "value": B / t_f * 1e3, "unit": "clouds/s",
        "config": {"workload": "GeometricStructureEmbedding forward, N=%d (196 + background point), hidden %d, angle_k %d, max"
                               % (N, C, k), "clouds_per_call": B},
        "fused_ms": t_f, "indices_ms": t_i, "reference_torch_gpu_ms": t_r, "speedup_vs_reference_torch_gpu": t_r / t_f,
        "roofline": {"kernel": "k_geo_embed<0> + k_geo_embed<1> (tcgen05 3xTF32, sinusoid operand generated in smem) + "
                               "k_geo_indices + 2 x k_split_tf32", "bound": "tensor", "achieved": ach, "peak": peak,
                     "unit": "TFLOP/s", "frac": ach / peak, "peak_source": src, "issued_tf32_tflops": issued / t_f * 1e-9,
                     "tf32_issue_peak": peak / 2, "frac_of_tf32_issue_peak": issued / t_f * 1e-9 / (peak / 2),
                     "traffic": None,
                     "note": "achieved = ALGORITHMIC flops 2 (1+k) C^2 per pair / whole-call time; 3 MMAs per product"},
        "out_bytes": B * N * N * C * 4}))


if __name__ == "__main__":
    main()
