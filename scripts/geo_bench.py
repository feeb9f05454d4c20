"""Dev tool: time the fused GeometricStructureEmbedding against the reference torch-op sequence on this GPU.
    python scripts/geo_bench.py [B] [N] [C]
"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import geo_oracle as G  # noqa: E402  (dev tool: the oracle is the timed baseline here)
from unopose_b200.modules import geo  # noqa: E402


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
    pts, dterm, w_d, b_d, w_a, b_a = G.make_inputs(1, B, N, C, dev)
    fa = 180.0 / (15 * math.pi)
    torch.backends.cuda.matmul.allow_tf32 = False
    t_f = timeit(lambda: geo.geometric_embedding(pts, dterm, w_d, b_d, w_a, b_a, 0.2, fa, 3))
    t_i = timeit(lambda: geo.geometric_embedding_indices(pts, 0.2, fa, 3))
    with torch.no_grad():
        t_r = timeit(lambda: G.geometric_embedding(pts, dterm, w_d, b_d, w_a, b_a, 0.2, 15, 3), it=3, warm=1)
    flops = 2.0 * B * N * N * 4 * C * C
    print(json.dumps({"B": B, "N": N, "C": C, "fused_ms": t_f, "indices_ms": t_i, "torch_ms": t_r,
                      "algorithmic_tflops": flops / t_f * 1e-9, "issued_tf32_tflops": 3 * flops / t_f * 1e-9 * (1275.0 / 1213.0),
                      "out_GBps": B * N * N * C * 4 * 3 / t_f * 1e-6}))


if __name__ == "__main__":
    main()
