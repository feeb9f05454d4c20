"""Deterministic, key-addressed module weights for tests and golden generators: every tensor of a state dict is drawn
from a numpy PCG64 stream seeded by (seed, crc32(key)), so the reference module (golden generation, this container)
and the product module (test time, any box) receive bit-identical weights without shipping them."""
import zlib

import numpy as np
import torch


def keyed_state_dict(template, seed):
    out = {}
    for k in sorted(template):
        v = template[k]
        if not v.is_floating_point():
            out[k] = v.clone()
            continue
        rng = np.random.default_rng([seed, zlib.crc32(k.encode())])
        shape = tuple(v.shape)
        n = rng.standard_normal(shape if shape else (1,)).astype(np.float32).reshape(shape)
        if "running_var" in k:
            a = 0.5 + np.abs(n)
        elif "running_mean" in k:
            a = 0.1 * n
        elif k.endswith(".gamma"):
            a = 1e-5 * (1.0 + 0.1 * n)             # LayerScale at its own initial value (DINOv2: init_values = 1e-5)
        elif k.endswith("bias"):
            a = 0.02 * n
        elif v.dim() <= 1 or "norm" in k.lower() or ".bn." in k:
            a = 1.0 + 0.1 * n                      # LayerNorm / BatchNorm scales
        else:
            fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
            a = n / np.sqrt(max(fan_in, 1))
        out[k] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    return out
