"""Golden output of the REFERENCE inference driver (core/unopose/engine/oneref_inference_utils_v1.py) on the
deterministic fake model / loader of tests/inference_fixtures.py, run on CPU in this container.

    python tests/golden/make_inference_golden.py        (needs /root/reference)

Shims (none touches the driver's logic): `orjson` is absent -> a stub that forwards to json;
Tensor.cuda() -> identity and torch.cuda.synchronize() -> no-op (no GPU here); time.perf_counter() -> 0 so that
the `time` column is the deterministic seg_time.  Writes tests/golden/inference_v1.json."""
import json
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")
from inference_fixtures import FakePoseModel, make_loader  # noqa: E402


def main():
    stub = types.ModuleType("orjson")
    stub.dumps = lambda o: json.dumps(o).encode()
    sys.modules["orjson"] = stub
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.cuda.synchronize = lambda *a, **k: None
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "ref_inference", "/root/reference/core/unopose/engine/oneref_inference_utils_v1.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    ref.time.perf_counter = lambda: 0.0
    out = {}
    for name, tem in (("with_tem_pose", True), ("plain", False)):
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "result.csv")
            ref.inference_and_save_oneref_v1(FakePoseModel(), make_loader(0, 3, tem), path, instance_batch_size=16)
            out[name] = {"csv": open(path).read(), "json": json.loads(open(path.replace(".csv", ".json")).read())}
    json.dump(out, open(os.path.join(HERE, "inference_v1.json"), "w"))
    print("wrote inference_v1.json:", {k: len(v["csv"].splitlines()) for k, v in out.items()}, "lines")


if __name__ == "__main__":
    main()
