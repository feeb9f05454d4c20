#!/bin/bash
# Round-2 opener (one gpurun call, ~2.5 min of box time): validates the opt-in 3xFP16 similarity path that round 1 could
# only compile and measures it against the default 3xTF32 path.
#   gpurun --timeout 300 -- 'bash scripts/r2_first_gpu_call.sh'
# Outputs under gpurun_out/: r2_fp16_test.log, bench_tf32.json, bench_fp16.json, geo_tf32.json, geo_fp16.json
set -u
mkdir -p gpurun_out
UPK_TEST_EXPERIMENTAL=1 timeout 120 python -m pytest tests/test_pose_gpu.py -x -q -k experimental_fp16 > gpurun_out/r2_fp16_test.log 2>&1
tail -3 gpurun_out/r2_fp16_test.log
# the whole fine-stage parity suite with the experimental mode forced on (statistics + solve consume its logits)
UPK_SIMILARITY_MODE=16 timeout 200 python -m pytest tests/test_pose_gpu.py tests/test_pipeline_gpu.py -x -q 2>&1 | tail -3
timeout 100 python bench.py --no-cpu-baseline --no-gpu-torch-baseline --no-widened > gpurun_out/bench_tf32.json 2>/dev/null
UPK_SIMILARITY_MODE=16 timeout 100 python bench.py --no-cpu-baseline --no-gpu-torch-baseline --no-widened > gpurun_out/bench_fp16.json 2>/dev/null
# the same for the geometric embedding (UPK_GEO_F16=1 is read once per process): parity suite, then timing in both modes
UPK_GEO_F16=1 timeout 120 python -m pytest tests/test_geo_gpu.py -x -q 2>&1 | tail -3
timeout 60 python scripts/geo_bench.py > gpurun_out/geo_tf32.json 2>/dev/null
UPK_GEO_F16=1 timeout 60 python scripts/geo_bench.py > gpurun_out/geo_fp16.json 2>/dev/null
python - <<'PY'
import json
for name in ("tf32", "fp16"):
    try:
        d = json.load(open("gpurun_out/geo_%s.json" % name))
        print("geo", name, "ms", round(d["fused_ms"], 4), "torch ms", round(d["reference_torch_gpu_ms"], 3))
    except Exception as e:  # noqa: BLE001
        print("geo", name, "failed:", e)
for name in ("tf32", "fp16"):
    try:
        d = json.load(open("gpurun_out/bench_%s.json" % name))
        print(name, "instances/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "fine_similarity ms",
              round(d["stage_ms"]["fine_similarity"], 4))
    except Exception as e:  # noqa: BLE001
        print(name, "failed:", e)
PY
