"""Stage the UNMODIFIED reference files of the hot path under baseline/_ref/ (git-ignored, shipped to the GPU box).

    python baseline/stage_reference.py          (needs /root/reference; a no-op elsewhere)

The reference (shanice-l/UNOPose) is not pip-installable (no setup.py / pyproject at its root), so "installing" it
means placing its own Python files where `import core.unopose...` finds them.  Only the files the correspondence-and-
pose path imports are staged, byte for byte (sha256 in MANIFEST.json); its compiled extension is built unmodified by
oracle/build_ref_ext.py.  Nothing staged here is tracked by git or imported by the product (unopose_b200/).
Consumers: baseline/refgpu.py -> bench.py (`gpu_reference` leg, `--impl reference`), tests/test_dropin_gpu.py.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
SRC = "/root/reference"
FILES = [
    "core/unopose/__init__.py",
    "core/unopose/utils/model_utils.py",
    "core/unopose/utils/loss_utils.py",
    "core/unopose/model/transformer.py",
    "core/unopose/model/oneref_predator_coarse_point_matching.py",
    "core/unopose/model/oneref_predator_fine_point_matching.py",
    "core/unopose/model/oneref_grf_predator_pose_estimation_model.py",
    "core/unopose/model/oneref_feature_extraction.py",
    "core/unopose/model/pointnet2/pointnet2_utils.py",
    "core/unopose/model/pointnet2/pytorch_utils.py",
    "core/unopose/engine/oneref_inference_utils_v1.py",
    "LICENSE",
]


def stage():
    if not os.path.isdir(SRC):
        return None
    manifest = {}
    for rel in FILES:
        src = os.path.join(SRC, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    json.dump({"source": SRC, "files": manifest}, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
    return DST


if __name__ == "__main__":
    print("staged:", stage())
