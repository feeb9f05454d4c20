"""Recipe: compile the reference's UNMODIFIED pointnet2 `_ext` (C++/CUDA) for sm_100a.

TEST INFRASTRUCTURE ONLY.  The sources are compiled *where they lie* under
/root/reference (core/unopose/model/pointnet2/_ext_src/{src,include}); nothing
is copied into this repository.  Outputs go to oracle/_ref/ (git-ignored, but
shipped to the GPU box by gpurun) as `ref_pointnet2_ext.so`, a torch/pybind
module exposing the 9 functions of _ext_src/src/bindings.cpp:11-24.

The build mirrors the reference's setup.py:22-38 flags (-O3, the CUDA_NO_HALF
defines, default -fmad=true) and adds only the sm_100a -gencode the reference
leaves to TORCH_CUDA_ARCH_LIST.

Usage:  python oracle/build_ref_ext.py            (needs /root/reference)
The GPU box has no /root/reference: there only the prebuilt .so is loaded
(see oracle/ref_ext.py).
"""
import glob
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF_SRC = "/root/reference/core/unopose/model/pointnet2/_ext_src"
NAME = "ref_pointnet2_ext"


def build(verbose=False):
    so = os.path.join(OUT, NAME + ".so")
    if os.path.exists(so):
        return so
    if not os.path.isdir(REF_SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load

    sources = sorted(glob.glob(REF_SRC + "/src/*.cpp") + glob.glob(REF_SRC + "/src/*.cu"))
    load(
        name=NAME,
        sources=sources,
        extra_include_paths=[REF_SRC + "/include"],
        extra_cuda_cflags=[
            "-O3",
            "-DCUDA_HAS_FP16=1",
            "-D__CUDA_NO_HALF_OPERATORS__",
            "-D__CUDA_NO_HALF_CONVERSIONS__",
            "-D__CUDA_NO_HALF2_OPERATORS__",
            "-gencode", "arch=compute_100a,code=sm_100a",
        ],
        build_directory=OUT,
        is_python_module=False,
        verbose=verbose,
    )
    return so if os.path.exists(so) else None


if __name__ == "__main__":
    p = build(verbose=True)
    print("built:", p)
    sys.exit(0 if p else 1)
