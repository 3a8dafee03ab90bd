"""Compile the REFERENCE's own mmcv CUDA extension (OpenSceneFlow/assets/cuda/mmcv/*.cu, *.cpp) into oracle/_ref/
for sm_100, from the sources where they lie under /root/reference (nothing is copied into the repo).

TEST INFRASTRUCTURE ONLY.  The resulting module (oracle/_ref/mmcv_ref_ext.so, git-ignored, travels to the GPU box)
is the reference's dynamic_voxelize / dynamic_point_to_voxel kernels, used by tests/test_gpu_vs_reference_ext.py to pin
integer parity of the CUDA path against the reference kernels themselves on a B200.  The recipe is a direct
torch.utils.cpp_extension.load of the six translation units the reference's setup.py lists (setup.py:13-18); the
reference's own build system is not run."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/OpenSceneFlow/assets/cuda/mmcv"
OUT = os.path.join(HERE, "_ref")
FILES = ["scatter_points_cuda.cu", "scatter_points.cpp", "voxelization_cuda.cu", "voxelization.cpp", "cudabind.cpp",
         "pybind.cpp"]


def build(verbose=False):
    if not os.path.isdir(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    load(name="mmcv_ref_ext", sources=[os.path.join(SRC, f) for f in FILES], build_directory=OUT, verbose=verbose,
         extra_cuda_cflags=["-O3"], is_python_module=False)
    return os.path.join(OUT, "mmcv_ref_ext.so")


def load_ref():
    """Import the prebuilt module (no compilation, no /root/reference access)."""
    path = os.path.join(OUT, "mmcv_ref_ext.so")
    if not os.path.exists(path):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location("mmcv_ref_ext", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
