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


# Python modules of the reference that the on-GPU checker / bar-to-beat need (oracle/ref_modules.py): staged verbatim into
# the git-ignored oracle/_ref/osf so that they travel to the GPU box next to mmcv_ref_ext.so.
PY_ROOT = "/root/reference/OpenSceneFlow"
PY_FILES = ["src/__init__.py", "src/models/__init__.py", "src/models/deflow.py", "src/models/fastflow3d.py",
            "src/models/flow4d.py", "src/models/ssf.py",
            "src/models/basic/__init__.py", "src/models/basic/encoder.py", "src/models/basic/decoder.py",
            "src/models/basic/unet.py", "src/models/basic/flow4d_module.py", "src/models/basic/ssf_module.py",
            "src/lossfuncs.py", "src/utils/mics.py", "src/utils/eval_metric.py", "src/utils/av2_eval.py", "src/trainer.py",
            "src/dataset.py",
            "assets/__init__.py", "assets/cuda/__init__.py", "assets/cuda/mmcv/__init__.py",
            "assets/cuda/mmcv/scatter_points.py", "assets/cuda/mmcv/voxelize.py",
            "assets/cuda/chamfer3D/__init__.py", "assets/cuda/chamfer3D/chamfer3D_cuda.cpp",
            "assets/cuda/chamfer3D/chamfer3D.cu", "assets/cuda/chamfer3D/setup.py",
            "assets/tests/test_pc0.npy", "assets/tests/test_pc1.npy"]


def stage():
    """Copy PY_FILES into oracle/_ref/osf (outputs only under oracle/_ref; nothing enters the repo's history)."""
    import shutil
    if not os.path.isdir(PY_ROOT):
        return None
    dst_root = os.path.join(OUT, "osf")
    n = 0
    for rel in PY_FILES:
        src = os.path.join(PY_ROOT, rel)
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
            n += 1
        elif rel.endswith("__init__.py"):
            open(dst, "a").close()   # namespace marker the reference tree does not have
    return dst_root, n


def build_chamfer(verbose=False):
    """The reference's chamfer3D extension (assets/cuda/chamfer3D/{chamfer3D.cu, chamfer3D_cuda.cpp}; its setup.py lists
    exactly these two files) -> oracle/_ref/chamfer3D_ref.so, the on-GPU checker of csrc/chamfer.cu."""
    src = "/root/reference/OpenSceneFlow/assets/cuda/chamfer3D"
    if not os.path.isdir(src):
        return None
    out = os.path.join(OUT, "chamfer")
    os.makedirs(out, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    load(name="chamfer3D_ref", sources=[os.path.join(src, "chamfer3D_cuda.cpp"), os.path.join(src, "chamfer3D.cu")],
         build_directory=out, verbose=verbose, is_python_module=False)
    import shutil
    shutil.copyfile(os.path.join(out, "chamfer3D_ref.so"), os.path.join(OUT, "chamfer3D_ref.so"))
    return os.path.join(OUT, "chamfer3D_ref.so")


def load_chamfer_ref():
    global _loaded_chamfer
    if _loaded_chamfer is not None:
        return _loaded_chamfer
    path = os.path.join(OUT, "chamfer3D_ref.so")
    if not os.path.exists(path):
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location("chamfer3D_ref", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _loaded_chamfer = mod
    return mod


_loaded = None
_loaded_chamfer = None


def load_ref():
    """Import the prebuilt module (no compilation, no /root/reference access)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    path = os.path.join(OUT, "mmcv_ref_ext.so")
    if not os.path.exists(path):
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch symbols must be loaded first)
    spec = importlib.util.spec_from_file_location("mmcv_ref_ext", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _loaded = mod
    return mod


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
    print(build_chamfer(verbose="-v" in sys.argv))
    print(stage())
