"""Load the REFERENCE's own Python modules (OpenSceneFlow/src/models/*, assets/cuda/mmcv/*.py, loss / metric functions)
UNMODIFIED, from /root/reference in the build container or from the staged copy oracle/_ref/osf on the GPU box
(oracle/build_ref.py stage(); git-ignored, never part of the repo's history).

TEST INFRASTRUCTURE ONLY: used by tests/golden/make_golden.py (fixtures), the GPU tests that run the reference's own
scatter_points.py / voxelize.py on top of deflow_b200's mmcv._ext drop-in, bench.py's CPU reference arm and
tools/ref_gpu_bench.py (the GPU bar-to-beat).  deflow_b200 never imports this.

Two stubs make the import possible: ``dztimer`` (wall-clock instrumentation, no arithmetic) and the
plugin module ``mmcv._ext`` the reference locates by name (scatter_points.py:11-19, voxelize.py:11-18), which is one of
  "numpy" -- oracle/mmcv_ext_oracle.py (CPU restatement, pinned against the CUDA kernels: tests/golden/ext_gpu_ref.npz),
  "cuda"  -- the reference's own CUDA extension compiled by oracle/build_ref.py (oracle/_ref/mmcv_ref_ext.so),
  "dfb"   -- deflow_b200.mmcv_ext (the product's drop-in; what INTEGRATION.md tells a maintainer to install).
Loss / metric functions are AST-extracted because their modules import chamfer3D / av2 / h5py at import time."""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# the staged copy first: what runs on the GPU box (where /root/reference does not exist) is then also what runs here
CANDIDATES = [os.path.join(HERE, "_ref", "osf"), "/root/reference/OpenSceneFlow"]


def root():
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "src", "models")):
            return c
    return None


class _Timing:
    def __getitem__(self, i):
        return self

    def start(self, *a, **k):
        return None

    def stop(self, *a, **k):
        return None

    def print(self, *a, **k):
        return None


def _numpy_ext():
    from . import mmcv_ext_oracle as ext_np
    e = types.ModuleType("mmcv._ext")

    def dynamic_voxelize_forward(points, voxel_size, coors_range, coors, NDim=3):
        out = ext_np.dynamic_voxelize_forward(points.detach().numpy(), voxel_size.numpy(), coors_range.numpy(), coors.numpy())
        coors.copy_(torch.from_numpy(out))

    def dynamic_point_to_voxel_forward(feats, coors, reduce_type):
        r = ext_np.dynamic_point_to_voxel_forward(feats.detach().numpy(), coors.numpy(), reduce_type)
        return [torch.from_numpy(np.ascontiguousarray(x)) for x in r]

    def dynamic_point_to_voxel_backward(grad_feats, grad_reduced_feats, feats, reduced_feats, coors_idx, reduce_count,
                                        reduce_type):
        g = ext_np.dynamic_point_to_voxel_backward(grad_reduced_feats.numpy(), feats.detach().numpy(),
                                                   reduced_feats.detach().numpy(), coors_idx.numpy(), reduce_count.numpy(),
                                                   reduce_type)
        grad_feats.copy_(torch.from_numpy(g))

    def hard_voxelize_forward(*a, **k):
        raise RuntimeError("hard_voxelize_forward is not on the DeFlow path")

    e.dynamic_voxelize_forward = dynamic_voxelize_forward
    e.dynamic_point_to_voxel_forward = dynamic_point_to_voxel_forward
    e.dynamic_point_to_voxel_backward = dynamic_point_to_voxel_backward
    e.hard_voxelize_forward = hard_voxelize_forward
    return e


def install_stubs(ext="numpy"):
    dz = types.ModuleType("dztimer")
    dz.Timing = _Timing
    sys.modules["dztimer"] = dz
    if ext == "numpy":
        e = _numpy_ext()
    elif ext == "cuda":
        from . import build_ref
        e = build_ref.load_ref()
        if e is None:
            raise RuntimeError("oracle/_ref/mmcv_ref_ext.so is missing (python oracle/build_ref.py in the build container)")
    elif ext == "dfb":
        from deflow_b200 import mmcv_ext
        e = mmcv_ext.install_as_mmcv_ext()
    else:
        raise ValueError(ext)
    pkg = sys.modules.get("mmcv")
    if pkg is None or getattr(pkg, "_ext", None) is not e:
        pkg = types.ModuleType("mmcv")
        pkg._ext = e
        sys.modules["mmcv"] = pkg
        sys.modules["mmcv._ext"] = e
    return e


def _purge():
    for k in list(sys.modules):
        if k == "src" or k.startswith("src.") or k == "assets" or k.startswith("assets."):
            del sys.modules[k]


def extract_functions(rel_path, names, extra_ns=None):
    """exec the named top-level function / class definitions of a reference file with only torch / numpy in scope."""
    r = root()
    tree = ast.parse(open(os.path.join(r, rel_path)).read())
    ns = {"torch": torch, "np": np}
    ns.update(extra_ns or {})
    for node in tree.body:
        if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name in names:
            exec(compile(ast.Module([node], []), rel_path, "exec"), ns)
    return ns


def load_reference(ext="numpy"):
    """-> (DeFlow, FastFlow3D, namespace with deflowLoss / ff3dLoss / zeroflowLoss).  Re-importable with another ext."""
    r = root()
    if r is None:
        raise RuntimeError("reference modules not found: neither /root/reference nor oracle/_ref/osf (oracle/build_ref.py)")
    install_stubs(ext)
    _purge()   # scatter_points.py / voxelize.py bind mmcv._ext at import time
    if r not in sys.path:
        sys.path.insert(0, r)
    from src.models.deflow import DeFlow  # noqa
    from src.models.fastflow3d import FastFlow3D  # noqa
    lossns = extract_functions("src/lossfuncs.py", ("deflowLoss", "ff3dLoss", "zeroflowLoss"))
    return DeFlow, FastFlow3D, lossns


def load_mmcv_wrappers(ext="dfb"):
    """The reference's own assets/cuda/mmcv/{scatter_points,voxelize}.py bound to the chosen mmcv._ext."""
    r = root()
    if r is None:
        raise RuntimeError("reference modules not found")
    install_stubs(ext)
    _purge()
    if r not in sys.path:
        sys.path.insert(0, r)
    import assets.cuda.mmcv as m  # noqa
    return m


def load_weights_init():
    import torch.nn as nn
    import torch.nn.init as init
    return extract_functions("src/utils/mics.py", ("weights_init",), {"nn": nn, "init": init})["weights_init"]


def load_eval_metric():
    """The reference's src/utils/eval_metric.py + av2_eval.py, imported unmodified.  ``av2`` (av2==0.2.1, OSF/environment.yaml:35)
    is not installed: stub modules carry the one thing the metric code takes from it, the AnnotationCategories enum (order
    restated in oracle/eval_oracle.py from the published av2-api source), plus placeholders for names it only imports."""
    import enum
    from . import eval_oracle
    r = root()
    if r is None:
        raise RuntimeError("reference modules not found")
    cats = enum.Enum("AnnotationCategories", {k: k for k in eval_oracle.ANNOTATION_CATEGORIES}, type=str)
    mods = {"av2": {}, "av2.datasets": {}, "av2.datasets.sensor": {}, "av2.datasets.sensor.constants": {"AnnotationCategories": cats},
            "av2.geometry": {}, "av2.geometry.geometry": {}, "av2.geometry.se3": {"SE3": object},
            "av2.utils": {}, "av2.utils.typing": {"NDArrayFloat": np.ndarray, "NDArrayBool": np.ndarray, "NDArrayInt": np.ndarray},
            "av2.utils.io": {"read_feather": None}}
    for name, attrs in mods.items():
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
    for name in mods:                      # parent.child attributes (import av2.geometry.geometry as ...)
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[name])
    if not hasattr(np, "NaN"):             # the reference targets numpy < 2 (eval_metric.py:133 uses np.NaN)
        np.NaN = np.nan
    _purge()
    if r not in sys.path:
        sys.path.insert(0, r)
    import src.utils.eval_metric as em  # noqa
    return em
