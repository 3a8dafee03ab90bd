"""The C-ABI library builds, loads and exports every symbol include/deflow_b200.h declares (CPU only,
no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from deflow_b200 import _lib
    return _lib.build_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "deflow_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    names = declared_symbols()
    assert "dfb_pillar_index" in names and "dfb_pfn_forward" in names and len(names) >= 15


def test_library_exports_every_declared_symbol(lib_path):
    handle = ctypes.CDLL(lib_path)
    missing = [n for n in declared_symbols() if not hasattr(handle, n)]
    assert not missing, missing


def test_python_binding_covers_header(lib_path):
    from deflow_b200 import _lib
    assert sorted(_lib.EXPORTS) == declared_symbols()
    lib = _lib.lib()
    assert lib.dfb_version() >= 100
    assert lib.dfb_launch_count() >= 0


def test_host_only_entry_points(lib_path):
    """dfb_grid_size / dfb_index_workspace are pure host arithmetic and must match the oracle."""
    from deflow_b200 import ops
    from oracle import mmcv_ext_oracle as ext
    for vs, rg in (([0.2, 0.2, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]), ([0.1, 0.1, 6], [-51.2, -51.2, -3, 51.2, 51.2, 3]),
                   ([0.2, 0.2, 6], [-6.4, -6.4, -3, 6.4, 6.4, 3]), ([0.3, 0.25, 0.2], [-10, -7, -3, 11.1, 8.3, 3.4])):
        assert ops.grid_size(vs, rg) == ext.grid_size(vs, rg)


def test_no_cpu_fallback():
    import torch
    import deflow_b200 as d
    m = d.DeFlow(grid_feature_size=[64, 64], point_cloud_range=[-6.4, -6.4, -3, 6.4, 6.4, 3])
    batch = {"pc0": torch.zeros(1, 4, 3), "pc1": torch.zeros(1, 4, 3), "pose0": [torch.eye(4)], "pose1": [torch.eye(4)]}
    with pytest.raises(RuntimeError):
        m(batch)
    with pytest.raises(RuntimeError):
        d.DynamicScatter([0.2, 0.2, 6], [-6.4, -6.4, -3, 6.4, 6.4, 3], True)(torch.zeros(4, 3), torch.zeros(4, 3, dtype=torch.int32))
