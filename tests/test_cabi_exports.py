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


def test_ctypes_structures_match_the_header_layout(tmp_path):
    """Every struct of include/deflow_b200.h against its ctypes mirror in deflow_b200/_lib.py: total size and the offset of
    every field, taken from the header by a C compiler (a field added on one side only would shift everything after it)."""
    import shutil
    import subprocess
    from deflow_b200 import _lib
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    pairs = {"dfb_index_args": _lib.IndexArgs, "dfb_conv_args": _lib.ConvArgs, "dfb_pack_desc": _lib.PackDesc, "dfb_unpack_desc": _lib.UnpackDesc,
             "dfb_pfn_args": _lib.PfnArgs, "dfb_pfn_bwd_args": _lib.PfnBwdArgs, "dfb_eval_tables": _lib.EvalTables}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "deflow_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, what, val = line.split()
        cls = pairs[cname]
        if what == "size":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, what).offset == int(val), (cname, what, getattr(cls, what).offset, val)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
