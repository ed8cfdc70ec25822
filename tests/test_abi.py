"""The C-ABI library loads here (no GPU) and exports every symbol include/roadsurf_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "roadsurf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rs_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_list_agree():
    from proj_roadsurf_b200 import _native
    assert header_symbols() == sorted(_native.EXPORTS)


def test_library_builds_loads_and_exports_every_symbol():
    from proj_roadsurf_b200 import _build, _native
    path = _build.build_native()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    for name in header_symbols():
        assert hasattr(lib, name), name
    L = _native.load()
    assert L.rs_version() == 100
    assert L.rs_status_string(0) == b"ok"
    assert L.rs_status_string(-3) == b"scanline crossing capacity exceeded"


def test_no_cpu_fallback_without_device():
    """without a CUDA device the product path fails loudly (RS_ERR_NO_DEVICE), it never computes on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from proj_roadsurf_b200._native import NativeError
    from proj_roadsurf_b200.engine import Engine
    with pytest.raises(NativeError) as ei:
        Engine(0)
    assert ei.value.status == -5


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "proj_roadsurf_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h")):
                s = open(os.path.join(d, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", s, flags=re.M), os.path.join(d, f)
                assert "libroadsurf_oracle" not in s
