"""CPU: the C-ABI library loads and exports every symbol include/bpp_b200.h declares.
No compute call is made (there is no GPU here and no CPU fallback to call)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from helpers import ROOT
from bpp_b200 import _lib, build


@pytest.fixture(scope="module")
def lib_path():
    return build.build_cuda()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "bpp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bppgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_loader_declare_the_same_symbols():
    assert header_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib_path):
    L = C.CDLL(lib_path)
    for name in header_symbols():
        assert hasattr(L, name), name


def test_library_is_sm100a_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_partial_op_layout_matches_header():
    assert C.sizeof(_lib.PartialOp) == 32
    assert [f[0] for f in _lib.PartialOp._fields_] == [
        "parent_clv_index", "left_clv_index", "right_clv_index", "left_pmatrix_index",
        "right_pmatrix_index", "parent_scaler_index", "left_scaler_index", "right_scaler_index"]


def test_no_cpu_fallback_without_device(lib_path):
    """Without a GPU the engine must refuse loudly instead of computing on the CPU."""
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("a GPU is present")
    from bpp_b200 import engine
    with pytest.raises(engine.BppGpuError):
        engine.Engine(0)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bpp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"\boracle\b", src), f
