"""The C-ABI library loads without a GPU and exports every symbol include/hyrise_b200.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from hyrise_b200 import capi

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols(header: str) -> set[str]:
    with open(os.path.join(REPO, "include", header)) as handle:
        text = re.sub(r"/\*.*?\*/", "", handle.read(), flags=re.S)
    return set(re.findall(r"\b(hyb_[a-z0-9_]+)\s*\(", text))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    declared = declared_symbols("hyrise_b200.h")
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in include/hyrise_b200.h but not exported"
    assert declared == set(capi.SYMBOLS), "capi.SYMBOLS and the header disagree"
    assert lib.hyb_abi_version() == 1


def test_generator_library_exports_its_header():
    from hyrise_b200 import tpch

    lib = tpch.load()
    for name in sorted(declared_symbols("hyrise_b200_tpch.h")):
        assert hasattr(lib, name), name


def test_struct_layouts_match_the_header():
    assert C.sizeof(capi.RowID) == 8
    assert C.sizeof(capi.SegmentDesc) == 56
    assert C.sizeof(capi.TableView) == 16
    assert C.sizeof(capi.ScanPredicate) == 32
    assert C.sizeof(capi.ExprNode) == 24
    assert C.sizeof(capi.AggregateDef) == 8 + 24 * capi.MAX_EXPR_NODES
    assert C.sizeof(capi.OperatorStats) == 40


def test_no_gpu_means_a_loud_failure():
    """Without a CUDA device the product refuses to run instead of falling back to the CPU."""
    lib = capi.load_library()
    count = C.c_int()
    status = lib.hyb_device_count(C.byref(count))
    if status == capi.HYB_OK and count.value > 0:
        pytest.skip("a GPU is present")
    context = C.c_void_p()
    assert lib.hyb_context_create(0, C.byref(context)) != capi.HYB_OK
    assert lib.hyb_last_error()
    from hyrise_b200.device import DeviceContext
    with pytest.raises(capi.HyriseB200Error):
        DeviceContext(0)


def test_product_never_imports_the_oracle():
    for root, _, files in os.walk(os.path.join(REPO, "hyrise_b200")):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp")):
                with open(os.path.join(root, name)) as handle:
                    text = handle.read()
                assert "oracle_lib" not in text and "liboracle" not in text and "oracle/" not in text, name
