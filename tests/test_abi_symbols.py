"""The C-ABI library loads without a GPU and exports exactly what include/*.h declares."""
import ctypes
import os
import re

import pytest

import nsb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if not fn.endswith(".h"):
            continue
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b(ns_[a-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    return names


def test_every_declared_symbol_is_exported_and_bound():
    lib = nsb200._lib.load()
    decl = declared_functions()
    assert len(decl) >= 35
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported by libnsb200.so"
    assert decl == set(nsb200._lib.SYMBOLS), (decl ^ set(nsb200._lib.SYMBOLS))


def test_header_cites_the_reference_interface():
    text = open(os.path.join(ROOT, "include", "nextsearch_b200.h")).read()
    for cite in ("src/api_engine.cpp:369-542", "src/api_engine.cpp:50", "include/api_types.hpp", "include/textutil.hpp:13-37"):
        assert cite in text


def test_library_has_no_torch_or_python_dependency():
    out = os.popen(f"ldd {nsb200._lib.LIB_PATH}").read()
    assert "torch" not in out and "python" not in out


def test_no_cpu_fallback_without_a_gpu():
    """Product path fails loudly when no CUDA device is usable."""
    lib = nsb200._lib.load()
    if lib.ns_device_count() > 0:
        pytest.skip("a GPU is present")
    h = ctypes.c_void_p()
    rc = lib.ns_index_create(0, ctypes.byref(h))
    assert rc == 2 and b"no CPU fallback" in lib.ns_last_error()


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "nextsearch-api_b200")
    for base, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cpp", ".cu", ".cuh", ".hpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(base, fn), errors="replace").read()
                assert "oracle" not in text.lower() or fn in ("corpus.hpp", "corpus.cpp", "engine.cpp") and "oracle/ref_driver" in text, \
                    f"{fn} mentions the oracle"
