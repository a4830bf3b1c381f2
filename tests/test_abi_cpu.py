"""CPU: the C-ABI library loads, exports every symbol include/b200hash.h declares, and refuses to
work (loudly, no CPU fallback) when no GPU is visible."""
import os
import re

import pytest

from modal_client_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200hash.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200h_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    _lib.build_library()
    lib = _lib.load_library()
    for sym in _declared_symbols():
        assert hasattr(lib, sym), sym
    assert b"sm_100a" in lib.b200h_version()


def test_no_cpu_fallback_without_gpu():
    lib = _lib.load_library()
    if lib.b200h_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(_lib.B200HashError) as ei:
        _lib.Context(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "modal_client_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert not re.search(r"^\s*(from|import)\s+hashlib\b", src, flags=re.M), f
