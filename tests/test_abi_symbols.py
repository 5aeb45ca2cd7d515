"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/secedo_b200.h declares; the Python binding covers exactly that set; and the product refuses
to run (loudly) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from secedo_b200 import _lib


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "secedo_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libsecedo_b200.so does not export {s}"


def test_binding_covers_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from secedo_b200.api import Context, SgpuError
    with pytest.raises(SgpuError):
        Context(0)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "secedo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "secedo_oracle" not in src, f
