"""The C-ABI library loads and exports exactly what include/torchcde_b200.h declares (no GPU)."""
import os
import re

import pytest

from conftest import ROOT
from torchcde_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "torchcde_b200.h")).read()
    return sorted(set(re.findall(r"TCDE_API\s+[\w\s\*]+?\b(tcde_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    names = _declared()
    assert len(names) >= 14
    assert sorted(_lib.SIGNATURES) == names


def test_library_exports_every_symbol():
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.tcde_abi_version() == 2


def test_argument_errors_do_not_need_a_gpu():
    lib = _lib.load()
    # null pointers are rejected before any CUDA call
    assert lib.tcde_hermite_bdiff_coeffs(None, None, None, 1, 4, 2, 0, None, None) == -1
    assert b"null" in lib.tcde_last_error()
    with pytest.raises(ValueError):
        _lib.call("tcde_linear_fill", None, None, None, 1, 4, 2, 0, None, None)


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torchcde_b200")
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            src = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(from|import)\s+\.*oracle", src, flags=re.M), name + " imports oracle/"
            assert "import_module(\"oracle" not in src and "__import__(\"oracle" not in src
