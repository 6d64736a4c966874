"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/mchrg_b200.h declares; without a GPU every compute entry fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "mchrg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mchrg_b200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from multicharge_b200 import _lib
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/mchrg_b200.h but not exported"
    # and the ctypes prototypes cover exactly the header
    assert sorted(_lib.PROTOTYPES) == syms


def test_version_and_error_string():
    from multicharge_b200 import _lib
    lib = _lib.load()
    assert lib.mchrg_b200_version() >= 100
    assert isinstance(lib.mchrg_b200_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import multicharge_b200 as m
    with pytest.raises(m.MchrgError) as exc:
        m.Context(0)
    assert exc.value.code == -2 and "no CPU fallback" in str(exc.value)


def test_null_arguments_rejected():
    from multicharge_b200 import _lib
    lib = _lib.load()
    assert lib.mchrg_b200_context_create(0, None) == _lib.ERR_ARG
    assert lib.mchrg_b200_solve(None, None, *([None] * 12)) == _lib.ERR_ARG
    assert lib.mchrg_b200_synchronize(None) == _lib.ERR_ARG


def test_product_does_not_import_oracle():
    """The product path must never route through oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "multicharge_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU oracle", ""), f"{f} mentions the oracle"
