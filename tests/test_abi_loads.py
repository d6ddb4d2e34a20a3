"""CPU test: the C-ABI library loads without a GPU and exports every symbol that
include/mtt_b200.h declares; creating a context without a device fails loudly."""
import os
import re

import pytest

import multitypetree_b200 as mtt
from multitypetree_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "mtt_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mtt_[a-z0-9_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported_and_bound():
    L = mtt.lib()
    names = _declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L, name), "libmtt_b200.so does not export %s" % name
        assert name in _lib.ABI, "%s is declared in the header but not bound in _lib.ABI" % name
    assert set(_lib.ABI) == set(names)
    assert L.mtt_abi_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(mtt.MttError) as e:
        mtt.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multitypetree_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), \
                    "%s refers to the oracle" % os.path.join(dirpath, f)
