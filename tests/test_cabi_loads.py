"""CPU-side checks: the shared library exists, loads, and exports every symbol include/tdsa.h declares."""
import os
import re

import pytest

from topdogspectrumanalyser_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tdsa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tdsa_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = _lib.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), s
    assert sorted(_lib.SIGNATURES) == syms          # the ctypes table and the header agree
    assert lib.tdsa_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from topdogspectrumanalyser_b200.engine import SpectrumPlan
    with pytest.raises(_lib.TdsaError):
        SpectrumPlan(1024)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "topdogspectrumanalyser_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)
