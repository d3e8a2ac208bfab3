"""CPU-side checks of the C-ABI library: it builds, loads, exports every symbol include/queasars_b200.h
declares, agrees on record layouts, and fails loudly (no CPU fallback) when no CUDA device is present."""
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    text = open(os.path.join(ROOT, "include", "queasars_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[a-z_0-9]+)\s*\(", text)))


def test_build_and_symbols():
    import __graft_entry__ as entry

    entry.build()
    from queasars_b200 import _native

    lib = _native.load()
    declared = header_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert sorted(_native.EXPORTED_SYMBOLS) == declared


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from queasars_b200 import _native
    from queasars_b200.engine import Engine

    with pytest.raises(_native.QbError):
        Engine(device=0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "queasars_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
