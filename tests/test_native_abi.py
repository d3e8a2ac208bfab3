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


def test_list_marshalling_helper_matches_numpy():
    """csrc/qb_pyhelper.c (CPython C API through ctypes.PyDLL) packs the reference's list[list[float]] exactly like the NumPy
    path: values, offsets, length errors, ints among the floats; rows it cannot read fall back to NumPy."""
    import numpy as np

    from queasars_b200 import _build, _native
    from queasars_b200.engine import Engine

    _build.build_pyhelper()
    assert _native.pyhelper(), "helper library not loadable"

    class Plan:
        def __init__(self, n):
            self.n_params, self.plan_id = n, 7

    rng = np.random.default_rng(1)
    plans = [Plan(n) for n in (5, 0, 3, 11)]
    rows = [[float(v) for v in rng.normal(size=p.n_params)] for p in plans]
    rows[0][2] = 3  # an int among the floats
    ids, flat, offsets = Engine._pack(plans, rows)
    assert list(offsets) == [0, 5, 5, 8, 19] and list(ids) == [7] * 4
    assert np.array_equal(flat[:19], np.concatenate([np.asarray(r, dtype=np.float64) for r in rows]))
    tuples = tuple(tuple(r) for r in rows)
    assert np.array_equal(Engine._pack(plans, tuples)[1][:19], flat[:19])
    import pytest

    with pytest.raises(ValueError, match="circuit 2 has 3 parameters but 2 values"):
        Engine._pack(plans, rows[:2] + [rows[2][:2]] + rows[3:])
    with pytest.raises((ValueError, TypeError)):
        Engine._pack(plans, rows[:3] + [["x"] * 11])
