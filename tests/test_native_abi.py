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


def test_single_evaluation_helper_marshals_and_passes_status_through():
    """qb_single_expectation (csrc/qb_pyhelper.c): the one-circuit call of the optimizer loop.  Checked against a stand-in for
    qb_evaluate_expectation with the same C signature (no GPU here): arguments as the C-ABI defines them, value returned as a
    Python float, native status / length error / unreadable row reported as ints."""
    import ctypes

    from queasars_b200 import _build, _native

    _build.build_pyhelper()
    helper = _native.pyhelper()
    assert helper, "helper library not loadable"
    seen = {}

    @ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_double),
                      ctypes.POINTER(ctypes.c_longlong), ctypes.c_longlong, ctypes.POINTER(ctypes.c_double))
    def fake_eval(ctx, batch, ids, params, offsets, ham_id, out):
        n = offsets[1] - offsets[0]
        seen.update(ctx=ctx, batch=batch, plan=ids[0], first=offsets[0], n=n, ham=ham_id)
        if ham_id == 99:
            return _native.QB_ERR_NOT_FOUND
        out[0] = sum(params[i] * (i + 1) for i in range(n)) + 0.5
        return 0

    fn = ctypes.cast(fake_eval, ctypes.c_void_p)
    ctx = ctypes.c_void_p(0x1234)
    row = [0.25, 2, -1.5]  # an int among the floats
    r = helper.qb_single_expectation(fn, ctx, 42, 3, row, 7)
    assert type(r) is float and r == 0.25 + 2 * 2 - 1.5 * 3 + 0.5
    assert seen == {"ctx": 0x1234, "batch": 1, "plan": 42, "first": 0, "n": 3, "ham": 7}
    assert helper.qb_single_expectation(fn, ctx, 42, 3, tuple(row), 7) == r
    assert helper.qb_single_expectation(fn, ctx, 42, 0, [], 7) == 0.5  # parameter-free circuit
    long_row = [float(i) for i in range(700)]  # beyond the helper's stack buffer
    assert helper.qb_single_expectation(fn, ctx, 1, 700, long_row, 7) == sum(v * (i + 1) for i, v in enumerate(long_row)) + 0.5
    assert helper.qb_single_expectation(fn, ctx, 42, 4, row, 7) == -1000  # wrong length
    assert helper.qb_single_expectation(fn, ctx, 42, 3, ["x", 1.0, 2.0], 7) == -2000  # not numbers: NumPy path decides
    assert helper.qb_single_expectation(fn, ctx, 42, 3, 5, 7) == -2000  # not a sequence
    assert helper.qb_single_expectation(fn, ctx, 42, 3, row, 99) == _native.QB_ERR_NOT_FOUND  # native status passed through
