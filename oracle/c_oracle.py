"""ctypes wrapper of the C/OpenMP oracle (oracle/c/statevector.c).  TEST / BASELINE INFRASTRUCTURE ONLY.

``evaluate(instructions, n, values, table)`` runs a circuit given in the oracle's instruction format
(oracle/qiskit_semantics.py) restricted to one-qubit gates with at most one control (the EVQE gate set ``u`` / ``cu3`` /
``id`` and the other controlled standard gates), and returns sum_k |psi_k|^2 table[k].  The gate matrices are produced
by the NumPy oracle's ``gate_matrix`` so both oracles share one definition of the gates."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from . import qiskit_semantics as oq

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "c", "liboracle.so")
_lib = None

_CONTROLLED = {"cu3": "u3", "cx": "x", "cy": "y", "cz": "z", "ch": "h", "cp": "p", "cu1": "p", "crz": "rz", "crx": "rx", "cry": "ry"}


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "c", "statevector.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(HERE, "c"), "-B", "liboracle.so"], check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(LIB)
        lib.oracle_run_circuit.restype = ctypes.c_double
        lib.oracle_run_circuit.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_diag_table.restype = None
        lib.oracle_diag_table.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_pauli_sum.restype = ctypes.c_double
        lib.oracle_pauli_sum.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_sample.restype = None
        lib.oracle_sample.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        lib.oracle_max_threads.restype = ctypes.c_int
        lib.oracle_set_threads.argtypes = [ctypes.c_int]
        _lib = lib
    return _lib


def lower(instructions, values):
    targets, controls, mats = [], [], []
    for name, qubits, params in oq.bind(instructions, values):
        if name in ("id", "i", "barrier", "measure"):
            continue
        if name in _CONTROLLED:
            m = oq.gate_matrix(_CONTROLLED[name], params)
            controls.append(qubits[0]), targets.append(qubits[1])
        elif len(qubits) == 1:
            m = oq.gate_matrix(name, params)
            controls.append(-1), targets.append(qubits[0])
        else:
            raise ValueError(f"C oracle: unsupported gate {name}")
        mats.append(np.asarray(m, dtype=np.complex128).reshape(-1))
    return (
        np.asarray(targets, dtype=np.int32),
        np.asarray(controls, dtype=np.int32),
        np.ascontiguousarray(np.asarray(mats, dtype=np.complex128).reshape(-1)).view(np.float64),
    )


def evaluate(instructions, n, values, table=None, state=None):
    lib = load()
    targets, controls, mats = lower(instructions, values)
    if state is None:
        state = np.empty(1 << n, dtype=np.complex128)
    tptr = table.ctypes.data if table is not None else None
    value = lib.oracle_run_circuit(state.ctypes.data, n, len(targets), targets.ctypes.data, controls.ctypes.data, mats.ctypes.data, tptr)
    return value, state


def diag_table(n, diag_terms):
    lib = load()
    z = np.asarray([t[0] for t in diag_terms], dtype=np.uint64)
    c = np.asarray([t[1] for t in diag_terms], dtype=np.float64)
    table = np.empty(1 << n, dtype=np.float64)
    lib.oracle_diag_table(table.ctypes.data, n, len(z), z.ctypes.data, c.ctypes.data)
    return table


def pauli_sum(state, n, terms):
    """Re sum_j c_j <psi|P_j|psi> for ``terms`` = [(label, coeff), ...] (labels little-endian like oq.pauli_label_to_masks): one
    pass over the state per term, like the upstream estimator."""
    lib = load()
    masks = [oq.pauli_label_to_masks(label) for label, _ in terms]
    x = np.asarray([m[0] for m in masks], dtype=np.uint64)
    z = np.asarray([m[1] for m in masks], dtype=np.uint64)
    cr = np.asarray([complex(c).real for _, c in terms], dtype=np.float64)
    ci = np.asarray([complex(c).imag for _, c in terms], dtype=np.float64)
    state = np.ascontiguousarray(state, dtype=np.complex128)
    return float(lib.oracle_pauli_sum(state.ctypes.data, n, len(x), x.ctypes.data, z.ctypes.data, cr.ctypes.data, ci.ctypes.data))


def sample_indices(state, n, uniforms):
    """cumsum -> /= last -> searchsorted(side='right'), the chunk-free restatement of oq.sample_indices."""
    lib = load()
    state = np.ascontiguousarray(state, dtype=np.complex128)
    uniforms = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1)
    out = np.empty(uniforms.size, dtype=np.int64)
    cdf = np.empty(1 << n, dtype=np.float64)
    lib.oracle_sample(state.ctypes.data, n, uniforms.size, uniforms.ctypes.data, out.ctypes.data, cdf.ctypes.data)
    return out


def max_threads() -> int:
    return int(load().oracle_max_threads())


def set_threads(n: int) -> None:
    load().oracle_set_threads(int(n))
