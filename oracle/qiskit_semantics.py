"""NumPy restatement of the Qiskit Estimator/Sampler semantics the reference's evaluators consume.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  **Parity unpinned** at the primitive boundary.

Reference call sites this restates (all paths relative to /root/reference):
  * queasars/circuit_evaluation/circuit_evaluation.py:29-59   measure_quasi_distributions
  * queasars/circuit_evaluation/circuit_evaluation.py:200-215 OperatorCircuitEvaluator.evaluate_circuits
  * queasars/circuit_evaluation/expectation_calculation.py:14-103
  * [upstream, un-vendored] qiskit 2.4.2 ``StatevectorEstimator._run_pub`` / ``StatevectorSampler._run_pub``,
    ``quantum_info.Statevector`` (evolve, expectation_value, probabilities, sample_memory),
    ``qiskit.result.sampled_expectation_value``,
    qiskit-algorithms 0.4.0 ``diagonal_estimator._evaluate_sparsepauli`` -- published algorithms restated.

Conventions: little-endian (qubit 0 = least significant bit of the amplitude index; the right-most
character of a Pauli label / bitstring is qubit 0), cf. queasars/utility/pauli_strings.py:38-41.

A circuit is described independently of the product's IR as a list of instructions
``(name, qubits, params)`` where every param is a float, a parameter *name* (str), or an affine
triple ``(coeff, name, const)``.  Flat parameter-value lists bind to ``sorted(names)`` (plain string
compare), which is what ``QuantumCircuit.parameters`` + ``EstimatorPub/SamplerPub.coerce`` do upstream.
"""
from __future__ import annotations

import cmath
import math
from typing import Callable, Iterable, Sequence, Union

import numpy as np

ParamLike = Union[float, str, tuple]
Instruction = tuple  # (name, qubits, params)

_S2 = 1.0 / math.sqrt(2.0)
_P0 = np.array([[1, 0], [0, 0]], dtype=complex)
_P1 = np.array([[0, 0], [0, 1]], dtype=complex)
_I2 = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)


# --------------------------------------------------------------------------------------------------
# gate matrices  [upstream qiskit.circuit.library.standard_gates]
# --------------------------------------------------------------------------------------------------
def u_matrix(theta: float, phi: float, lam: float) -> np.ndarray:
    """UGate.__array__: [[cos, -e^{i lam} sin], [e^{i phi} sin, e^{i(phi+lam)} cos]] with half angle."""
    c, s = math.cos(theta / 2), math.sin(theta / 2)
    return np.array(
        [[c, -cmath.exp(1j * lam) * s], [cmath.exp(1j * phi) * s, cmath.exp(1j * (phi + lam)) * c]],
        dtype=complex,
    )


def _controlled(u: np.ndarray) -> np.ndarray:
    """control = first qarg (low bit of the 2-qubit index), target = second qarg."""
    return np.kron(_I2, _P0) + np.kron(u, _P1)


def gate_matrix(name: str, params: Sequence[float]) -> np.ndarray:
    p = list(params)
    if name in ("id", "i"):
        return _I2.copy()
    if name in ("u", "u3"):
        return u_matrix(*p)
    if name == "u2":
        return u_matrix(math.pi / 2, p[0], p[1])
    if name in ("u1", "p"):
        return np.array([[1, 0], [0, cmath.exp(1j * p[0])]], dtype=complex)
    if name == "rz":
        return np.array([[cmath.exp(-0.5j * p[0]), 0], [0, cmath.exp(0.5j * p[0])]], dtype=complex)
    if name == "rx":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -1j * s], [-1j * s, c]], dtype=complex)
    if name == "ry":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return np.array([[c, -s], [s, c]], dtype=complex)
    if name == "x":
        return _X.copy()
    if name == "y":
        return _Y.copy()
    if name == "z":
        return _Z.copy()
    if name == "h":
        return _S2 * np.array([[1, 1], [1, -1]], dtype=complex)
    if name == "s":
        return np.diag([1, 1j]).astype(complex)
    if name == "sdg":
        return np.diag([1, -1j]).astype(complex)
    if name == "t":
        return np.diag([1, cmath.exp(0.25j * math.pi)]).astype(complex)
    if name == "tdg":
        return np.diag([1, cmath.exp(-0.25j * math.pi)]).astype(complex)
    if name == "sx":
        return 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex)
    if name == "sxdg":
        return 0.5 * np.array([[1 - 1j, 1 + 1j], [1 + 1j, 1 - 1j]], dtype=complex)
    if name == "cx":
        return _controlled(_X)
    if name == "cy":
        return _controlled(_Y)
    if name == "cz":
        return _controlled(_Z)
    if name == "ch":
        return _controlled(gate_matrix("h", []))
    if name in ("cp", "cu1"):
        return _controlled(gate_matrix("p", p))
    if name == "crz":
        return _controlled(gate_matrix("rz", p))
    if name == "crx":
        return _controlled(gate_matrix("rx", p))
    if name == "cry":
        return _controlled(gate_matrix("ry", p))
    if name == "cu3":
        return _controlled(u_matrix(*p))
    if name == "cu":
        return _controlled(cmath.exp(1j * p[3]) * u_matrix(p[0], p[1], p[2]))
    if name == "swap":
        return np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)
    if name == "rzz":
        a, b = cmath.exp(-0.5j * p[0]), cmath.exp(0.5j * p[0])
        return np.diag([a, b, b, a]).astype(complex)
    if name == "rxx":
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return c * np.eye(4, dtype=complex) - 1j * s * np.kron(_X, _X)
    if name == "rzx":  # Z on the first qarg, X on the second
        c, s = math.cos(p[0] / 2), math.sin(p[0] / 2)
        return c * np.eye(4, dtype=complex) - 1j * s * np.kron(_X, _Z)
    if name == "ecr":
        return _S2 * np.array(
            [[0, 1, 0, 1j], [1, 0, -1j, 0], [0, 1j, 0, 1], [-1j, 0, 1, 0]], dtype=complex
        )
    raise ValueError(f"oracle: unsupported gate '{name}'")


# --------------------------------------------------------------------------------------------------
# parameter binding
# --------------------------------------------------------------------------------------------------
def parameter_names(circuit: Iterable[Instruction]) -> list[str]:
    """``QuantumCircuit.parameters`` order: unique names, sorted by plain string comparison."""
    names = set()
    for _, _, params in circuit:
        for prm in params:
            if isinstance(prm, str):
                names.add(prm)
            elif isinstance(prm, tuple):
                names.add(prm[1])
    return sorted(names)


def bind(circuit: Sequence[Instruction], values: Sequence[float]) -> list[Instruction]:
    names = parameter_names(circuit)
    if len(values) != len(names):
        raise ValueError(f"oracle: {len(values)} values for {len(names)} parameters")
    table = dict(zip(names, (float(v) for v in values)))
    out = []
    for name, qubits, params in circuit:
        num = []
        for prm in params:
            if isinstance(prm, str):
                num.append(table[prm])
            elif isinstance(prm, tuple):
                num.append(prm[0] * table[prm[1]] + prm[2])
            else:
                num.append(float(prm))
        out.append((name, tuple(qubits), tuple(num)))
    return out


# --------------------------------------------------------------------------------------------------
# statevector evolution  [upstream Statevector._evolve_instruction: einsum on the qargs' axes]
# --------------------------------------------------------------------------------------------------
def apply_matrix(state: np.ndarray, n: int, mat: np.ndarray, qubits: Sequence[int]) -> np.ndarray:
    k = len(qubits)
    psi = state.reshape((2,) * n)
    m = mat.reshape((2,) * (2 * k))
    # matrix row/col bit j (msb first) <-> qubits[k-1-j]; amplitude axis of qubit q is n-1-q
    tgt_axes = [n - 1 - qubits[k - 1 - j] for j in range(k)]
    out = np.tensordot(m, psi, axes=(list(range(k, 2 * k)), tgt_axes))
    out = np.moveaxis(out, list(range(k)), tgt_axes)
    return np.ascontiguousarray(out).reshape(-1)


def statevector(circuit: Sequence[Instruction], n: int, values: Sequence[float] = (), dtype=np.complex128):
    state = np.zeros(1 << n, dtype=dtype)
    state[0] = 1.0
    for name, qubits, params in bind(circuit, values):
        if name in ("barrier", "measure", "id", "i", "delay"):
            continue
        mat = gate_matrix(name, params).astype(dtype)
        state = apply_matrix(state, n, mat, qubits)
    return state


# --------------------------------------------------------------------------------------------------
# observables
# --------------------------------------------------------------------------------------------------
def pauli_label_to_masks(label: str) -> tuple[int, int, int]:
    """label right-most char = qubit 0 -> (x_mask, z_mask, n_y)."""
    x = z = ny = 0
    n = len(label)
    for pos, ch in enumerate(label):
        q = n - 1 - pos
        if ch == "X":
            x |= 1 << q
        elif ch == "Z":
            z |= 1 << q
        elif ch == "Y":
            x |= 1 << q
            z |= 1 << q
            ny += 1
        elif ch != "I":
            raise ValueError(f"bad Pauli label char {ch!r}")
    return x, z, ny


def _parity(arr: np.ndarray) -> np.ndarray:
    a = arr.astype(np.uint64).copy()
    for s in (32, 16, 8, 4, 2, 1):
        a ^= a >> np.uint64(s)
    return (a & np.uint64(1)).astype(np.int8)


def pauli_expectation(state: np.ndarray, label: str) -> complex:
    """<psi|P|psi>  [upstream expval_pauli_no_x / expval_pauli_with_x]; P|k> = i^{nY} (-1)^{pc(k&z)} |k^x>."""
    x, z, ny = pauli_label_to_masks(label)
    idx = np.arange(state.size, dtype=np.uint64)
    sign = 1.0 - 2.0 * _parity(idx & np.uint64(z))
    if x == 0:
        return complex(np.sum((state.real**2 + state.imag**2) * sign))
    partner = state[(idx ^ np.uint64(x)).astype(np.int64)]
    return complex((1j**ny) * np.sum(np.conj(partner) * sign * state))


def estimator_expectation(state: np.ndarray, terms: Sequence[tuple[str, complex]]) -> float:
    """Re sum_j c_j <psi|P_j|psi>  (the imaginary part is dropped: circuit_evaluation.py:215)."""
    total = 0.0 + 0.0j
    for label, coeff in terms:
        total += complex(coeff) * pauli_expectation(state, label)
    return float(total.real)


def estimator_value(ev: float, precision: float, seed) -> float:
    """[upstream StatevectorEstimator]: Gaussian noise of std ``precision`` from a fresh default_rng(seed)."""
    if precision == 0:
        return ev
    return float(np.random.default_rng(seed).normal(ev, precision))


def diagonal_energy(state_index: int, diag_terms: Sequence[tuple[int, float]]) -> float:
    """E(k) = sum_j c_j (-1)^{popcount(k & z_j)}  [upstream _evaluate_sparsepauli, .real]."""
    e = 0.0
    for zmask, coeff in diag_terms:
        e += coeff * (1 - 2 * (bin(state_index & zmask).count("1") & 1))
    return e


def diagonal_table(n: int, diag_terms: Sequence[tuple[int, float]]) -> np.ndarray:
    idx = np.arange(1 << n, dtype=np.uint64)
    out = np.zeros(1 << n, dtype=np.float64)
    for zmask, coeff in diag_terms:
        out += coeff * (1.0 - 2.0 * _parity(idx & np.uint64(zmask)))
    return out


def diag_terms_from_labels(terms: Sequence[tuple[str, complex]]) -> list[tuple[int, float]]:
    out = []
    for label, coeff in terms:
        x, z, _ = pauli_label_to_masks(label)
        if x:
            raise ValueError("operator is not diagonal")
        out.append((z, float(np.real(coeff))))
    return out


# --------------------------------------------------------------------------------------------------
# sampling  [upstream Statevector.sample_memory -> numpy Generator.choice(p=...)]
# --------------------------------------------------------------------------------------------------
def sample_indices(state: np.ndarray, shots: int, seed=None, uniforms: np.ndarray | None = None) -> np.ndarray:
    probs = state.real.astype(np.float64) ** 2 + state.imag.astype(np.float64) ** 2
    cdf = probs.cumsum()
    cdf /= cdf[-1]
    if uniforms is None:
        uniforms = np.random.default_rng(seed).random(shots)
    return cdf.searchsorted(uniforms, side="right").astype(np.int64)


def counts_from_indices(indices: np.ndarray, n: int) -> dict[str, int]:
    vals, cnts = np.unique(indices, return_counts=True)
    return {format(int(v), f"0{n}b"): int(c) for v, c in zip(vals, cnts)}


def quasi_distribution(counts: dict[str, int], shots: int) -> dict[int, float]:
    """circuit_evaluation.py:56-59: QuasiDistribution({bitstring: count/shots}) -> integer keys."""
    return {int(b, 2): c / shots for b, c in counts.items()}


# --------------------------------------------------------------------------------------------------
# expectation / CVaR post-processing  (expectation_calculation.py)
# --------------------------------------------------------------------------------------------------
def cvar_accumulate(state_list: list[tuple[object, float, float]], alpha: float) -> float:
    """expectation_calculation.py:14-32 -- lower-alpha-tail mean with isclose() stop."""
    if not np.isclose(alpha, 1):
        state_list = sorted(state_list, key=lambda t: t[2])
    gathered = 0.0
    acc = 0.0
    for _, prob, value in state_list:
        take = min(alpha - gathered, prob)
        acc += take * value
        gathered += take
        if np.isclose(gathered, alpha):
            break
    return acc / alpha


def expectation_with_operator(dist: dict[int, float], diag_terms, alpha: float = 1.0) -> float:
    """expectation_calculation.py:35-69."""
    if alpha <= 0 or alpha > 1:
        raise ValueError("alpha must be in the range (0, 1]!")
    if np.isclose(alpha, 1):
        # [upstream sampled_expectation_value]: sum_b p(b) sum_j c_j (-1)^{pc(b & z_j)}
        return float(sum(p * diagonal_energy(k, diag_terms) for k, p in dist.items()))
    evals = [(k, p, diagonal_energy(k, diag_terms)) for k, p in dist.items()]
    evals = sorted(evals, key=lambda t: t[2])
    return cvar_accumulate(evals, alpha)


def expectation_with_bitstring_function(
    dist: dict[int, float], n: int, fn: Callable[[str], float], alpha: float = 1.0, fixed_width: bool = True
) -> float:
    """expectation_calculation.py:72-103.  ``fixed_width=False`` reproduces the upstream
    ``binary_probabilities()`` padding quirk (pads to the longest *observed* key only)."""
    if alpha <= 0 or alpha > 1:
        raise ValueError("alpha must be in the range (0, 1]!")
    width = n if fixed_width else max(1, max(dist).bit_length())
    evals = [(format(k, f"0{width}b"), p, fn(format(k, f"0{width}b"))) for k, p in dist.items()]
    return cvar_accumulate(evals, alpha)
