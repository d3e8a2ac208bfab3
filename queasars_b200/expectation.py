"""Host-side post-processing of sampled distributions: expectation / CVaR(alpha).

Mirrors /root/reference/queasars/circuit_evaluation/expectation_calculation.py:
  * ``lower_tail_expectation``  <-> ``_get_expectation`` (:14-32): optional sort by value, greedy fill of
    the lower alpha tail, early stop once ``isclose(gathered, alpha)`` (numpy defaults rtol=1e-5, atol=1e-8),
    division by alpha.
  * ``expectation_with_operator`` <-> ``get_expectation_with_operator`` (:35-69)
  * ``expectation_with_bitstring_evaluator`` <-> ``get_expectation_with_bitstring_evaluator`` (:72-103)

The diagonal energies of the *distinct sampled states* are computed vectorised on the host here when the
caller has no device table at hand; the evaluators normally hand in energies gathered on the GPU.
"""
from __future__ import annotations

from typing import Mapping, Optional, Sequence

import numpy as np

_RTOL, _ATOL = 1e-5, 1e-8  # numpy.isclose defaults, which the reference relies on


def _close(a: float, b: float) -> bool:
    return abs(a - b) <= _ATOL + _RTOL * abs(b)


def check_alpha(alpha: float) -> None:
    if alpha <= 0 or 1 < alpha:
        raise ValueError("alpha must be in the range (0, 1]!")


def lower_tail_expectation(probabilities: Sequence[float], values: Sequence[float], alpha: float) -> float:
    probs = np.asarray(probabilities, dtype=np.float64)
    vals = np.asarray(values, dtype=np.float64)
    if not _close(alpha, 1):
        order = np.argsort(vals, kind="stable")
        probs, vals = probs[order], vals[order]
    gathered = 0.0
    acc = 0.0
    for p, v in zip(probs.tolist(), vals.tolist()):
        take = min(alpha - gathered, p)
        acc += take * v
        gathered += take
        if _close(gathered, alpha):
            break
    return acc / alpha


def lower_tail_expectation_arrays(probs: np.ndarray, vals: np.ndarray, alpha: float) -> float:
    """``lower_tail_expectation`` without the per-entry Python loop (10 000-shot distributions): the greedy fill takes whole
    entries while the running mass stays below alpha, stops *early* at the first entry whose running mass is already
    ``isclose`` to alpha, and otherwise clips the entry that crosses alpha.  The running mass is the same sequential float64
    sum as in the loop (``np.cumsum``), so the stopping entry is identical; only the final dot product may differ in the
    last bits."""
    probs = np.asarray(probs, dtype=np.float64)
    vals = np.asarray(vals, dtype=np.float64)
    if probs.size == 0:
        return 0.0
    if not _close(alpha, 1):
        order = np.argsort(vals, kind="stable")
        probs, vals = probs[order], vals[order]
    cum = np.cumsum(probs)
    tol = _ATOL + _RTOL * abs(alpha)
    stop = np.nonzero(cum >= alpha - tol)[0]
    if stop.size == 0:  # the whole distribution weighs less than alpha (cannot happen for a normalised one)
        return float(np.dot(probs, vals)) / alpha
    k = int(stop[0])
    acc = float(np.dot(probs[:k], vals[:k]))
    before = float(cum[k - 1]) if k else 0.0
    acc += min(alpha - before, float(probs[k])) * float(vals[k])
    return acc / alpha


def diagonal_energies(states: np.ndarray, z_masks: np.ndarray, coeffs: np.ndarray) -> np.ndarray:
    """E(k) = sum_j c_j (-1)^{popcount(k & z_j)} for every k in ``states`` (uint64), vectorised."""
    states = np.asarray(states, dtype=np.uint64).reshape(-1, 1)
    z = np.asarray(z_masks, dtype=np.uint64).reshape(1, -1)
    odd = (np.bitwise_count(states & z) & 1).astype(np.float64)
    return (1.0 - 2.0 * odd) @ np.asarray(coeffs, dtype=np.float64)


def merge_diagonal_terms(z_masks: np.ndarray, coeffs: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Sum the coefficients of repeated z-masks (the JSSP encoder emits 346 raw terms for 84 distinct masks at 26
    qubits: SURVEY.md section 8a); first-appearance order is kept."""
    order: dict = {}
    for z, c in zip(np.asarray(z_masks, dtype=np.uint64).tolist(), np.asarray(coeffs, dtype=np.float64).tolist()):
        order[z] = order.get(z, 0.0) + c
    return np.fromiter(order.keys(), dtype=np.uint64, count=len(order)), np.fromiter(order.values(), dtype=np.float64, count=len(order))


def expectation_with_operator(
    distribution: Mapping[int, float],
    z_masks: np.ndarray,
    coeffs: np.ndarray,
    alpha: float = 1.0,
    energies: Optional[np.ndarray] = None,
) -> float:
    check_alpha(alpha)
    states = np.fromiter(distribution.keys(), dtype=np.uint64, count=len(distribution))
    probs = np.fromiter(distribution.values(), dtype=np.float64, count=len(distribution))
    if energies is None:
        energies = diagonal_energies(states, z_masks, coeffs)
    if _close(alpha, 1):
        return float(np.dot(probs, energies))
    return lower_tail_expectation(probs, energies, alpha)


def expectation_with_bitstring_evaluator(distribution, bitstring_evaluator, alpha: float = 1.0, num_bits: Optional[int] = None) -> float:
    """``num_bits=None`` reproduces upstream ``binary_probabilities()`` (pads to the widest observed key);
    the B200 evaluators pass ``num_bits = n_qubits`` so short strings can never reach the user's callable."""
    check_alpha(alpha)
    binary = distribution.binary_probabilities(num_bits=num_bits)
    probs, vals = [], []
    for bitstring, prob in binary.items():
        probs.append(prob)
        vals.append(bitstring_evaluator.evaluate_bitstring(bitstring=bitstring))
    return lower_tail_expectation(probs, vals, alpha)
