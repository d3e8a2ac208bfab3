"""Bitstring -> float evaluator type accepted by ``B200BitstringCircuitEvaluator``.

Same surface as /root/reference/queasars/circuit_evaluation/bitstring_evaluation.py:7-61 (constructor
arguments, ``evaluate_bitstring``, ``input_length``, length/charset validation and exception type); any
object with ``input_length`` and ``evaluate_bitstring(bitstring=...)`` -- e.g. the reference's own class --
is accepted by the evaluators.
"""
from __future__ import annotations

from typing import Callable


class BitstringEvaluatorException(Exception):
    """Raised for malformed bitstrings."""


class BitstringEvaluator:
    def __init__(self, input_length: int, evaluation_function: Callable[[str], float]):
        self._input_length = int(input_length)
        self._evaluation_function = evaluation_function

    @property
    def input_length(self) -> int:
        return self._input_length

    def evaluate_bitstring(self, bitstring: str) -> float:
        if len(bitstring) != self._input_length:
            raise BitstringEvaluatorException(
                f"Bitstring must be of the length {self._input_length} but was of length {len(bitstring)}!"
            )
        if bitstring.strip("01"):
            raise BitstringEvaluatorException("Bitstring may not contain characters other than '0' or '1'!")
        return self._evaluation_function(bitstring)


class DiagonalEnergyBitstringEvaluator(BitstringEvaluator):
    """Bitstring evaluator whose value is a diagonal energy ``E(b) = sum_t c_t (-1)^{popcount(b & z_t)}`` -- what the
    JSSP examples of the reference compute per string through ``translate_result_bitstring`` + a cost function
    (job_shop_scheduling/domain_wall_hamiltonian_encoder.py:106-144).  It still satisfies the
    ``BitstringEvaluator`` contract string by string, and additionally offers ``evaluate_states(uint64 array)``, which
    ``B200BitstringCircuitEvaluator`` evaluates its ``diagonal_terms`` for all distinct sampled states of a population in ONE
    device call instead of one Python call per string (bit q of a state = character ``n-1-q`` of its bitstring)."""

    def __init__(self, input_length: int, z_masks, coeffs):
        import numpy as np

        from . import expectation as ex

        self._z, self._c = ex.merge_diagonal_terms(np.asarray(z_masks, dtype=np.uint64), np.asarray(coeffs, dtype=np.float64))
        super().__init__(input_length, self._evaluate_one)

    def _evaluate_one(self, bitstring: str) -> float:
        import numpy as np

        return float(self.evaluate_states(np.asarray([int(bitstring, 2)], dtype=np.uint64))[0])

    @property
    def diagonal_terms(self):
        return self._z, self._c

    def evaluate_states(self, states):
        """Host evaluation (vectorised NumPy); ``B200BitstringCircuitEvaluator`` evaluates ``diagonal_terms`` on the GPU."""
        import numpy as np

        from . import expectation as ex

        return ex.diagonal_energies(np.ascontiguousarray(states, dtype=np.uint64).reshape(-1), self._z, self._c)
