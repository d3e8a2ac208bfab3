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
