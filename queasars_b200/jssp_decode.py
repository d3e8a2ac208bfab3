"""Vectorised decoding of sampled basis states into job-shop schedules -- the array form of
``JSSPDomainWallHamiltonianEncoder.translate_result_bitstring``
(/root/reference/queasars/job_shop_scheduling/domain_wall_hamiltonian_encoder.py:106-144) and
``DomainWallVariable.value_from_bitlist`` (/root/reference/queasars/utility/domain_wall_variables.py:145-170), SURVEY.md
section 8f-4.  The sampler route hands back up to ``shots`` distinct basis states per individual as integers
(``B200SamplerV2.sample_indices``); the reference decodes them one bitstring at a time in Python.

Bit convention: the reference reverses the measured bitstring (``bitstring[::-1]``), so ``bit_list[i]`` is qubit ``i`` = bit
``i`` of the basis-state integer.  A domain-wall variable on qubits ``[s, s + w)`` holds ``values[d]`` where ``d`` is the number
of leading ones, and is *invalid* (operation unscheduled, ``None`` in the reference) unless every bit after the wall is zero.
"""
from __future__ import annotations

from typing import NamedTuple, Sequence

import numpy as np


class OperationSlot(NamedTuple):
    job: str
    operation: str
    start: int  # first qubit of the variable
    width: int  # number of qubits
    values: tuple  # width + 1 start times


def layout_of(encoder) -> list[OperationSlot]:
    """Domain-wall layout of a reference ``JSSPDomainWallHamiltonianEncoder`` (duck typed: jobs in instance order, operations
    in job order -- the order ``translate_result_bitstring`` walks)."""
    if not getattr(encoder, "_encoding_prepared", True):
        encoder._prepare_encoding()
    slots = []
    for job in encoder.jssp_instance.jobs:
        for operation in job.operations:
            var = encoder._operation_start_variables[operation]
            slots.append(OperationSlot(job.name, operation.name, int(var._qubit_start_index), int(var.n_qubits), tuple(var.values)))
    return slots


def decode_start_times(states, layout: Sequence[OperationSlot]) -> np.ndarray:
    """states: basis-state integers [n_states] -> int64 [n_states, n_operations]: the start time of every operation, or -1 where
    the reference returns an ``UnscheduledOperation`` (bits set behind the domain wall)."""
    k = np.ascontiguousarray(states, dtype=np.uint64).reshape(-1)
    out = np.empty((k.size, len(layout)), dtype=np.int64)
    for col, slot in enumerate(layout):
        if slot.width == 0:
            out[:, col] = slot.values[0]
            continue
        field = (k >> np.uint64(slot.start)) & np.uint64((1 << slot.width) - 1)
        # number of trailing ones of the field = position of the domain wall; valid iff the field is exactly 2^d - 1
        wall = np.zeros(k.size, dtype=np.int64)
        run = np.ones(k.size, dtype=bool)
        for i in range(slot.width):
            run &= ((field >> np.uint64(i)) & np.uint64(1)).astype(bool)
            wall += run
        valid = field == ((np.uint64(1) << wall.astype(np.uint64)) - np.uint64(1))
        values = np.asarray(slot.values, dtype=np.int64)
        out[:, col] = np.where(valid, values[wall], -1)
    return out


def makespans(start_times: np.ndarray, durations: Sequence[int]) -> np.ndarray:
    """max over operations of start + duration, -1 for states with an unscheduled operation."""
    d = np.asarray(durations, dtype=np.int64)
    ok = np.all(start_times >= 0, axis=1)
    return np.where(ok, np.max(start_times + d[None, :], axis=1), -1)
