"""EVQE genome types for the direct genome -> gate-list front end and for synthetic workloads.

The reference's genome (``EVQEIndividual`` = layers of one gate per qubit:
/root/reference/queasars/minimum_eigensolvers/evqe/quantum_circuit/quantum_gate.py,
.../circuit_layer.py, .../evolutionary_algorithm/individual.py) can be passed to
``gate_list.from_evqe_individual`` directly by duck typing.  These lightweight equivalents exist so that
benchmarks and tests can create populations with the reference's *generation rules and RNG call sequence*
(circuit_layer.py:38-135, individual.py:34-66, population.py:33-77, utility/random.py) on machines where
the reference package is not installed; tests/golden/genomes.json pins them against the reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from random import Random
from typing import Optional, Sequence

from .circuit import CU3Gate, Parameter, QuantumCircuit

_SEED_TOP = 2147483647


@dataclass(frozen=True)
class Gene:
    qubit_index: int

    @staticmethod
    def n_parameters() -> int:
        return 0


@dataclass(frozen=True)
class Identity(Gene):
    pass


@dataclass(frozen=True)
class Rotation(Gene):
    @staticmethod
    def n_parameters() -> int:
        return 3


@dataclass(frozen=True)
class Control(Gene):
    controlled_qubit_index: int = -1


@dataclass(frozen=True)
class ControlledRotation(Gene):
    control_qubit_index: int = -1

    @staticmethod
    def n_parameters() -> int:
        return 3


@dataclass(frozen=True)
class Layer:
    n_qubits: int
    gates: tuple

    @property
    def n_parameters(self) -> int:
        return sum(g.n_parameters() for g in self.gates)

    @staticmethod
    def random(n_qubits: int, previous: Optional["Layer"], seed: Optional[int]) -> "Layer":
        rng = Random(seed)
        slots: list[Gene] = [Identity(q) for q in range(n_qubits)]
        waiting: list[int] = []
        for q in range(n_qubits):
            free_choice = previous is None or isinstance(previous.gates[q], (Control, ControlledRotation))
            if free_choice and rng.choice(("rotation", "controlled")) == "rotation":
                slots[q] = Rotation(q)
            else:
                waiting.append(q)
        while len(waiting) > 1:
            target, control = rng.sample(waiting, 2)
            pair = (ControlledRotation(target, control), Control(control, target))
            if previous is not None and (previous.gates[target] == pair[0] or previous.gates[control] == pair[1]):
                continue  # would only duplicate the previous layer's gate: draw again
            slots[target], slots[control] = pair
            waiting.remove(target)
            waiting.remove(control)
        for q in waiting:
            blocked = previous is not None and isinstance(previous.gates[q], Rotation)
            slots[q] = Identity(q) if blocked else Rotation(q)
        return Layer(n_qubits, tuple(slots))


@dataclass(frozen=True)
class Individual:
    n_qubits: int
    layers: tuple
    parameter_values: tuple

    @property
    def layer_parameter_indices(self) -> dict:
        out, start = {}, 0
        for i, layer in enumerate(self.layers):
            out[i] = tuple(range(start, start + layer.n_parameters))
            start += layer.n_parameters
        return out

    @staticmethod
    def random(n_qubits: int, n_layers: int, randomize_parameter_values: bool, seed: Optional[int] = None) -> "Individual":
        rng = Random(seed)
        layers: list[Layer] = []
        for _ in range(n_layers):
            layers.append(Layer.random(n_qubits, layers[-1] if layers else None, rng.randint(0, _SEED_TOP)))
        count = sum(layer.n_parameters for layer in layers)
        values = tuple(2 * math.pi * rng.random() for _ in range(count)) if randomize_parameter_values else (0,) * count
        return Individual(n_qubits, tuple(layers), values)

    def layer_values(self, layer_id: int) -> tuple:
        return tuple(self.parameter_values[i] for i in self.layer_parameter_indices[layer_id % len(self.layers)])

    def to_circuit(self, parameterized_layers: Optional[set] = None) -> QuantumCircuit:
        """The circuit ``get_partially_parameterized_quantum_circuit`` would build (individual.py:288-322):
        ops ``u`` / ``cu3`` / ``id``, parameters named ``layer{L}_q{Q}_{theta,phi,lambda}``, layers outside
        ``parameterized_layers`` pre-bound with the stored values in name-sorted order (circuit_layer.py:233-235)."""
        n_layers = len(self.layers)
        chosen = set(range(n_layers)) if parameterized_layers is None else {i % n_layers for i in parameterized_layers}
        circuit = QuantumCircuit(self.n_qubits)
        for i, layer in enumerate(self.layers):
            sub = QuantumCircuit(self.n_qubits, name=f"layer_{i}")
            for gate in layer.gates:
                pre = f"layer{i}_q{gate.qubit_index}_"
                if isinstance(gate, Rotation):
                    sub.u(Parameter(pre + "theta"), Parameter(pre + "phi"), Parameter(pre + "lambda"), gate.qubit_index)
                elif isinstance(gate, ControlledRotation):
                    sub.append(
                        CU3Gate(Parameter(pre + "theta"), Parameter(pre + "phi"), Parameter(pre + "lambda")),
                        (gate.control_qubit_index, gate.qubit_index),
                    )
                elif isinstance(gate, Identity):
                    sub.id(gate.qubit_index)
            if i not in chosen:
                sub.assign_parameters(list(self.layer_values(i)), inplace=True)
            circuit.compose(sub, inplace=True)
        return circuit


def random_population(n_qubits: int, n_layers: int, n_individuals: int, randomize_parameter_values: bool, seed: Optional[int] = None) -> list:
    rng = Random(seed)
    return [Individual.random(n_qubits, n_layers, randomize_parameter_values, rng.randint(0, _SEED_TOP)) for _ in range(n_individuals)]


def ising_operator(n_qubits: int, seed: int = 1234):
    """Synthetic diagonal Ising Hamiltonian of BASELINE config C2: h_i ~ N(0,1) on every
    qubit, J_ij ~ N(0,1) on every pair, ``numpy.random.default_rng(seed)``."""
    import numpy as np

    from .operators import SparsePauliOp

    rng = np.random.default_rng(seed)
    xs, zs, cs = [], [], []
    for i in range(n_qubits):
        xs.append(0), zs.append(1 << i), cs.append(float(rng.normal()))
    for i in range(n_qubits):
        for j in range(i + 1, n_qubits):
            xs.append(0), zs.append((1 << i) | (1 << j)), cs.append(float(rng.normal()))
    return SparsePauliOp._raw(n_qubits, xs, zs, cs)


def tfim_operator(n_qubits: int, field: float = 0.5):
    """Open-chain transverse-field Ising Pauli sum of config C3: -sum Z_i Z_{i+1} - field * sum X_i."""
    from .operators import SparsePauliOp

    xs, zs, cs = [], [], []
    for i in range(n_qubits - 1):
        xs.append(0), zs.append(3 << i), cs.append(-1.0)
    for i in range(n_qubits):
        xs.append(1 << i), zs.append(0), cs.append(-float(field))
    return SparsePauliOp._raw(n_qubits, xs, zs, cs)
