"""EVQE genome -> circuit description, restated for the oracle.  TEST INFRASTRUCTURE ONLY.

Follows (paths relative to /root/reference/queasars/minimum_eigensolvers/evqe):
  * quantum_circuit/quantum_gate.py:78-79, 96-102, 125-126, 157-165  which instruction each gene emits,
    the parameter names ``layer{L}_q{Q}_theta|_phi|_lambda`` and cu3 qargs = (control, target)
  * quantum_circuit/circuit_layer.py:38-135   random layer generation (RNG call sequence reproduced)
  * quantum_circuit/circuit_layer.py:191-235  parameterised vs. numerically pre-bound layers
  * evolutionary_algorithm/individual.py:34-66, 288-322   random individual, partial parameterisation
  * evolutionary_algorithm/population.py:33-77  seed chaining
  * queasars/utility/random.py:7-15  new seed = randint(0, 2**31-1)

A genome layer is a tuple with one gene per qubit:
  ("id",) | ("rot",) | ("ctrl", target_qubit) | ("crot", control_qubit)
Pinned against the reference's own classes by tests/golden/genomes.json.
"""
from __future__ import annotations

import math
from random import Random
from typing import Optional, Sequence

Gene = tuple
Layer = tuple  # of Gene
SEED_MAX = 2147483647


def gene_n_params(gene: Gene) -> int:
    return 3 if gene[0] in ("rot", "crot") else 0


def layer_n_params(layer: Layer) -> int:
    return sum(gene_n_params(g) for g in layer)


def random_layer(n_qubits: int, previous: Optional[Layer], seed: Optional[int]) -> Layer:
    rng = Random(seed)
    genes: list[Gene] = [("id",)] * n_qubits
    pending: list[int] = []  # qubits waiting to take part in a controlled rotation
    for q in range(n_qubits):
        forced = previous is not None and previous[q][0] in ("rot", "id")
        if forced:
            pending.append(q)
        elif rng.choice(["rot", "crot"]) == "crot":
            pending.append(q)
        else:
            genes[q] = ("rot",)
    while len(pending) >= 2:
        tgt, ctl = rng.sample(pending, 2)
        crot, ctrl = ("crot", ctl), ("ctrl", tgt)
        # reject a pair that repeats the previous layer's gene on either of the two qubits
        if previous is None or (previous[tgt] != crot and previous[ctl] != ctrl):
            genes[ctl] = ctrl
            genes[tgt] = crot
            pending.remove(tgt)
            pending.remove(ctl)
    if pending:
        q = pending[0]
        genes[q] = ("id",) if (previous is not None and previous[q][0] == "rot") else ("rot",)
    return tuple(genes)


def random_individual(n_qubits: int, n_layers: int, randomize: bool, seed: Optional[int]):
    """-> (layers, parameter_values)"""
    rng = Random(seed)
    layers: list[Layer] = []
    prev: Optional[Layer] = None
    for _ in range(n_layers):
        prev = random_layer(n_qubits, prev, rng.randint(0, SEED_MAX))
        layers.append(prev)
    n_params = sum(layer_n_params(layer) for layer in layers)
    values = tuple(2 * math.pi * rng.random() for _ in range(n_params)) if randomize else (0,) * n_params
    return tuple(layers), values


def random_population(n_qubits: int, n_layers: int, n_individuals: int, randomize: bool, seed: Optional[int]):
    rng = Random(seed)
    return [random_individual(n_qubits, n_layers, randomize, rng.randint(0, SEED_MAX)) for _ in range(n_individuals)]


def layer_instructions(layer: Layer, layer_id: int, values: Optional[Sequence[float]] = None) -> list[tuple]:
    """Instructions of one layer in qubit order.  ``values`` None -> named parameters; otherwise the
    layer's genome-order values (theta, phi, lambda per gate), bound *positionally in name-sorted order*
    exactly like ``assign_parameters(sequence)`` does at circuit_layer.py:233-235."""
    named = []
    for q, gene in enumerate(layer):
        pre = f"layer{layer_id}_q{q}_"
        if gene[0] == "id":
            named.append(("id", (q,), ()))
        elif gene[0] == "rot":
            named.append(("u", (q,), (pre + "theta", pre + "phi", pre + "lambda")))
        elif gene[0] == "crot":
            named.append(("cu3", (gene[1], q), (pre + "theta", pre + "phi", pre + "lambda")))
    if values is None:
        return named
    names = sorted({p for _, _, ps in named for p in ps})
    if len(values) != len(names):
        raise ValueError("wrong number of layer parameter values")
    table = dict(zip(names, values))
    return [(nm, qs, tuple(float(table[p]) for p in ps)) for nm, qs, ps in named]


def individual_circuit(layers: Sequence[Layer], values: Sequence[float], parameterized: Optional[set] = None):
    """individual.py:288-322: layers in ``parameterized`` keep named parameters, the others are pre-bound
    with the individual's stored values.  Returns the instruction list (ops named u / cu3 / id)."""
    n_layers = len(layers)
    parameterized = set(range(n_layers)) if parameterized is None else {i % n_layers for i in parameterized}
    out: list[tuple] = []
    offset = 0
    for i, layer in enumerate(layers):
        npar = layer_n_params(layer)
        if i in parameterized:
            out += layer_instructions(layer, i)
        else:
            out += layer_instructions(layer, i, values[offset : offset + npar])
        offset += npar
    return out


def layer_value_slice(layers: Sequence[Layer], layer_id: int) -> slice:
    layer_id %= len(layers)
    start = sum(layer_n_params(layer) for layer in layers[:layer_id])
    return slice(start, start + layer_n_params(layers[layer_id]))
