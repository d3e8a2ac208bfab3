"""Pins the oracle (oracle/) against the fixtures produced by the reference's own code
(tests/golden/make_golden.py) and against analytic known answers."""
import math

import numpy as np
import pytest

from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq


# ---------------------------------------------------------------- genome / naming / binding order
def _layers(entry):
    return tuple(tuple(tuple(g) for g in layer) for layer in entry["layers"])


def test_random_population_matches_reference(genome_golden):
    pop = og.random_population(4, 2, 10, True, 0)
    for (layers, values), entry in zip(pop, genome_golden["population_4q_2l_seed0"]):
        assert layers == _layers(entry)
        assert list(values) == entry["parameter_values"]
    pop12 = og.random_population(12, 3, 3, True, 11)
    for (layers, values), entry in zip(pop12, genome_golden["population_12q_3l_seed11"]):
        assert layers == _layers(entry)
        assert list(values) == entry["parameter_values"]
    pop20 = og.random_population(20, 2, 2, True, 0)
    for (layers, values), entry in zip(pop20, genome_golden["population_20q_2l_seed0_genes_only"]):
        assert layers == _layers(entry)
        assert list(values) == entry["parameter_values"]


def test_random_individual_matches_reference(genome_golden):
    layers, values = og.random_individual(4, 2, False, 0)
    assert layers == _layers(genome_golden["individual_4q_2l_seed0"])
    assert all(v == 0 for v in values)
    layers, values = og.random_individual(3, 12, True, 5)
    assert layers == _layers(genome_golden["individual_3q_12l_seed5"])


def _norm_ops(ops):
    return [(name, tuple(qs), tuple(ps)) for name, qs, ps in ops]


@pytest.mark.parametrize("key", ["population_4q_2l_seed0", "population_12q_3l_seed11"])
def test_circuit_instructions_and_parameter_order(genome_golden, key):
    for entry in genome_golden[key]:
        layers, values = _layers(entry), entry["parameter_values"]
        full = og.individual_circuit(layers, values)
        assert _norm_ops(full) == _norm_ops(entry["full_ops"])
        assert oq.parameter_names(full) == entry["full_parameters"]
        part = og.individual_circuit(layers, values, set(entry["partial_layers"]))
        assert oq.parameter_names(part) == entry["partial_parameters"]
        for (n1, q1, p1), (n2, q2, p2) in zip(part, entry["partial_ops"]):
            assert n1 == n2 and tuple(q1) == tuple(q2)
            assert list(p1) == list(p2)


def test_layer_ordering_quirk_for_many_layers(genome_golden):
    entry = genome_golden["individual_3q_12l_seed5"]
    full = og.individual_circuit(_layers(entry), entry["parameter_values"])
    names = oq.parameter_names(full)
    assert names == entry["full_parameters"]
    # 'layer10_' sorts before 'layer1_' ('0' < '_'): the reference quirk a drop-in must reproduce
    firsts = [n.split("_")[0] for n in names]
    assert firsts.index("layer10") < firsts.index("layer1")


# ---------------------------------------------------------------- diagonal energies (JSSP goldens)
@pytest.mark.parametrize("key", ["jssp_4q", "jssp_5q", "jssp_8q", "jssp_12q", "jssp_unit_test"])
def test_diagonal_table_matches_reference_encoder(jssp_golden, key):
    entry = jssp_golden[key]
    n = entry["n_qubits"]
    terms = list(zip(entry["z_masks"], entry["coeffs"]))
    table = oq.diagonal_table(n, terms)
    for bitstring, value in entry["lowest"]:
        assert table[int(bitstring, 2)] == pytest.approx(value, rel=1e-12, abs=1e-9)
    order = np.lexsort((np.arange(table.size), table))
    assert format(int(order[0]), f"0{n}b") == entry["lowest"][0][0]
    assert table.sum() == pytest.approx(entry["energy_sum"], rel=1e-12, abs=1e-6)
    if entry.get("energies"):
        np.testing.assert_allclose(table, entry["energies"], rtol=1e-12, atol=1e-9)


def test_notebook_minima(jssp_golden):
    # convergence values printed in the reference notebooks
    expect = {"jssp_4q": 63.5, "jssp_5q": 61.6, "jssp_8q": 22.75, "jssp_12q": 22.75}
    for key, value in expect.items():
        entry = jssp_golden[key]
        table = oq.diagonal_table(entry["n_qubits"], list(zip(entry["z_masks"], entry["coeffs"])))
        assert table.min() == pytest.approx(value, abs=1e-9)


def test_26q_probe_energies(jssp_golden):
    entry = jssp_golden["jssp_26q"]
    assert entry["n_qubits"] == 26 and entry["n_raw_terms"] == 346 and entry["n_distinct_terms"] == 84
    terms = list(zip(entry["z_masks"], entry["coeffs"]))
    for state, value in zip(entry["probe_states"], entry["probe_energies"]):
        assert oq.diagonal_energy(state, terms) == pytest.approx(value, rel=1e-12)


# ---------------------------------------------------------------- CVaR
def test_cvar_matches_reference(cvar_golden):
    for case in cvar_golden:
        states = [(i, p, v) for i, (p, v) in enumerate(zip(case["probs"], case["values"]))]
        assert oq.cvar_accumulate(states, case["alpha"]) == pytest.approx(case["expected"], rel=1e-14, abs=1e-14)


def test_product_cvar_matches_reference_and_its_vectorised_form(cvar_golden):
    """queasars_b200.expectation: the loop restatement of _get_expectation and the array version the sampler evaluator uses
    (no per-entry Python loop) against the fixtures produced by the reference's own function, then against each other on
    random distributions with ties, early isclose stops and clipped entries."""
    from queasars_b200 import expectation as ex

    for case in cvar_golden:
        want = case["expected"]
        assert ex.lower_tail_expectation(case["probs"], case["values"], case["alpha"]) == pytest.approx(want, rel=1e-14, abs=1e-14)
        assert ex.lower_tail_expectation_arrays(case["probs"], case["values"], case["alpha"]) == pytest.approx(want, rel=1e-13, abs=1e-13)
    rng = np.random.default_rng(0)
    for trial in range(400):
        n = int(rng.integers(1, 80))
        counts = rng.integers(1, 40, size=n)
        probs = counts / counts.sum()
        values = rng.integers(-4, 5, size=n).astype(float) if trial % 2 else rng.normal(size=n)
        alpha = float(rng.choice([1.0, 0.5, 0.25, 0.05, 0.999995, float(probs[: max(1, n // 2)].sum())]))
        a = ex.lower_tail_expectation(probs, values, alpha)
        b = ex.lower_tail_expectation_arrays(probs, values, alpha)
        assert b == pytest.approx(a, rel=1e-13, abs=1e-13)


# ---------------------------------------------------------------- analytic known answers for the simulator
def test_u_matrix_and_little_endian():
    # X on qubit 0 of 2 qubits -> |01> = index 1
    st = oq.statevector([("u", (0,), (math.pi, 0.0, math.pi))], 2)
    np.testing.assert_allclose(np.abs(st) ** 2, [0, 1, 0, 0], atol=1e-15)
    # cu3 with control (first qarg) = qubit 0 set -> flips qubit 1
    st = oq.statevector([("x", (0,), ()), ("cu3", (0, 1), (math.pi, 0.0, math.pi))], 2)
    np.testing.assert_allclose(np.abs(st) ** 2, [0, 0, 0, 1], atol=1e-15)
    st = oq.statevector([("cu3", (0, 1), (math.pi, 0.0, math.pi))], 2)
    np.testing.assert_allclose(np.abs(st) ** 2, [1, 0, 0, 0], atol=1e-15)


def test_single_qubit_expectations():
    theta, phi = 0.7, 1.3
    st = oq.statevector([("u", (0,), (theta, phi, 0.2))], 1)
    assert oq.pauli_expectation(st, "Z").real == pytest.approx(math.cos(theta), abs=1e-14)
    assert oq.pauli_expectation(st, "X").real == pytest.approx(math.sin(theta) * math.cos(phi), abs=1e-14)
    assert oq.pauli_expectation(st, "Y").real == pytest.approx(math.sin(theta) * math.sin(phi), abs=1e-14)


def test_bell_and_ghz():
    st = oq.statevector([("h", (0,), ()), ("cx", (0, 1), ()), ("cx", (1, 2), ())], 3)
    assert oq.estimator_expectation(st, [("ZZI", 1.0), ("IZZ", 1.0), ("XXX", 1.0), ("ZII", 1.0)]) == pytest.approx(3.0, abs=1e-14)
    assert oq.estimator_expectation(st, [("YYX", 1.0)]) == pytest.approx(-1.0, abs=1e-14)


def test_two_qubit_gate_identities():
    rng = np.random.default_rng(0)
    n = 3
    prep = [("u", (q,), tuple(rng.uniform(0, 6, 3))) for q in range(n)]
    th = 0.813
    a = oq.statevector(prep + [("rzz", (0, 2), (th,))], n)
    b = oq.statevector(prep + [("cx", (0, 2), ()), ("rz", (2,), (th,)), ("cx", (0, 2), ())], n)
    np.testing.assert_allclose(a, b, atol=1e-14)
    a = oq.statevector(prep + [("rzx", (0, 1), (th,))], n)
    b = oq.statevector(prep + [("h", (1,), ()), ("rzz", (0, 1), (th,)), ("h", (1,), ())], n)
    np.testing.assert_allclose(a, b, atol=1e-14)
    a = oq.statevector(prep + [("ecr", (0, 1), ())], n)
    b = oq.statevector(prep + [("rzx", (0, 1), (math.pi / 4,)), ("x", (0,), ()), ("rzx", (0, 1), (-math.pi / 4,))], n)
    np.testing.assert_allclose(a, b, atol=1e-14)
    a = oq.statevector(prep + [("swap", (0, 2), ())], n)
    b = oq.statevector(prep + [("cx", (0, 2), ()), ("cx", (2, 0), ()), ("cx", (0, 2), ())], n)
    np.testing.assert_allclose(a, b, atol=1e-14)


def test_sampler_semantics():
    st = oq.statevector([("h", (0,), ()), ("cx", (0, 1), ())], 2)
    idx = oq.sample_indices(st, 1000, seed=5)
    assert set(np.unique(idx)) <= {0, 3}
    counts = oq.counts_from_indices(idx, 2)
    assert sum(counts.values()) == 1000 and set(counts) <= {"00", "11"}
    # identical uniforms -> identical indices, and they follow searchsorted(side='right')
    u = np.array([0.0, 0.49999, 0.5, 0.99999])
    np.testing.assert_array_equal(oq.sample_indices(st, 4, uniforms=u), [0, 0, 3, 3])


def test_test_model_hamiltonian_ground_state():
    # min x^2 - y^2, x,y in [0,3] -> H = -1.5 Z0 - 3 Z1 + Z0Z1 + 1.5 Z2 + 3 Z3 - Z2Z3
    terms = [("IIIZ", -1.5), ("IIZI", -3.0), ("IIZZ", 1.0), ("IZII", 1.5), ("ZIII", 3.0), ("ZZII", -1.0)]
    table = oq.diagonal_table(4, oq.diag_terms_from_labels(terms))
    assert format(int(np.argmin(table)), "04b") == "1100"
    assert table.min() == pytest.approx(-9.0)
