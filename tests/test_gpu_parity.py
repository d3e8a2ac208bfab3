"""GPU parity tests proper: the CUDA path (through the C-ABI via ctypes) against the CPU oracle on the same
seeded inputs, against the committed golden fixtures, and -- at BASELINE.json's full sizes -- through
size-independent properties (norm, analytic product states, linearity of the expectation).

Tolerances (BASELINE.json north_star): expectation values 1e-10 relative in complex128, 1e-4 in complex64;
sampling: identical indices for identical uniforms up to <= 1 boundary flip per 10^4 shots, plus a
total-variation bound against the exact distribution.
"""
import math
import threading

import numpy as np
import pytest

from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq
from queasars_b200 import gate_list as gl
from queasars_b200.circuit import Parameter, QuantumCircuit
from queasars_b200.operators import SparsePauliOp
from tests.test_frontend_planner import build_circuit

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from queasars_b200.engine import Engine

    return Engine(device=0, dtype="complex128")


@pytest.fixture(scope="module")
def engine32():
    from queasars_b200.engine import Engine

    return Engine(device=0, dtype="complex64")


def rel_err(got, want):
    return abs(got - want) / max(1.0, abs(want))


def evqe_case(n, layers, seed, parameterized=None):
    genome, values = og.random_individual(n, layers, True, seed)
    instr = og.individual_circuit(genome, values, parameterized)
    if parameterized is not None:
        (lid,) = parameterized
        values = values[og.layer_value_slice(genome, lid)]
    return instr, list(values), build_circuit(instr, n)


def random_ising(n, seed=1234):
    rng = np.random.default_rng(seed)
    terms = []
    for i in range(n):
        lab = ["I"] * n
        lab[n - 1 - i] = "Z"
        terms.append(("".join(lab), float(rng.normal())))
    for i in range(n):
        for j in range(i + 1, n):
            lab = ["I"] * n
            lab[n - 1 - i] = lab[n - 1 - j] = "Z"
            terms.append(("".join(lab), float(rng.normal())))
    return terms


def tfim(n, h=0.5):
    terms = []
    for i in range(n - 1):
        lab = ["I"] * n
        lab[n - 1 - i] = lab[n - 2 - i] = "Z"
        terms.append(("".join(lab), -1.0))
    for i in range(n):
        lab = ["I"] * n
        lab[n - 1 - i] = "X"
        terms.append(("".join(lab), -h))
    return terms


# ------------------------------------------------------------------------------------ statevectors
@pytest.mark.parametrize("n,layers,seed", [(1, 1, 0), (2, 2, 1), (4, 2, 0), (5, 3, 2), (8, 4, 3), (12, 3, 4), (13, 3, 5), (16, 4, 6), (20, 3, 7)])
def test_statevector_matches_oracle(engine, n, layers, seed):
    instr, values, circ = evqe_case(n, layers, seed)
    plan = engine.compile(gl.from_circuit(circ))
    got = engine.statevector(plan, values)
    want = oq.statevector(instr, n, values)
    assert np.max(np.abs(got - want)) < 1e-13
    assert abs(np.vdot(got, got).real - 1.0) < 1e-13


def test_partially_parameterised_circuit(engine):
    for lid in (0, 1, -1):
        instr, values, circ = evqe_case(11, 3, 21, {lid})
        plan = engine.compile(gl.from_circuit(circ))
        assert np.max(np.abs(engine.statevector(plan, values) - oq.statevector(instr, 11, values))) < 1e-13


def test_transpiled_basis_gate_set(engine):
    n = 6
    a, b = Parameter("a"), Parameter("b")
    circ = QuantumCircuit(n)
    circ.h(0), circ.sx(1), circ.rz(a, 1), circ.cx(0, 1), circ.rzz(b * 2 + 0.3, 1, 3), circ.ecr(2, 4), circ.p(-a, 2)
    circ.cz(2, 3), circ.swap(0, 5), circ.ry(0.4, 3), circ.rx(b, 0), circ.t(1), circ.sdg(2), circ.cp(0.7, 4, 1)
    circ.crz(a * 0.5, 3, 0), circ.cu(0.1, 0.2, 0.3, 0.4, 1, 2), circ.rzx(0.9, 0, 2), circ.rxx(a, 3, 4), circ.y(5), circ.z(1)
    instr = []
    for inst in circ.data:
        ps = []
        for p in inst.operation.params:
            if hasattr(p, "parameters") and p.parameters:
                (prm,) = p.parameters
                ps.append((p._terms[prm], prm.name, p._const))
            else:
                ps.append(float(p))
        instr.append((inst.operation.name, tuple(q._index for q in inst.qubits), tuple(ps)))
    values = [0.37, -1.2]
    plan = engine.compile(gl.from_circuit(circ))
    assert np.max(np.abs(engine.statevector(plan, values) - oq.statevector(instr, n, values))) < 1e-13


def test_empty_circuit_and_identity(engine):
    circ = QuantumCircuit(3)
    circ.id(0)
    plan = engine.compile(gl.from_circuit(circ))
    got = engine.statevector(plan, [])
    want = np.zeros(8, dtype=complex)
    want[0] = 1
    np.testing.assert_array_equal(got, want)


# ------------------------------------------------------------------------------------ expectation values
@pytest.mark.parametrize("key", ["jssp_4q", "jssp_5q", "jssp_8q", "jssp_12q", "jssp_unit_test"])
def test_jssp_diagonal_expectation(engine, jssp_golden, key):
    entry = jssp_golden[key]
    n = entry["n_qubits"]
    terms = list(zip(entry["z_masks"], entry["coeffs"]))
    op = SparsePauliOp._raw(n, [0] * len(terms), entry["z_masks"], entry["coeffs"])
    for build_table in (True, False):
        ham = engine.hamiltonian(op, build_table=build_table)
        plans, params, want = [], [], []
        for seed in range(6):
            instr, values, circ = evqe_case(n, 2 + seed % 3, 100 + seed)
            plans.append(engine.compile(gl.from_circuit(circ)))
            params.append(values)
            st = oq.statevector(instr, n, values)
            want.append(float(np.dot(np.abs(st) ** 2, oq.diagonal_table(n, terms))))
        got = engine.expectation(plans, params, ham)
        for g, w in zip(got, want):
            assert rel_err(g, w) < 1e-10
    # known answer: the golden minimum-energy basis state (notebook convergence values)
    best = entry["lowest"][0]
    circ = QuantumCircuit(n)
    for q in range(n):
        if best[0][n - 1 - q] == "1":
            circ.x(q)
        else:
            circ.id(q)
    plan = engine.compile(gl.from_circuit(circ))
    assert engine.expectation([plan], [[]], engine.hamiltonian(op))[0] == pytest.approx(best[1], rel=1e-12)


def test_random_ising_16q_batch(engine):
    n = 16
    terms = random_ising(n)
    ham = engine.hamiltonian(SparsePauliOp.from_list(terms))
    table = oq.diagonal_table(n, oq.diag_terms_from_labels(terms))
    plans, params, want = [], [], []
    for seed in range(5):  # heterogeneous depths -> different sweep counts inside one batch
        instr, values, circ = evqe_case(n, 1 + seed, 40 + seed)
        plans.append(engine.compile(gl.from_circuit(circ)))
        params.append(values)
        want.append(float(np.dot(np.abs(oq.statevector(instr, n, values)) ** 2, table)))
    got = engine.expectation(plans, params, ham)
    for g, w in zip(got, want):
        assert rel_err(g, w) < 1e-10
    single = [engine.expectation([p], [v], ham)[0] for p, v in zip(plans, params)]
    np.testing.assert_allclose(got, single, rtol=0, atol=1e-12)


@pytest.mark.parametrize("n", [3, 10, 14])
def test_pauli_sum_expectation(engine, n):
    terms = tfim(n)
    rng = np.random.default_rng(n)
    for _ in range(3):  # add a few random mixed X/Y/Z strings with complex coefficients
        lab = "".join(rng.choice(list("IXYZ"), size=n))
        terms.append((lab, complex(rng.normal(), rng.normal())))
    ham = engine.hamiltonian(SparsePauliOp.from_list(terms))
    for seed in range(3):
        instr, values, circ = evqe_case(n, 3, 70 + seed)
        plan = engine.compile(gl.from_circuit(circ))
        want = oq.estimator_expectation(oq.statevector(instr, n, values), terms)
        assert rel_err(engine.expectation([plan], [values], ham)[0], want) < 1e-10


def test_complex64_tolerance(engine32):
    n = 14
    terms = random_ising(n, 5) + tfim(n)
    ham = engine32.hamiltonian(SparsePauliOp.from_list(terms))
    for seed in range(3):
        instr, values, circ = evqe_case(n, 4, 90 + seed)
        plan = engine32.compile(gl.from_circuit(circ))
        want = oq.estimator_expectation(oq.statevector(instr, n, values), terms)
        assert rel_err(engine32.expectation([plan], [values], ham)[0], want) < 1e-4
        assert np.max(np.abs(engine32.statevector(plan, values) - oq.statevector(instr, n, values))) < 1e-5


# ------------------------------------------------------------------------------------ sampling
@pytest.mark.parametrize("n,shots", [(4, 1000), (10, 10000), (13, 10000), (16, 4096)])
def test_sampler_indices_match_searchsorted(engine, n, shots):
    instr, values, circ = evqe_case(n, 3, 7 * n)
    plan = engine.compile(gl.from_circuit(circ))
    uniforms = np.random.default_rng(99).random((2, shots))
    got = engine.sample([plan, plan], [values, values], shots, uniforms)
    state = oq.statevector(instr, n, values)
    for row in range(2):
        want = oq.sample_indices(state, shots, uniforms=uniforms[row])
        mismatches = np.nonzero(got[row] != want)[0]
        assert len(mismatches) <= max(1, shots // 10000)
        for i in mismatches:  # only a flip to an adjacent state at a CDF boundary is tolerated
            assert abs(int(got[row][i]) - int(want[i])) <= 1 or abs(state[got[row][i]]) ** 2 > 0


def test_sampler_total_variation_bound(engine):
    n, shots = 10, 10000
    instr, values, circ = evqe_case(n, 4, 3)
    plan = engine.compile(gl.from_circuit(circ))
    probs = np.abs(oq.statevector(instr, n, values)) ** 2
    idx = engine.sample([plan], [values], shots, np.random.default_rng(1).random((1, shots)))[0]
    emp = np.bincount(idx, minlength=1 << n) / shots
    tv = 0.5 * np.abs(emp - probs).sum()
    bound = 1.5 * 0.5 * np.sum(np.sqrt(2 * probs * (1 - probs) / (math.pi * shots)))
    assert tv <= bound


# ------------------------------------------------------------------------------------ evaluators (drop-in boundary)
def test_operator_evaluator_matches_oracle(jssp_golden):
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator

    entry = jssp_golden["jssp_8q"]
    n = entry["n_qubits"]
    op = SparsePauliOp._raw(n, [0] * entry["n_raw_terms"], entry["z_masks"], entry["coeffs"])
    table = oq.diagonal_table(n, list(zip(entry["z_masks"], entry["coeffs"])))
    ev = B200OperatorCircuitEvaluator(B200EstimatorV2(seed=0), 0.0, op)
    assert ev.n_qubits == 8
    circuits, params, want = [], [], []
    for seed in range(10):
        instr, values, circ = evqe_case(n, 2, seed)
        circuits.append(circ)
        params.append(values)
        want.append(float(np.dot(np.abs(oq.statevector(instr, n, values)) ** 2, table)))
    got = ev.evaluate_circuits(circuits, params)
    assert isinstance(got, list) and len(got) == 10
    for g, w in zip(got, want):
        assert rel_err(g, w) < 1e-10
    # same circuit object, many parameter vectors (the optimizer's calling pattern, mutation.py:63-75)
    instr, values, circ = evqe_case(n, 2, 0, {-1})
    rng = np.random.default_rng(0)
    batch = [list(rng.uniform(0, 2 * math.pi, len(values))) for _ in range(7)]
    got = ev.evaluate_circuits([circ] * 7, batch)
    for g, v in zip(got, batch):
        assert rel_err(g, float(np.dot(np.abs(oq.statevector(instr, n, v)) ** 2, table))) < 1e-10
    # precision > 0 adds N(0, precision) noise from default_rng(seed)
    noisy = B200OperatorCircuitEvaluator(B200EstimatorV2(seed=5), 0.05, op).evaluate_circuits([circ], [batch[0]])[0]
    assert noisy == pytest.approx(float(np.random.default_rng(5).normal(got[0], 0.05)), abs=1e-9)


def test_initial_state_circuit(engine):
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator

    n = 5
    init = QuantumCircuit(n)
    for q in range(n):
        init.h(q)
    op = SparsePauliOp.from_list(tfim(n))
    ev = B200OperatorCircuitEvaluator(B200EstimatorV2(), 0.0, op, initial_state_circuit=init)
    instr, values, circ = evqe_case(n, 2, 4)
    full = [("h", (q,), ()) for q in range(n)] + instr
    want = oq.estimator_expectation(oq.statevector(full, n, values), tfim(n))
    assert rel_err(ev.evaluate_circuits([circ], [values])[0], want) < 1e-10
    assert len(circ.data) == len(instr)  # inputs are not mutated
    with pytest.raises(ValueError):
        B200OperatorCircuitEvaluator(B200EstimatorV2(), 0.0, op, initial_state_circuit=QuantumCircuit(n + 1))


@pytest.mark.parametrize("alpha", [1.0, 0.5, 0.1])
def test_sampler_evaluators_match_oracle(jssp_golden, alpha):
    from queasars_b200 import (
        B200BitstringCircuitEvaluator,
        B200OperatorSamplerCircuitEvaluator,
        B200SamplerV2,
        BitstringEvaluator,
        DiagonalEnergyBitstringEvaluator,
    )

    entry = jssp_golden["jssp_5q"]
    n, shots, seed = entry["n_qubits"], 512, 11
    terms = list(zip(entry["z_masks"], entry["coeffs"]))
    op = SparsePauliOp._raw(n, [0] * len(terms), entry["z_masks"], entry["coeffs"])
    sampler = B200SamplerV2(seed=seed)
    ev_op = B200OperatorSamplerCircuitEvaluator(sampler, shots, op, alpha=alpha)
    fn = lambda bits: oq.diagonal_energy(int(bits, 2), terms)  # noqa: E731
    ev_bs = B200BitstringCircuitEvaluator(sampler, shots, BitstringEvaluator(n, fn), alpha=alpha)
    # the same energy as a vectorised evaluator: all distinct sampled states valued in one device call
    ev_vec = B200BitstringCircuitEvaluator(sampler, shots, DiagonalEnergyBitstringEvaluator(n, entry["z_masks"], entry["coeffs"]), alpha=alpha)
    for s in range(4):
        instr, values, circ = evqe_case(n, 2, 30 + s)
        state = oq.statevector(instr, n, values)
        idx = oq.sample_indices(state, shots, seed=seed)
        dist = oq.quasi_distribution(oq.counts_from_indices(idx, n), shots)
        want_op = oq.expectation_with_operator(dist, terms, alpha)
        want_bs = oq.expectation_with_bitstring_function(dist, n, fn, alpha)
        assert ev_op.evaluate_circuits([circ], [values])[0] == pytest.approx(want_op, rel=1e-10, abs=1e-10)
        assert ev_bs.evaluate_circuits([circ], [values])[0] == pytest.approx(want_bs, rel=1e-10, abs=1e-10)
        assert ev_vec.evaluate_circuits([circ], [values])[0] == pytest.approx(want_bs, rel=1e-10, abs=1e-10)
    with pytest.raises(ValueError):
        B200OperatorSamplerCircuitEvaluator(sampler, shots, op, alpha=0.0)
    with pytest.raises(ValueError):
        B200BitstringCircuitEvaluator(sampler, shots, BitstringEvaluator(n, fn), alpha=1.5)
    with pytest.raises(ValueError):
        B200OperatorSamplerCircuitEvaluator(sampler, shots, "not an operator")


def test_primitive_contract_objects(jssp_golden):
    """run(pubs=..., precision=/shots=).result()[i].data.evs / .data['meas'].get_counts(), tuple and Pub inputs,
    generator of pubs (transpiling_primitives.py:82-83), metadata attribute (mutex_primitives.py:259)."""
    from queasars_b200 import B200EstimatorV2, B200SamplerV2
    from queasars_b200.containers import EstimatorPub, SamplerPub

    n = 4
    op = SparsePauliOp.from_list(tfim(n))
    instr, values, circ = evqe_case(n, 2, 0)
    want = oq.estimator_expectation(oq.statevector(instr, n, values), tfim(n))
    est = B200EstimatorV2()
    res = est.run(pubs=((circ, op, values), EstimatorPub(circ, op, values, None)), precision=0.0).result()
    assert len(res) == 2 and hasattr(res, "metadata")
    for r in res:
        assert np.shape(r.data.evs) == () and rel_err(float(np.real(r.data.evs)), want) < 1e-10
    res = est.run((pub for pub in [(circ, op, values)]), precision=None).result()
    assert rel_err(float(res[0].data.evs), want) < 1e-10
    smp = B200SamplerV2(seed=3)
    measured = circ.measure_all(inplace=False)
    res = smp.run(pubs=((measured, values), SamplerPub(measured, values, None)), shots=256).result()
    counts = res[0].data["meas"].get_counts()
    assert sum(counts.values()) == 256 and all(len(k) == n for k in counts)
    assert counts == res[1].data["meas"].get_counts()  # fresh default_rng(seed) per pub
    idx = oq.sample_indices(oq.statevector(instr, n, values), 256, seed=3)
    assert counts == oq.counts_from_indices(idx, n)
    with pytest.raises(Exception):
        est.run(pubs=((circ, op, values[:-1]),), precision=0.0).result()


def test_concurrent_threads_coalesce(jssp_golden):
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator

    entry = jssp_golden["jssp_12q"]
    n = entry["n_qubits"]
    op = SparsePauliOp._raw(n, [0] * entry["n_raw_terms"], entry["z_masks"], entry["coeffs"])
    ev = B200OperatorCircuitEvaluator(B200EstimatorV2(), 0.0, op)
    cases = [evqe_case(n, 2 + s % 2, 200 + s) for s in range(16)]
    sequential = [ev.evaluate_circuits([c], [v])[0] for _, v, c in cases]
    results = [None] * len(cases)

    def work(i):
        for _ in range(5):
            results[i] = ev.evaluate_circuits([cases[i][2]], [cases[i][1]])[0]

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    np.testing.assert_allclose(results, sequential, rtol=0, atol=1e-12)


# ------------------------------------------------------------------------------------ full-size properties
def product_state_circuit(n, thetas):
    circ = QuantumCircuit(n)
    for q in range(n):
        circ.u(float(thetas[q]), 0.3 * q, 0.1, q)
    return circ


@pytest.mark.parametrize("n", [20, 24, 26])
def test_full_size_properties(engine, n):
    rng = np.random.default_rng(n)
    # (1) analytic product state: <Z_i> = cos(theta_i), <Z_i Z_j> = cos*cos, <X_i> = sin(theta_i) cos(phi_i)
    thetas = rng.uniform(0, math.pi, n)
    plan = engine.compile(gl.from_circuit(product_state_circuit(n, thetas)))
    zz_terms, want = [], 0.0
    for i in range(0, n, 3):
        j = (i + 5) % n
        lab = ["I"] * n
        lab[n - 1 - i] = "Z"
        c1 = float(rng.normal())
        zz_terms.append(("".join(lab), c1))
        want += c1 * math.cos(thetas[i])
        lab[n - 1 - j] = "Z"
        c2 = float(rng.normal())
        zz_terms.append(("".join(lab), c2))
        want += c2 * math.cos(thetas[i]) * math.cos(thetas[j])
    for build_table in (True, False):
        ham = engine.hamiltonian(SparsePauliOp.from_list(zz_terms), build_table=build_table)
        assert rel_err(engine.expectation([plan], [[]], ham)[0], want) < 1e-10
    lab = ["I"] * n
    lab[n - 1 - (n - 2)] = "X"
    hx = engine.hamiltonian(SparsePauliOp.from_list([("".join(lab), 1.0)]))
    assert rel_err(engine.expectation([plan], [[]], hx)[0], math.sin(thetas[n - 2]) * math.cos(0.3 * (n - 2))) < 1e-10
    # (2) entangling EVQE circuit: norm is preserved, and the expectation is linear in H
    instr, values, circ = evqe_case(n, 3, n)
    plan = engine.compile(gl.from_circuit(circ))
    ident = engine.hamiltonian(SparsePauliOp.from_list([("I" * n, 1.0)]), build_table=False)
    assert abs(engine.expectation([plan], [values], ident)[0] - 1.0) < 1e-11
    h1 = SparsePauliOp.from_list(zz_terms)
    h2 = SparsePauliOp.from_list(tfim(n))
    e1 = engine.expectation([plan], [values], engine.hamiltonian(h1))[0]
    e2 = engine.expectation([plan], [values], engine.hamiltonian(h2))[0]
    e12 = engine.expectation([plan], [values], engine.hamiltonian(h1 * 0.7 + h2 * (-1.3)))[0]
    assert rel_err(e12, 0.7 * e1 - 1.3 * e2) < 1e-10


def test_26q_jssp_sampler_route(engine, jssp_golden):
    """C4 shape: 26-qubit JSSP QUBO, 10k shots; checked through exact sampled-energy identities."""
    entry = jssp_golden["jssp_26q"]
    n, shots = 26, 10000
    op = SparsePauliOp._raw(n, [0] * entry["n_raw_terms"], entry["z_masks"], entry["coeffs"])
    ham = engine.hamiltonian(op)
    np.testing.assert_allclose(
        engine.diag_energies(ham, np.array(entry["probe_states"], dtype=np.uint64)), entry["probe_energies"], rtol=1e-12
    )
    # basis state circuit -> every shot returns that state
    target = entry["probe_states"][0]
    circ = QuantumCircuit(n)
    for q in range(n):
        if (target >> q) & 1:
            circ.x(q)
    plan = engine.compile(gl.from_circuit(circ))
    idx = engine.sample([plan], [[]], shots, np.random.default_rng(0).random((1, shots)))[0]
    assert np.all(idx == target)
    # product state: sampled mean energy within 5 sigma / sqrt(S) of the exact expectation
    thetas = np.random.default_rng(3).uniform(0, math.pi, n)
    plan = engine.compile(gl.from_circuit(product_state_circuit(n, thetas)))
    exact = engine.expectation([plan], [[]], ham)[0]
    idx = engine.sample([plan], [[]], shots, np.random.default_rng(4).random((1, shots)))[0]
    energies = engine.diag_energies(ham, idx.astype(np.uint64))
    assert abs(energies.mean() - exact) < 5 * energies.std() / math.sqrt(shots)
    marg = np.array([np.mean((idx >> q) & 1) for q in range(n)])
    assert np.max(np.abs(marg - np.sin(thetas / 2) ** 2)) < 5 * 0.5 / math.sqrt(shots)


# ------------------------------------------------------------------------------------ device-pointer entry points
def test_sharded_api_single_rank(engine):
    """ShardedStatevector with one rank drives qb_apply_plan_device / qb_expectation_device on a torch-owned buffer."""
    import torch

    from queasars_b200.sharded import CudaShardBackend, ShardedStatevector

    n = 14
    instr, values, circ = evqe_case(n, 3, 8)
    sv = ShardedStatevector(n, backend=CudaShardBackend(engine), device=torch.device("cuda", 0))
    sv.run(gl.from_circuit(circ), values)
    want = oq.statevector(instr, n, values)
    assert np.max(np.abs(sv.gather_logical() - want)) < 1e-13
    rng = np.random.default_rng(2)
    z = [int(v) for v in rng.integers(0, 1 << n, size=7)]
    c = [float(v) for v in rng.normal(size=7)]
    e_want = float(np.dot(np.abs(want) ** 2, oq.diagonal_table(n, list(zip(z, c)))))
    assert rel_err(sv.diagonal_expectation(z, c), e_want) < 1e-10
    assert abs(sv.norm_squared() - 1.0) < 1e-12
    # general Pauli sum and sampling through the device-pointer entry points (qb_expectation_device / qb_sample_device)
    terms = tfim(n) + [("Y" + "I" * (n - 2) + "X", 0.3)]
    assert rel_err(sv.expectation(SparsePauliOp.from_list(terms)), oq.estimator_expectation(want, terms)) < 1e-10
    uniforms = np.random.default_rng(4).random(5000)
    drawn = sv.sample(len(uniforms), uniforms=uniforms)
    expect = oq.sample_indices(want, len(uniforms), uniforms=uniforms)  # one rank, identity permutation: same enumeration
    assert np.count_nonzero(drawn != expect) <= 1


def test_pipelined_submission_matches_single_call(engine):
    """Large lists of Python float lists are handed to the engine in two chunks (qb_evaluate_expectation_submit / _collect)
    so that the second chunk's value conversion overlaps the first chunk's GPU work: same values, same order, and a bad
    parameter vector in the second chunk still raises cleanly."""
    was = engine._pipeline
    engine._pipeline = True  # (off by default when the C marshalling helper is present: the submit / collect path stays tested)
    try:
        _pipelined_submission_case(engine)
    finally:
        engine._pipeline = was


def _pipelined_submission_case(engine):
    n, count = 13, 24
    terms = random_ising(n, 9)
    ham = engine.hamiltonian(SparsePauliOp.from_list(terms))
    plans, params, want = [], [], []
    for seed in range(count):
        instr, values, circ = evqe_case(n, 4 + seed % 2, 300 + seed)
        plans.append(engine.compile(gl.from_circuit(circ)))
        params.append([float(v) for v in values])
        if seed % 6 == 0:
            want.append((seed, oq.estimator_expectation(oq.statevector(instr, n, values), terms)))
    assert sum(p.n_params for p in plans) >= 2048 and engine._pipeline_split(plans, params) is not None
    piped = engine.expectation(plans, params, ham)
    single = engine.expectation(plans, [np.asarray(p) for p in params], ham)  # arrays: submitted in one piece
    assert engine._pipeline_split(plans, [np.asarray(p) for p in params]) is None
    assert np.array_equal(piped, single)
    for i, w in want:
        assert rel_err(piped[i], w) < 1e-10
    bad = list(params)
    bad[-1] = bad[-1][:-1]
    with pytest.raises(ValueError):
        engine.expectation(plans, bad, ham)
    assert np.array_equal(engine.expectation(plans, params, ham), single)  # nothing left pending after the failure


@pytest.mark.parametrize("n_local,lp", [(12, [11]), (13, [3, 12]), (14, [0, 5, 13]), (12, [2, 3, 4])])
def test_fused_swap_kernel_virtual_ranks(engine, n_local, lp):
    """qb_swap_global_p2p with all "ranks" on one GPU: every rank's kernel stores into the destination buffers of all ranks
    (here plain device buffers instead of peer-mapped ones); the result must be the bit permutation rank bit j <-> lp[j]."""
    import torch

    g = len(lp)
    world = 1 << g
    rng = np.random.default_rng(n_local + g)
    full = rng.normal(size=world << n_local) + 1j * rng.normal(size=world << n_local)
    src = [torch.from_numpy(full[r << n_local : (r + 1) << n_local].copy()).cuda() for r in range(world)]
    dst = [torch.zeros(1 << n_local, dtype=torch.complex128, device="cuda") for _ in range(world)]
    torch.cuda.synchronize()
    for r in range(world):
        engine.swap_global_p2p("complex128", n_local, src[r].data_ptr(), [d.data_ptr() for d in dst], r, lp)
    engine.synchronize()
    got = np.concatenate([d.cpu().numpy() for d in dst])
    idx = np.arange(world << n_local, dtype=np.int64)
    swapped = idx.copy()
    for j, p in enumerate(lp):
        a, b = (idx >> (n_local + j)) & 1, (idx >> p) & 1
        swapped &= ~((1 << (n_local + j)) | (1 << p))
        swapped |= (b << (n_local + j)) | (a << p)
    want = np.empty_like(full)
    want[swapped] = full
    assert np.array_equal(got, want)


def test_prefix_state_reuse_matches_full_evaluation(engine):
    """Optimizer pattern (mutation.py:57-81): one layer parameterised, the others bound numerically.  The cached
    prefix state must give the same expectation values as evaluating the whole circuit from |0...0>."""
    n = 14
    terms = random_ising(n, 9)
    ham = engine.hamiltonian(SparsePauliOp.from_list(terms))
    table = oq.diagonal_table(n, oq.diag_terms_from_labels(terms))
    rng = np.random.default_rng(4)
    for lid in (-1, 1):
        instr, values, circ = evqe_case(n, 4, 55, {lid})
        gates = gl.from_circuit(circ)
        reuse = engine.compile_with_prefix_reuse(gates)
        full = engine.compile(gates)
        if lid == -1:
            assert reuse.prefix is not None and reuse.n_ops < full.n_ops
        batch = [list(rng.uniform(0, 2 * math.pi, len(values))) for _ in range(5)]
        got = engine.expectation([reuse] * 5, batch, ham)
        ref = engine.expectation([full] * 5, batch, ham)
        np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
        for g, v in zip(got, batch):
            assert rel_err(g, float(np.dot(np.abs(oq.statevector(instr, n, v)) ** 2, table))) < 1e-10
        assert np.max(np.abs(engine.statevector(reuse, batch[0]) - oq.statevector(instr, n, batch[0]))) < 1e-13
        idx = engine.sample([reuse], [batch[0]], 2000, np.random.default_rng(1).random((1, 2000)))[0]
        want = oq.sample_indices(oq.statevector(instr, n, batch[0]), 2000, uniforms=np.random.default_rng(1).random(2000))
        assert np.count_nonzero(idx != want) <= 1


def test_term_counts_beyond_the_shared_memory_staging(engine):
    """Operators with more terms than one launch stages in shared memory (3 000 diagonal terms per table-builder launch, 2 000
    terms per x-mask group) and with duplicate strings (merged on the host): the reference accepts arbitrary SparsePauliOps."""
    n = 14
    rng = np.random.default_rng(77)
    zs = rng.choice(1 << n, size=5000, replace=False)
    diag = [(int(z), float(c)) for z, c in zip(zs, rng.normal(size=5000))]
    diag += diag[:300]  # duplicates: coefficients add up
    xmask = 0b101 << 5
    group = [(int(z), complex(a, b)) for z, a, b in zip(rng.choice(1 << n, size=2500, replace=False), rng.normal(size=2500), rng.normal(size=2500))]
    op_diag = SparsePauliOp._raw(n, [0] * len(diag), [z for z, _ in diag], [c for _, c in diag])
    op_full = SparsePauliOp._raw(n, [0] * len(diag) + [xmask] * len(group), [z for z, _ in diag] + [z for z, _ in group], [c for _, c in diag] + [c for _, c in group])
    instr, values, circ = evqe_case(n, 3, 5)
    plan = engine.compile(gl.from_circuit(circ))
    state = oq.statevector(instr, n, values)
    want_diag = float(np.dot(np.abs(state) ** 2, oq.diagonal_table(n, diag)))
    for build_table in (True, False):
        got = engine.expectation([plan, plan], [values, values], engine.hamiltonian(op_diag, build_table=build_table))
        assert rel_err(got[0], want_diag) < 1e-10 and got[0] == got[1]
    want_full = oq.estimator_expectation(state, op_full.to_list())
    assert rel_err(engine.expectation([plan], [values], engine.hamiltonian(op_full))[0], want_full) < 1e-10


def test_single_circuit_calls_replay_a_cuda_graph(engine):
    """Calls with one circuit at 1-4 parameter points (the optimizer loop: mutation.py:63-75) go through a cached CUDA graph per
    (plan, Hamiltonian, points): same values
    as the batched path for changing parameter values, for prefixed plans, and after the plan / Hamiltonian are destroyed and
    rebuilt (the cached graph must go with them)."""
    import gc

    n = 13
    terms = random_ising(n, 21)
    rng = np.random.default_rng(8)
    for lid in (None, -1):
        instr, values, circ = evqe_case(n, 4, 77, None if lid is None else {lid})
        gates = gl.from_circuit(circ)
        for round_ in range(2):  # second round: new plan + Hamiltonian objects after the first ones were released
            ham = engine.hamiltonian(SparsePauliOp.from_list(terms))
            plan = engine.compile_with_prefix_reuse(gates, drop_final_phases=True) if lid is not None else engine.compile(gates, cache=False)
            rows = [list(rng.uniform(0, 2 * math.pi, len(values))) for _ in range(6)]
            single = [engine.expectation([plan], [r], ham)[0] for r in rows]
            batched = engine.expectation([plan] * 6, rows, ham)
            np.testing.assert_allclose(single, batched, rtol=0, atol=1e-12)
            # one circuit at 2 / 3 / 4 parameter points (SPSA's theta +- c delta): graph path as well, rows given as lists and as a block
            for copies in (2, 3, 4):
                np.testing.assert_allclose(engine.expectation([plan] * copies, rows[:copies], ham), batched[:copies], rtol=0, atol=1e-12)
                np.testing.assert_allclose(engine.expectation([plan] * copies, np.asarray(rows[1 : copies + 1]), ham), batched[1 : copies + 1], rtol=0, atol=1e-12)
            with pytest.raises(Exception):
                engine.expectation([plan] * 2, [rows[0], rows[1][:-1]], ham)
            table = oq.diagonal_table(n, oq.diag_terms_from_labels(terms))
            assert rel_err(single[0], float(np.dot(np.abs(oq.statevector(instr, n, rows[0])) ** 2, table))) < 1e-10
            with pytest.raises(Exception):
                engine.expectation([plan], [rows[0][:-1]], ham)
            del plan, ham
            gc.collect()
