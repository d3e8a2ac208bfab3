"""The primitives' and evaluators' HOST logic end to end on CPU, with the oracle-backed ``FakeEngine`` in place of the CUDA
engine (tests/fake_engine.py): caches, locks, device split, coalescing queue, pub containers, phase-deferred plans.  The same
flows run against the real engine in the ``-m gpu`` tests; this file exists so that a host-side deadlock or ordering bug is
caught here, not on the GPU box."""
import threading

import numpy as np
import pytest

from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq
from queasars_b200 import primitives as pr
from queasars_b200.operators import SparsePauliOp
from tests.fake_engine import FakeEngine
from tests.test_frontend_planner import build_circuit


@pytest.fixture()
def fake_engines(monkeypatch):
    made = {}

    def get_engine(device=0, dtype="complex128"):
        return made.setdefault((device, dtype), FakeEngine(device, dtype))

    monkeypatch.setattr(pr, "get_engine", get_engine)
    monkeypatch.setattr(pr._native, "device_count", lambda: 3)
    return made


def case(n, layers, seed):
    genome, values = og.random_individual(n, layers, True, seed)
    instr = og.individual_circuit(genome, values)
    return instr, list(values), build_circuit(instr, n)


def terms_for(n):
    rng = np.random.default_rng(n)
    out = []
    for q in range(n):
        lab = ["I"] * n
        lab[n - 1 - q] = "Z"
        out.append(("".join(lab), float(rng.normal())))
        lab[n - 1 - (q + 1) % n] = "Z"
        out.append(("".join(lab), float(rng.normal())))
    return out


@pytest.mark.timeout(120)
@pytest.mark.parametrize("devices,coalesce", [(None, True), ("all", True), (None, False), ("all", False)])
def test_operator_evaluator_host_flow(fake_engines, devices, coalesce):
    """(coalesce=False on one device is the direct engine call of the optimizer loop; every other combination goes through
    the queue / the device split)"""
    from queasars_b200.evaluators import B200OperatorCircuitEvaluator

    n = 5
    terms = terms_for(n)
    op = SparsePauliOp.from_list(terms)
    est = pr.B200EstimatorV2(devices=devices, coalesce=coalesce)
    ev = B200OperatorCircuitEvaluator(est, 0.0, op)
    cases = [case(n, 2, s) for s in range(7)]
    got = ev.evaluate_circuits([c for _, _, c in cases], [v for _, v, _ in cases])
    for g, (instr, values, _) in zip(got, cases):
        want = oq.estimator_expectation(oq.statevector(instr, n, values), terms)
        assert abs(g - want) < 1e-10
    if devices == "all":
        assert len(fake_engines) == 3 and all(e.launch_count > 0 for e in fake_engines.values())
    # non-diagonal operator: final phases must be kept
    xterms = terms + [("X" * n, 0.7), ("IYIZI"[:n], -0.3)]
    ev2 = B200OperatorCircuitEvaluator(est, 0.0, SparsePauliOp.from_list(xterms))
    instr, values, circ = cases[0]
    assert abs(ev2.evaluate_circuits([circ], [values])[0] - oq.estimator_expectation(oq.statevector(instr, n, values), xterms)) < 1e-10
    # threads around one primitive
    results = [None] * len(cases)

    def work(i):
        for _ in range(3):
            results[i] = ev.evaluate_circuits([cases[i][2]], [cases[i][1]])[0]

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    [t.start() for t in threads]
    [t.join(60) for t in threads]
    assert not any(t.is_alive() for t in threads)
    np.testing.assert_allclose(results, got, atol=1e-12)
    with pytest.raises(ValueError):
        ev.evaluate_circuits([cases[0][2]], [cases[0][1][:-1]])


@pytest.mark.timeout(120)
def test_sampler_evaluators_and_pub_contract_host_flow(fake_engines):
    from queasars_b200.evaluators import B200OperatorSamplerCircuitEvaluator, measure_quasi_distributions

    n, shots = 4, 300
    terms = terms_for(n)
    op = SparsePauliOp.from_list(terms)
    smp = pr.B200SamplerV2(devices="all", seed=9)
    instr, values, circ = case(n, 2, 3)
    dist = measure_quasi_distributions([circ, circ], [values, values], smp, shots)
    want_idx = oq.sample_indices(oq.statevector(instr, n, values), shots, seed=9)
    want = oq.quasi_distribution(oq.counts_from_indices(want_idx, n), shots)
    assert dict(dist[0]) == want and dict(dist[1]) == want
    for alpha in (1.0, 0.4):
        got = B200OperatorSamplerCircuitEvaluator(smp, shots, op, alpha=alpha).evaluate_circuits([circ], [values])[0]
        assert got == pytest.approx(oq.expectation_with_operator(want, oq.diag_terms_from_labels(terms), alpha), rel=1e-10, abs=1e-10)
    res = smp.run(pubs=[(circ.measure_all(inplace=False), values)], shots=shots).result()
    assert res[0].data["meas"].get_counts() == oq.counts_from_indices(want_idx, n)
    est = pr.B200EstimatorV2(devices="all", seed=1)
    res = est.run(pubs=[(circ, op, values)], precision=0.0).result()
    assert abs(float(res[0].data.evs) - oq.estimator_expectation(oq.statevector(instr, n, values), terms)) < 1e-10


@pytest.mark.timeout(120)
def test_individuals_are_accepted_in_place_of_circuits(fake_engines):
    """the evaluators take EVQE individuals directly (genome -> gate list, no QuantumCircuit) and give the
    values of the circuit the reference would have built from them (individual.py:288-322)."""
    from queasars_b200 import genome as gn
    from queasars_b200.evaluators import B200OperatorCircuitEvaluator, B200OperatorSamplerCircuitEvaluator

    n = 5
    op = SparsePauliOp.from_list(terms_for(n))
    pop = gn.random_population(n, 3, 4, True, 2)
    values = [list(i.parameter_values) for i in pop]
    ev = B200OperatorCircuitEvaluator(pr.B200EstimatorV2(), 0.0, op)
    direct = ev.evaluate_circuits(pop, values)
    via_circuit = ev.evaluate_circuits([i.to_circuit() for i in pop], values)
    np.testing.assert_allclose(direct, via_circuit, atol=1e-12)
    smp = B200OperatorSamplerCircuitEvaluator(pr.B200SamplerV2(seed=3), 200, op, alpha=0.5)
    np.testing.assert_allclose(smp.evaluate_circuits(pop, values), smp.evaluate_circuits([i.to_circuit() for i in pop], values), atol=1e-12)
