"""Multi-GPU behind the drop-in API: ONE ``evaluate_circuits`` call on a primitive configured with ``devices="all"`` is split
over every visible GPU (one native engine + one worker thread per device) and returns the values of the single-GPU path in
submission order.  On a one-GPU box the same code runs with a device set of one.  Also here: the cache guards against circuits
and operators edited in place after their first evaluation (the reference re-transpiles / re-binds on every call:
transpiling_primitives.py:47, circuit_evaluation.py:200-215)."""
import threading

import numpy as np
import pytest

from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq
from queasars_b200.operators import SparsePauliOp
from tests.test_frontend_planner import build_circuit
from tests.test_gpu_parity import random_ising, rel_err

pytestmark = pytest.mark.gpu


def evqe_case(n, layers, seed):
    genome, values = og.random_individual(n, layers, True, seed)
    instr = og.individual_circuit(genome, values)
    return instr, list(values), build_circuit(instr, n)


def test_one_call_is_split_over_all_visible_gpus():
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator, _native

    n, count = 16, 32
    n_dev = _native.device_count()
    terms = random_ising(n, 3)
    op = SparsePauliOp.from_list(terms)
    table = oq.diagonal_table(n, oq.diag_terms_from_labels(terms))
    cases = [evqe_case(n, 2 + s % 4, 500 + s) for s in range(count)]
    circuits, params = [c for _, _, c in cases], [v for _, v, _ in cases]
    multi = B200EstimatorV2(devices="all", coalesce=False)
    single = B200EstimatorV2(device=0, coalesce=False)
    assert multi.devices_used() == list(range(n_dev))
    before = [e.launch_count for e in multi.engines]
    got = B200OperatorCircuitEvaluator(multi, 0.0, op).evaluate_circuits(circuits, params)
    used = [e.launch_count - b for e, b in zip(multi.engines, before)]
    assert all(u > 0 for u in used), f"kernel launches per device: {used}"
    ref = B200OperatorCircuitEvaluator(single, 0.0, op).evaluate_circuits(circuits, params)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)
    for g, (instr, values, _) in zip(got[::5], cases[::5]):
        assert rel_err(g, float(np.dot(np.abs(oq.statevector(instr, n, values)) ** 2, table))) < 1e-10
    # one circuit, many parameter vectors (batched optimizer evaluation): the rows are dealt out over the devices
    rng = np.random.default_rng(1)
    rows = [list(rng.uniform(0, 6.28, len(params[0]))) for _ in range(24)]
    before = [e.launch_count for e in multi.engines]
    got = B200OperatorCircuitEvaluator(multi, 0.0, op).evaluate_circuits([circuits[0]] * 24, rows)
    assert all(e.launch_count > b for e, b in zip(multi.engines, before))
    ref = B200OperatorCircuitEvaluator(single, 0.0, op).evaluate_circuits([circuits[0]] * 24, rows)
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-12)


def test_sampler_on_all_devices_matches_single_device():
    from queasars_b200 import B200SamplerV2

    n, shots = 14, 3000
    cases = [evqe_case(n, 3, 700 + s) for s in range(9)]
    circuits, params = [c for _, _, c in cases], [v for _, v, _ in cases]
    multi = B200SamplerV2(devices="all", seed=11, coalesce=False).sample_indices(circuits, params, shots)
    single = B200SamplerV2(device=0, seed=11, coalesce=False).sample_indices(circuits, params, shots)
    assert multi.shape == (9, shots) and np.array_equal(multi, single)
    want = oq.sample_indices(oq.statevector(cases[4][0], n, cases[4][1]), shots, seed=11)
    assert np.count_nonzero(multi[4] != want) <= 1


def test_threads_sharing_one_multi_device_primitive():
    """The reference's pattern: ``population_size`` threads around one primitive (evqe.py:232-236), here on a device set."""
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator

    n = 13
    op = SparsePauliOp.from_list(random_ising(n, 8))
    ev = B200OperatorCircuitEvaluator(B200EstimatorV2(devices="all", coalesce=True), 0.0, op)
    cases = [evqe_case(n, 2 + s % 3, 900 + s) for s in range(16)]
    sequential = [ev.evaluate_circuits([c], [v])[0] for _, v, c in cases]
    results = [None] * len(cases)

    def work(i):
        for _ in range(6):
            results[i] = ev.evaluate_circuits([cases[i][2]], [cases[i][1]])[0]

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    [t.start() for t in threads]
    [t.join() for t in threads]
    np.testing.assert_allclose(results, sequential, rtol=0, atol=1e-12)


def test_circuit_and_operator_edited_in_place_are_recompiled():
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator
    from queasars_b200.circuit import QuantumCircuit

    n = 6
    terms = random_ising(n, 2)
    op = SparsePauliOp.from_list(terms)
    ev = B200OperatorCircuitEvaluator(B200EstimatorV2(device=0), 0.0, op)
    circ = QuantumCircuit(n)
    instr = []
    for q in range(n):
        circ.ry(0.3 + 0.2 * q, q)
        instr.append(("ry", (q,), (0.3 + 0.2 * q,)))
    table = oq.diagonal_table(n, oq.diag_terms_from_labels(terms))
    first = ev.evaluate_circuits([circ], [[]])[0]
    assert rel_err(first, float(np.dot(np.abs(oq.statevector(instr, n)) ** 2, table))) < 1e-10
    circ.cx(0, 3), circ.h(2)  # the caller appends gates to the SAME circuit object
    instr += [("cx", (0, 3), ()), ("h", (2,), ())]
    second = ev.evaluate_circuits([circ], [[]])[0]
    assert rel_err(second, float(np.dot(np.abs(oq.statevector(instr, n)) ** 2, table))) < 1e-10
    assert abs(second - first) > 1e-6
    op._c[0] = op._c[0] + 2.0  # and edits a coefficient of the SAME operator object
    terms2 = [(terms[0][0], terms[0][1] + 2.0)] + terms[1:]
    table2 = oq.diagonal_table(n, oq.diag_terms_from_labels(terms2))
    third = ev.evaluate_circuits([circ], [[]])[0]
    assert rel_err(third, float(np.dot(np.abs(oq.statevector(instr, n)) ** 2, table2))) < 1e-10


def test_individuals_in_place_of_circuits_on_the_gpu():
    """Direct genome -> gate-list front end through the evaluator: same values as the circuit route, and
    as the oracle on the circuit the reference builds from the genome (individual.py:288-322)."""
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator
    from queasars_b200 import genome as gn

    n = 12
    terms = random_ising(n, 6)
    op = SparsePauliOp.from_list(terms)
    table = oq.diagonal_table(n, oq.diag_terms_from_labels(terms))
    pop = gn.random_population(n, 4, 6, True, 3)
    values = [list(i.parameter_values) for i in pop]
    ev = B200OperatorCircuitEvaluator(B200EstimatorV2(device=0, coalesce=False), 0.0, op)
    direct = ev.evaluate_circuits(pop, values)
    circuits = [i.to_circuit() for i in pop]
    np.testing.assert_allclose(direct, ev.evaluate_circuits(circuits, values), rtol=0, atol=1e-12)
    for g, circ, vals in zip(direct, circuits, values):
        instr = []
        for inst in circ.data:
            ps = tuple(p.name if hasattr(p, "name") else float(p) for p in inst.operation.params)
            instr.append((inst.operation.name, tuple(q._index for q in inst.qubits), ps))
        assert rel_err(g, float(np.dot(np.abs(oq.statevector(instr, n, vals)) ** 2, table))) < 1e-10


def test_engine_calls_leave_the_callers_current_device_alone():
    """A host thread that drives engines on several GPUs (or shares the thread with torch) must find its current CUDA device
    unchanged after every native call: torch would otherwise allocate -- and run NCCL collectives -- on the wrong GPU."""
    import torch

    from queasars_b200 import _native
    from queasars_b200.primitives import get_engine

    n_dev = _native.device_count()
    torch.cuda.set_device(0)
    probe = torch.zeros(1, device="cuda")
    last = get_engine(n_dev - 1)
    instr, values, circ = evqe_case(12, 2, 5)
    from queasars_b200 import gate_list as gl

    plan = last.compile(gl.from_circuit(circ))
    ham = last.hamiltonian(SparsePauliOp.from_list(random_ising(12, 2)))
    last.expectation([plan], [values], ham)
    last.expectation([plan] * 5, [values] * 5, ham)
    last.statevector(plan, values)
    assert torch.cuda.current_device() == 0
    assert torch.zeros(1, device="cuda").device == probe.device
