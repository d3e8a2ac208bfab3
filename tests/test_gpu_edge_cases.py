"""Edge cases of the drop-in boundary on the GPU: empty and ragged inputs, None entries, one-qubit registers, one shot, batches
larger than the workspace (memory chunking) and larger than a launch's grid.y (65 535), identical circuits repeated in one call
(the reference passes ``[circuit] * batch_size``: mutation.py:66-69)."""
import numpy as np
import pytest

from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq
from queasars_b200 import gate_list as gl
from queasars_b200.circuit import Parameter, QuantumCircuit
from queasars_b200.operators import SparsePauliOp
from tests.test_frontend_planner import build_circuit
from tests.test_gpu_parity import random_ising, rel_err

pytestmark = pytest.mark.gpu


def evqe_case(n, layers, seed):
    genome, values = og.random_individual(n, layers, True, seed)
    instr = og.individual_circuit(genome, values)
    return instr, list(values), build_circuit(instr, n)


def test_empty_none_and_ragged_inputs():
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator, B200OperatorSamplerCircuitEvaluator, B200SamplerV2

    n = 5
    op = SparsePauliOp.from_list(random_ising(n, 1))
    ev = B200OperatorCircuitEvaluator(B200EstimatorV2(device=0), 0.0, op)
    smp = B200OperatorSamplerCircuitEvaluator(B200SamplerV2(device=0, seed=1), 100, op)
    assert ev.evaluate_circuits([], []) == []
    assert smp.evaluate_circuits([], []) == []
    instr, values, circ = evqe_case(n, 2, 0)
    # None entries are skipped like the reference's tuple(... if circuit is not None and parameter_values is not None)
    got = ev.evaluate_circuits([circ, None, circ], [values, values, values])
    assert len(got) == 2 and got[0] == got[1]
    with pytest.raises(ValueError):
        ev.evaluate_circuits([circ], [values[:-1]])
    with pytest.raises(ValueError):
        ev.evaluate_circuits([circ], [values + [0.1]])
    with pytest.raises(ValueError):
        smp.evaluate_circuits([circ], [values[:-2]])
    # a failing call leaves the engine usable
    table = oq.diagonal_table(n, oq.diag_terms_from_labels(random_ising(n, 1)))
    assert rel_err(ev.evaluate_circuits([circ], [values])[0], float(np.dot(np.abs(oq.statevector(instr, n, values)) ** 2, table))) < 1e-10


def test_one_qubit_register_and_one_shot():
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator, B200SamplerV2
    from queasars_b200.evaluators import measure_quasi_distributions

    theta = Parameter("t")
    circ = QuantumCircuit(1)
    circ.ry(theta, 0)
    op = SparsePauliOp.from_list([("Z", 1.0), ("X", 0.5)])
    ev = B200OperatorCircuitEvaluator(B200EstimatorV2(device=0), 0.0, op)
    for t in (0.0, 0.7, np.pi):
        assert ev.evaluate_circuits([circ], [[t]])[0] == pytest.approx(np.cos(t) + 0.5 * np.sin(t), abs=1e-12)
    dist = measure_quasi_distributions([circ], [[np.pi]], B200SamplerV2(device=0, seed=0), 1)[0]
    assert dict(dist) == {1: 1.0}


def test_batch_larger_than_the_workspace_is_chunked():
    from queasars_b200.engine import Engine

    n = 14
    eng = Engine(device=0, workspace_limit=5 * (16 << n))  # room for 5 states: 23 evaluations run as 5 chunks
    terms = random_ising(n, 3)
    ham = eng.hamiltonian(SparsePauliOp.from_list(terms), build_table=True)
    table = oq.diagonal_table(n, oq.diag_terms_from_labels(terms))
    cases = [evqe_case(n, 2 + s % 2, 300 + s) for s in range(23)]
    plans = [eng.compile(gl.from_circuit(c), drop_final_phases=True) for _, _, c in cases]
    params = [v for _, v, _ in cases]
    got = eng.expectation(plans, params, ham)
    for g, (instr, values, _) in zip(got[::4], cases[::4]):
        assert rel_err(g, float(np.dot(np.abs(oq.statevector(instr, n, values)) ** 2, table))) < 1e-10
    full = Engine(device=0)
    ham2 = full.hamiltonian(SparsePauliOp.from_list(terms), build_table=True)
    plans2 = [full.compile(gl.from_circuit(c), drop_final_phases=True) for _, _, c in cases]
    np.testing.assert_allclose(got, full.expectation(plans2, params, ham2), rtol=0, atol=1e-12)
    idx = eng.sample(plans, params, 50, np.random.default_rng(0).random((23, 50)))
    np.testing.assert_array_equal(idx, full.sample(plans2, params, 50, np.random.default_rng(0).random((23, 50))))
    eng.close(), full.close()


def test_more_evaluations_than_one_launch_can_index():
    """70 000 evaluations of ONE small circuit in one call: beyond the 65 535 entries a launch's grid.y can address."""
    from queasars_b200.engine import Engine

    n, count = 6, 70000
    eng = Engine(device=0)
    theta = [Parameter(f"t{q}") for q in range(n)]
    circ = QuantumCircuit(n)
    for q in range(n):
        circ.ry(theta[q], q)
    circ.cx(0, 1), circ.cx(2, 3)
    plan = eng.compile(gl.from_circuit(circ), drop_final_phases=True)
    ham = eng.hamiltonian(SparsePauliOp.from_list([("I" * (n - 1) + "Z", 1.0), ("I" * (n - 2) + "ZI", 0.5)]))
    rng = np.random.default_rng(5)
    params = rng.uniform(0, np.pi, size=(count, n))
    got = eng.expectation([plan] * count, params, ham)
    want = np.cos(params[:, 0]) + 0.5 * np.cos(params[:, 0]) * np.cos(params[:, 1])  # <Z0> + 0.5 <Z1> after cx(0, 1)
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)
    eng.close()
