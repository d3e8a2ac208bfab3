"""The UNMODIFIED reference package (imported from /root/reference, build container only) driven end to end through
the stand-in Qiskit types of ``queasars_b200.qiskit_compat`` -- the same containers / pub contract / optimizers the
B200 primitives use -- with oracle-backed CPU primitives in place of the GPU engine.  Mirrors the reference's own
end-to-end test (test/minimum_eigensolvers/evqe/test_evqe_algorithm.py:23-38, solver.py:17-53): EVQE on
min x^2 - y^2, x, y in [0, 3] must find [0, 3] (ground state '1100' of the Ising form)."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from tests import reference_loop

pytestmark = [pytest.mark.reference, pytest.mark.skipif(reference_loop.locate_reference() is None, reason="reference package not present (/root/reference or baseline/_ref)")]


def _instructions(circuit):
    out = []
    for inst in circuit.data:
        params = []
        for p in inst.operation.params:
            if hasattr(p, "parameters") and p.parameters:
                (prm,) = p.parameters
                params.append((p._terms[prm], prm.name, p._const))
            else:
                params.append(float(p))
        out.append((inst.operation.name, tuple(circuit.find_bit(q).index for q in inst.qubits), tuple(params)))
    return out


class OracleEstimatorV2:
    """EstimatorV2 contract on top of the NumPy oracle (checker, CPU)."""

    def __init__(self, seed=None):
        self.seed = seed
        self.calls = 0

    def run(self, pubs, *, precision=None):
        from oracle import qiskit_semantics as oq
        from queasars_b200 import containers as ct

        results = []
        for pub in pubs:
            pub = ct.EstimatorPub.coerce(pub, precision)
            circuit = pub.circuit
            state = oq.statevector(_instructions(circuit), circuit.num_qubits, list(np.asarray(pub.parameter_values).reshape(-1)))
            ev = oq.estimator_expectation(state, pub.observables.to_list())
            ev = oq.estimator_value(ev, pub.precision or 0.0, self.seed)
            results.append(ct.PubResult(ct.DataBin(evs=np.asarray(ev), stds=np.asarray(pub.precision or 0.0))))
            self.calls += 1
        return ct.FinishedJob(ct.PrimitiveResult(results, metadata={"version": 2}))


class OracleSamplerV2:
    def __init__(self, seed=None):
        self.seed = seed

    def run(self, pubs, *, shots=None):
        from oracle import qiskit_semantics as oq
        from queasars_b200 import containers as ct

        results = []
        for pub in pubs:
            pub = ct.SamplerPub.coerce(pub, shots)
            circuit = pub.circuit
            state = oq.statevector(_instructions(circuit), circuit.num_qubits, list(np.asarray(pub.parameter_values).reshape(-1)))
            idx = oq.sample_indices(state, pub.shots, seed=self.seed)
            results.append(ct.SamplerPubResult(ct.DataBin(meas=ct.ShotRegister(idx, circuit.num_qubits))))
        return ct.FinishedJob(ct.PrimitiveResult(results, metadata={"version": 2}))


@pytest.fixture(scope="module")
def reference_modules():
    return reference_loop.import_reference()


def test_reference_evqe_finds_ground_state(reference_modules):
    with ThreadPoolExecutor(max_workers=4) as pool:
        solver = reference_loop.sample_solver(reference_modules, OracleEstimatorV2(seed=2), OracleSamplerV2(seed=1), pool, mutex=False)
        result = solver.compute_minimum_eigenvalue(operator=reference_loop.test_model_hamiltonian())
    assert reference_loop.likeliest_bitstring(result) == "1100"  # x = 0, y = 3
    assert result.eigenvalue == pytest.approx(-9.0, abs=0.5)
    assert result.circuit_evaluations and sum(result.circuit_evaluations) > 100


def test_reference_batching_mutex_wrappers_accept_the_contract(reference_modules):
    """One generation through BatchingMutex* + Transpiling* (0.1 s batching sleep per call: keep it short)."""
    with ThreadPoolExecutor(max_workers=10) as pool:
        solver = reference_loop.sample_solver(reference_modules, OracleEstimatorV2(seed=2), OracleSamplerV2(seed=1), pool, mutex=True, max_generations=1)
        result = solver.compute_minimum_eigenvalue(operator=reference_loop.test_model_hamiltonian())
    assert result.eigenvalue < 0
