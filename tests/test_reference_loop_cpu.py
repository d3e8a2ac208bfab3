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

REFERENCE = "/root/reference"
pytestmark = [pytest.mark.reference, pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not present")]


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
    from queasars_b200 import qiskit_compat

    qiskit_compat.install()
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    from queasars.circuit_evaluation.configured_primitives import ConfiguredEstimatorV2, ConfiguredSamplerV2
    from queasars.minimum_eigensolvers.base.termination_criteria import BestIndividualRelativeChangeTolerance
    from queasars.minimum_eigensolvers.evqe.evqe import EVQEMinimumEigensolver, EVQEMinimumEigensolverConfiguration

    return ConfiguredEstimatorV2, ConfiguredSamplerV2, BestIndividualRelativeChangeTolerance, EVQEMinimumEigensolver, EVQEMinimumEigensolverConfiguration


def _hamiltonian():
    from queasars_b200.operators import SparsePauliOp

    return SparsePauliOp.from_list([("IIIZ", -1.5), ("IIZI", -3.0), ("IIZZ", 1.0), ("IZII", 1.5), ("ZIII", 3.0), ("ZZII", -1.0)])


def _solver(mods, executor, mutex, max_generations=None):
    ConfiguredEstimatorV2, ConfiguredSamplerV2, Criterion, Solver, Configuration = mods
    from qiskit_algorithms.optimizers import NFT

    configuration = Configuration(
        configured_sampler=ConfiguredSamplerV2(sampler=OracleSamplerV2(seed=1), shots=1000),
        configured_estimator=ConfiguredEstimatorV2(estimator=OracleEstimatorV2(seed=2), precision=0.05),
        pass_manager=None,
        optimizer=NFT(maxiter=40),
        optimizer_n_circuit_evaluations=40,
        max_generations=max_generations,
        max_circuit_evaluations=None,
        termination_criterion=None if max_generations else Criterion(minimum_relative_change=0.005),
        random_seed=0,
        population_size=10,
        randomize_initial_population_parameters=False,
        speciation_genetic_distance_threshold=3,
        selection_alpha_penalty=0.1,
        selection_beta_penalty=0.1,
        parameter_search_probability=0.24,
        topological_search_probability=0.2,
        layer_removal_probability=0.05,
        parallel_executor=executor,
        mutually_exclusive_primitives=mutex,
    )
    return Solver(configuration=configuration)


def test_reference_evqe_finds_ground_state(reference_modules):
    with ThreadPoolExecutor(max_workers=4) as pool:
        solver = _solver(reference_modules, pool, mutex=False)
        result = solver.compute_minimum_eigenvalue(operator=_hamiltonian())
    probs = result.eigenstate.binary_probabilities()
    best = max(probs, key=probs.get)
    assert best == "1100"  # x = 0, y = 3
    assert result.eigenvalue == pytest.approx(-9.0, abs=0.5)
    assert result.circuit_evaluations and sum(result.circuit_evaluations) > 100


def test_reference_batching_mutex_wrappers_accept_the_contract(reference_modules):
    """One generation through BatchingMutex* + Transpiling* (0.1 s batching sleep per call: keep it short)."""
    with ThreadPoolExecutor(max_workers=10) as pool:
        solver = _solver(reference_modules, pool, mutex=True, max_generations=1)
        result = solver.compute_minimum_eigenvalue(operator=_hamiltonian())
    assert result.eigenvalue < 0
