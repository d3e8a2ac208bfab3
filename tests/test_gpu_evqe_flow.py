"""The callers of the hot path on the GPU: per-individual optimizer chains over one
layer's parameters, run concurrently from a thread pool like ``BaseEVQEMutationOperator.apply_operator``
(/root/reference/queasars/minimum_eigensolvers/evqe/evolutionary_algorithm/mutation.py:28-89, 194-235), with the same
monotone assertions the reference's operator tests make (test/minimum_eigensolvers/evqe/test_evqe_operators.py:64-93)."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from queasars_b200 import genome as gn
from queasars_b200.operators import SparsePauliOp
from queasars_b200.optimizers import NFT, SPSA

pytestmark = pytest.mark.gpu


def optimize_last_layer(individual, evaluator, optimizer):
    """Same shape as mutation.py:28-89: a partially parameterised circuit, a batched objective, optimizer.minimize."""
    circuit = individual.to_circuit({-1})
    x0 = np.asarray(individual.layer_values(-1))
    n_params = len(x0)

    def objective(x):
        rows = np.reshape(x, (-1, n_params)).tolist()
        vals = evaluator.evaluate_circuits([circuit] * len(rows), rows)
        return vals[0] if len(vals) == 1 else np.asarray(vals)

    result = optimizer.minimize(fun=objective, x0=x0, bounds=[(None, None)] * n_params)
    return float(result.fun), int(result.nfev)


@pytest.mark.parametrize("route", ["estimator", "sampler_cvar"])
def test_last_layer_search_lowers_population_energy(jssp_golden, route):
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator, B200OperatorSamplerCircuitEvaluator, B200SamplerV2, qiskit_compat

    qiskit_compat.install()
    entry = jssp_golden["jssp_4q"]
    n = entry["n_qubits"]
    op = SparsePauliOp._raw(n, [0] * entry["n_raw_terms"], entry["z_masks"], entry["coeffs"])
    minimum = entry["lowest"][0][1]  # 63.5: the notebook's converged objective
    if route == "estimator":
        evaluator = B200OperatorCircuitEvaluator(B200EstimatorV2(seed=0), 0.0, op)
        make_optimizer = lambda: NFT(maxfev=40)  # noqa: E731  (the reference's operator tests use NFT(maxfev=40))
    else:
        evaluator = B200OperatorSamplerCircuitEvaluator(B200SamplerV2(seed=0), 512, op, alpha=0.5)
        make_optimizer = lambda: SPSA(maxiter=33, perturbation=0.35, learning_rate=0.43, trust_region=True)  # noqa: E731
    population = gn.random_population(n, 2, 10, True, 0)
    before = evaluator.evaluate_circuits([i.to_circuit() for i in population], [list(i.parameter_values) for i in population])
    optimizers = []
    for _ in population:
        opt = make_optimizer()
        opt.set_max_evals_grouped(2)
        optimizers.append(opt)
    with ThreadPoolExecutor(max_workers=len(population)) as pool:
        results = list(pool.map(lambda args: optimize_last_layer(args[0], evaluator, args[1]), zip(population, optimizers)))
    after = [r[0] for r in results]
    assert sum(after) < sum(before)
    assert all(v >= minimum - 1e-9 for v in after)  # nothing can undercut the ground-state energy
    assert sum(r[1] for r in results) >= 10 * 30
    if route == "estimator":
        assert min(after) < min(before) + 1e-9
