"""BASELINE config C1: the EVQE last-layer search on the smallest JSSP instances (4 / 5 / 8 qubits, population 10, NFT with
maxfev = 40 per individual, grouped evaluations of 2, ten optimizer chains in a thread pool exactly like
evqe/evolutionary_algorithm/mutation.py:194-235) through the B200 evaluators.  Reports objective evaluations per second and
the time per optimizer pass; at these sizes every call is launch-latency bound (a 4-qubit state is 256 B)."""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator, B200OperatorSamplerCircuitEvaluator, B200SamplerV2  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402
from queasars_b200.operators import SparsePauliOp  # noqa: E402
from queasars_b200.optimizers import NFT  # noqa: E402

golden = json.load(open(os.path.join(ROOT, "tests", "golden", "jssp_hamiltonians.json")))


def optimize_last_layer(individual, evaluator, optimizer, counter):
    circuit = individual.to_circuit({-1})
    x0 = np.asarray(individual.layer_values(-1))
    n_params = len(x0)

    def objective(x):
        rows = np.reshape(x, (-1, n_params)).tolist()
        counter[0] += len(rows)
        vals = evaluator.evaluate_circuits([circuit] * len(rows), rows)
        return vals[0] if len(vals) == 1 else np.asarray(vals)

    return float(optimizer.minimize(fun=objective, x0=x0, bounds=[(None, None)] * n_params).fun)


out = {}
for key in ("jssp_4q", "jssp_5q", "jssp_8q"):
    entry = golden[key]
    n = entry["n_qubits"]
    op = SparsePauliOp._raw(n, [0] * entry["n_raw_terms"], entry["z_masks"], entry["coeffs"])
    for route in ("estimator", "sampler_cvar"):
        if route == "estimator":
            evaluator = B200OperatorCircuitEvaluator(B200EstimatorV2(seed=0), 0.0, op)
        else:
            evaluator = B200OperatorSamplerCircuitEvaluator(B200SamplerV2(seed=0), 512, op, alpha=0.5)
        population = gn.random_population(n, 2, 10, True, 0)
        best = None
        for rep in range(3):  # first repetition compiles plans / caches prefix states
            counters = [[0] for _ in population]
            optimizers = []
            for _ in population:
                opt = NFT(maxfev=40)
                opt.set_max_evals_grouped(2)
                optimizers.append(opt)
            t0 = time.perf_counter()
            with ThreadPoolExecutor(max_workers=len(population)) as pool:
                after = list(pool.map(lambda a: optimize_last_layer(a[0], evaluator, a[1], a[2]), zip(population, optimizers, counters)))
            dt = time.perf_counter() - t0
            evals = sum(c[0] for c in counters)
            if best is None or dt < best[0]:
                best = (dt, evals, min(after))
        out[f"{key}_{route}"] = {"pass_s": best[0], "objective_evals": best[1], "evals_per_s": best[1] / best[0], "best_value": best[2], "ground_state": entry["lowest"][0][1]}
print(json.dumps(out))
