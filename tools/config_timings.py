"""Wall-clock timings of BASELINE configs C3 (24 q TFIM Pauli sum, estimator route) and C4 (26 q JSSP-shaped diagonal
operator, sampler route with 10k shots) through the public evaluators."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator, B200OperatorSamplerCircuitEvaluator, B200SamplerV2  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402
from queasars_b200.operators import SparsePauliOp  # noqa: E402


def timed(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    return (time.perf_counter() - t0) / reps, out


out = {}
pop = gn.random_population(24, 6, 4, True, 0)
circuits, params = [i.to_circuit() for i in pop], [list(i.parameter_values) for i in pop]
ev = B200OperatorCircuitEvaluator(B200EstimatorV2(coalesce=False), 0.0, gn.tfim_operator(24))
dt, vals = timed(lambda: ev.evaluate_circuits(circuits, params), 5)
out["C3_24q_tfim"] = {"batch": 4, "s_per_call": dt, "evals_per_s": 4 / dt, "values": vals[:2]}

golden = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "jssp_hamiltonians.json")))["jssp_26q"]
op = SparsePauliOp._raw(26, [0] * golden["n_raw_terms"], golden["z_masks"], golden["coeffs"])
pop = gn.random_population(26, 4, 4, True, 1)
circuits, params = [i.to_circuit() for i in pop], [list(i.parameter_values) for i in pop]
for alpha in (1.0, 0.5):
    evs = B200OperatorSamplerCircuitEvaluator(B200SamplerV2(seed=3, coalesce=False), 10000, op, alpha=alpha)
    dt, vals = timed(lambda: evs.evaluate_circuits(circuits, params), 3)
    out[f"C4_26q_jssp_10k_shots_alpha{alpha}"] = {"batch": 4, "s_per_call": dt, "evals_per_s": 4 / dt, "values": vals[:2]}
print(json.dumps(out))
