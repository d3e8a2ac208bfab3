"""cProfile of the end-to-end evaluator call (host overhead hunting)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402

inds = gn.random_population(20, 6, 32, True, 0)
circuits = [i.to_circuit() for i in inds]
params = [list(i.parameter_values) for i in inds]
est = B200EstimatorV2(coalesce=False)
ev = B200OperatorCircuitEvaluator(est, 0.0, gn.ising_operator(20))
for _ in range(200):
    ev.evaluate_circuits(circuits, params)
t0 = time.perf_counter()
for _ in range(50):
    ev.evaluate_circuits(circuits, params)
print("ms per call", (time.perf_counter() - t0) / 50 * 1e3)
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    ev.evaluate_circuits(circuits, params)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
