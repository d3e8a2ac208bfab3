"""Where the time of BASELINE config C3 goes (24-qubit TFIM through the estimator evaluator, batch of 4): end-to-end call,
device-resident run, per-sweep CUDA-event times, cProfile of the host side."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator
from queasars_b200 import genome as gn
pop = gn.random_population(24, 6, 4, True, 0)
circuits, params = [i.to_circuit() for i in pop], [list(i.parameter_values) for i in pop]
est = B200EstimatorV2(coalesce=False)
ev = B200OperatorCircuitEvaluator(est, 0.0, gn.tfim_operator(24))
for _ in range(5): ev.evaluate_circuits(circuits, params)
t0=time.perf_counter()
for _ in range(10): ev.evaluate_circuits(circuits, params)
print("ms per call", (time.perf_counter()-t0)/10*1e3)
plans = [est._cache.plan_for(c) for c in circuits]
ham = est.hamiltonian_for(gn.tfim_operator(24))
rb = est.engine.resident_batch(plans, ham); rb.set_params(params)
for _ in range(3): rb.run()
est.engine.synchronize()
t0=time.perf_counter()
for _ in range(10): rb.run()
est.engine.synchronize()
print("resident ms per run", (time.perf_counter()-t0)/10*1e3)
ms, states = rb.run_timed(); print("sweep ms", [round(float(m),3) for m in ms], "sum", float(ms.sum()))
pr = cProfile.Profile(); pr.enable()
for _ in range(10): ev.evaluate_circuits(circuits, params)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(12)
