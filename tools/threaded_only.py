"""bench.py's e2e_threaded probe alone (32 threads x single-circuit evaluate_circuits calls through the coalescing queue)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402
from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator  # noqa: E402

args = argparse.Namespace(steps=100, layers=6)
individuals, circuits, params = bench.build_workload(20, 6, 32, 0)
op = gn.ising_operator(20)
values = B200OperatorCircuitEvaluator(B200EstimatorV2(device=0, coalesce=False), 0.0, op).evaluate_circuits(circuits, params)
for _ in range(3):
    print(bench.threaded_probe(0, op, circuits, params, values, args))
