"""bench.py's optimizer_calls probe alone."""
import json
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import bench  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402

individuals = gn.random_population(20, 6, 32, True, 1000)
out = bench.optimizer_calls_probe(0, gn.ising_operator(20), individuals)
print(json.dumps(out, indent=1))
