"""Small driver for ncu captures: one resident batch of `--batch` random EVQE individuals on `--n` qubits,
`--runs` evaluations.  Usage (on the GPU box):  python tools/profile_case.py --n 26 --layers 6 --runs 3"""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))

from queasars_b200 import gate_list as gl  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402
from queasars_b200.engine import Engine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=26)
ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--runs", type=int, default=3)
ap.add_argument("--dtype", default="complex128")
args = ap.parse_args()

engine = Engine(0, args.dtype)
inds = gn.random_population(args.n, args.layers, args.batch, True, 7)
plans = [engine.compile(gl.from_evqe_individual(i)) for i in inds]
ham = engine.hamiltonian(gn.ising_operator(args.n)) if args.n <= 26 else None
rb = engine.resident_batch(plans, ham)
rb.set_params([list(i.parameter_values) for i in inds])
for _ in range(args.runs):
    rb.run()
engine.synchronize()
ms, states = rb.run_timed()
print("sweeps", len(ms), "ms", [round(float(m), 4) for m in ms], "states", list(states), "ops", [p.n_ops for p in plans][:4])
if ham is not None:
    print("values", rb.read()[:4])
