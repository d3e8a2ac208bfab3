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
ap.add_argument("--oploop", type=int, default=0, help="op-loop probe: u on every qubit (absorbed), then this many layers of u on the 4 highest qubits (one pass)")
ap.add_argument("--hbm-regime", type=int, default=0, help="the bench's hbm_regime circuit: u on every qubit (absorbed into the product start), then 3 layers of 7 u gates on non-low qubits")
ap.add_argument("--keep-phases", type=int, default=0, help="1: plans keep the final phases (statevector semantics); default: probabilities only, like the evaluators with a diagonal Hamiltonian")
ap.add_argument("--simple", type=int, default=0, help="instead of an EVQE genome: this many u gates on distinct high qubits, repeated --layers times")
args = ap.parse_args()

engine = Engine(0, args.dtype)
inds = gn.random_population(args.n, args.layers, args.batch, True, 7)
if args.oploop:
    from queasars_b200.circuit import QuantumCircuit

    circ = QuantumCircuit(args.n)
    for q in range(args.n):
        circ.u(0.1 + 0.01 * q, 0.2, 0.3, q)
    for layer in range(args.oploop):
        for g in range(4):
            circ.u(0.3 + g, 0.2 * layer, 0.1, args.n - 1 - g)
    plans = [engine.compile(gl.from_circuit(circ))] * args.batch
    params = [[] for _ in range(args.batch)]
elif args.hbm_regime:
    from queasars_b200.circuit import QuantumCircuit

    n = args.n
    circ = QuantumCircuit(n)
    for q in range(n):
        circ.u(0.1 + 0.01 * q, 0.2, 0.3, q)
    for layer in range(3):
        for g in range(7):
            circ.u(0.3 + g, 0.2 * layer, 0.1, 4 + ((g * 3 + 7 * layer) % (n - 4)))
    plans = [engine.compile(gl.from_circuit(circ), drop_final_phases=not args.keep_phases)] * args.batch
    params = [[] for _ in range(args.batch)]
elif args.simple:
    from queasars_b200.circuit import QuantumCircuit

    circ = QuantumCircuit(args.n)
    for layer in range(args.layers):
        for g in range(args.simple):
            circ.u(0.3 + g, 0.2 * layer, 0.1, args.n - 1 - ((g + 8 * layer) % (args.n - 4)))
    plans = [engine.compile(gl.from_circuit(circ))] * args.batch
    params = [[] for _ in range(args.batch)]
else:
    plans = [engine.compile(gl.from_evqe_individual(i), drop_final_phases=not args.keep_phases) for i in inds]
    params = [list(i.parameter_values) for i in inds]
ham = engine.hamiltonian(gn.ising_operator(args.n)) if args.n <= 26 else None
rb = engine.resident_batch(plans, ham)
rb.set_params(params)
for _ in range(args.runs):
    rb.run()
engine.synchronize()
ms, states = rb.run_timed()
gbs = [round(rb.stats()["sweep_bytes"] * int(st) / (float(m) * 1e-3) / 1e9) for m, st in zip(ms, states)]
print("GB/s per sweep launch", gbs, "passes", [p.n_passes for p in plans][:2])
print("sweeps", len(ms), "ms", [round(float(m), 4) for m in ms], "states", list(states), "ops", [p.n_ops for p in plans][:4])
if ham is not None:
    print("values", rb.read()[:4])
