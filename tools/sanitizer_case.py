"""A small pass over every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): 12- and 13-qubit EVQE circuits
through batched expectation (diagonal and Pauli-sum Hamiltonians), the one-circuit CUDA-graph path, prefix-state reuse, sampling,
diagonal energies and the statevector read-back.
    compute-sanitizer --tool racecheck python tools/sanitizer_case.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import gate_list as gl  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402
from queasars_b200.engine import Engine  # noqa: E402
from queasars_b200.operators import SparsePauliOp  # noqa: E402

eng = Engine(0)
rng = np.random.default_rng(0)
for n in (12, 13):
    ham = eng.hamiltonian(gn.ising_operator(n))
    labels = [("".join(rng.choice(list("IXYZ"), size=n)), float(rng.normal())) for _ in range(6)]
    ham_x = eng.hamiltonian(SparsePauliOp.from_list(labels))
    inds = [gn.Individual.random(n, 3, True, s) for s in range(3)]
    gates = [gl.from_evqe_individual(i) for i in inds]
    plans = [eng.compile(g, drop_final_phases=True) for g in gates]
    rows = [list(i.parameter_values) for i in inds]
    print(n, "diag batch", eng.expectation(plans, rows, ham))
    full = [eng.compile(g) for g in gates]
    print(n, "pauli batch", eng.expectation(full, rows, ham_x))
    print(n, "graph 1pt", eng.expectation(plans[:1], rows[:1], ham), "2pt", eng.expectation([plans[0]] * 2, [rows[0], rows[0]], ham))
    last = gl.from_evqe_individual(inds[0], {-1})
    pre = eng.compile_with_prefix_reuse(last, drop_final_phases=True)
    vals = list(rng.uniform(0, 6.28, last.n_params))
    print(n, "prefix", eng.expectation([pre], [vals], ham), eng.expectation([pre] * 5, [vals] * 5, ham)[:2])
    idx = eng.sample(plans, rows, 64, rng.random((3, 64)))
    print(n, "sample", idx[:, :4].tolist(), "energies", eng.diag_energies(ham, idx[0, :4].astype(np.uint64)))
    sv = eng.statevector(full[0], rows[0])
    print(n, "norm", float(np.vdot(sv, sv).real))
print("done")
