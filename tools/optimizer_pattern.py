"""Evaluations/s in the optimizer's calling pattern (one circuit object, one layer parameterised, sequential calls of
batch 1 or 2) with and without prefix-state reuse.  20 qubits, L layers, random diagonal Ising Hamiltonian."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import gate_list as gl  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402
from queasars_b200.engine import Engine  # noqa: E402

n, layers = 20, 6
eng = Engine(0)
ham = eng.hamiltonian(gn.ising_operator(n))
ind = gn.Individual.random(n, layers, True, 3)
gates = gl.from_evqe_individual(ind, {-1})
rng = np.random.default_rng(0)
out = {"n_qubits": n, "layers": layers, "ops_total": len(gates.ops)}
# plans as the evaluators compile them for a diagonal Hamiltonian (trailing phases deferred and dropped)
for name, plan in (("full_circuit", eng.compile(gates, drop_final_phases=True)), ("prefix_reuse", eng.compile_with_prefix_reuse(gates, drop_final_phases=True))):
    for batch in (1, 2, 8):
        params = [list(rng.uniform(0, 6.28, gates.n_params)) for _ in range(batch)]
        for _ in range(20):
            eng.expectation([plan] * batch, params, ham)
        t0 = time.perf_counter()
        reps = 300
        for _ in range(reps):
            eng.expectation([plan] * batch, params, ham)
        dt = time.perf_counter() - t0
        out[f"{name}_batch{batch}"] = {"evals_per_s": batch * reps / dt, "us_per_call": 1e6 * dt / reps, "ops_per_eval": plan.n_ops, "sweeps": plan.n_sweeps}
print(json.dumps(out, indent=1))
