"""Where a batch-1 evaluation of the optimizer loop spends its time (20 qubits, 6 layers, last layer parameterised, prefix reuse):
whole Python call, native call alone (pre-packed arrays), packing alone, and the device time of the suffix sweeps."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import _native  # noqa: E402
from queasars_b200 import gate_list as gl  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402
from queasars_b200.engine import Engine  # noqa: E402

n, layers, reps = 20, 6, 2000
eng = Engine(0)
ham = eng.hamiltonian(gn.ising_operator(n))
ind = gn.Individual.random(n, layers, True, 3)
gates = gl.from_evqe_individual(ind, {-1})
rng = np.random.default_rng(0)
out = {}
for name, plan in (("full_circuit", eng.compile(gates, drop_final_phases=True)), ("prefix_reuse", eng.compile_with_prefix_reuse(gates, drop_final_phases=True))):
    params = [list(rng.uniform(0, 6.28, gates.n_params))]
    plans = [plan]
    for _ in range(50):
        eng.expectation(plans, params, ham)
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.expectation(plans, params, ham)
    whole = (time.perf_counter() - t0) / reps
    ids, flat, offsets = eng._pack(plans, params)
    res = np.empty(1)
    args = (eng._ctx, 1, _native.ptr(ids), _native.ptr(flat), _native.ptr(offsets), ham.ham_id, _native.ptr(res))
    fn = eng._lib.qb_evaluate_expectation
    for _ in range(50):
        fn(*args)
    t0 = time.perf_counter()
    for _ in range(reps):
        fn(*args)
    native = (time.perf_counter() - t0) / reps
    t0 = time.perf_counter()
    for _ in range(reps):
        eng._pack(plans, params)
    pack = (time.perf_counter() - t0) / reps
    rb = eng.resident_batch(plans, ham)
    rb.set_params(params)
    for _ in range(5):
        rb.run()
    ms = np.zeros(plan.n_sweeps)
    for _ in range(20):
        m, _s = rb.run_timed()
        ms += m
    rb.close()
    out[name] = {"whole_call_us": 1e6 * whole, "native_call_us": 1e6 * native, "pack_us": 1e6 * pack,
                 "python_other_us": 1e6 * (whole - native - pack), "sweeps": int(plan.n_sweeps),
                 "sweep_device_us": [round(1e3 * float(v) / 20, 2) for v in ms]}
print(json.dumps(out, indent=1))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
