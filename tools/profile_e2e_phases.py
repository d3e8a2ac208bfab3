"""Host-side phases of one pipelined evaluate_circuits call on the bench workload (20 q x 32): where the time between the
call and the first kernel goes.  Times are medians over 200 calls, in microseconds."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import B200EstimatorV2, _native  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402

inds = gn.random_population(20, 6, 32, True, 0)
circuits = [i.to_circuit() for i in inds]
params = [list(i.parameter_values) for i in inds]
est = B200EstimatorV2(coalesce=False)
op = gn.ising_operator(20)
for _ in range(20):
    est.expectation_values(circuits, params, op)
eng, lib = est.engine, est.engine._lib
rows = []
for _ in range(200):
    t = [time.perf_counter()]
    ham = est.hamiltonian_for(op)
    resolved = [est._resolve(c, v) for c, v in zip(circuits, params)]
    plans, vals = [r[0] for r in resolved], [r[1] for r in resolved]
    t.append(time.perf_counter())
    split = eng._pipeline_split(plans, vals)
    t.append(time.perf_counter())
    ids, flat, offsets = eng._pack(plans[:split], vals[:split])
    t.append(time.perf_counter())
    _native.check(lib.qb_evaluate_expectation_submit(eng._ctx, split, _native.ptr(ids), _native.ptr(flat), _native.ptr(offsets), ham.ham_id))
    t.append(time.perf_counter())
    ids, flat, offsets = eng._pack(plans[split:], vals[split:])
    t.append(time.perf_counter())
    _native.check(lib.qb_evaluate_expectation_submit(eng._ctx, len(plans) - split, _native.ptr(ids), _native.ptr(flat), _native.ptr(offsets), ham.ham_id))
    t.append(time.perf_counter())
    out = np.empty(len(plans))
    _native.check(lib.qb_evaluate_expectation_collect(eng._ctx, len(plans), _native.ptr(out)))
    t.append(time.perf_counter())
    rows.append(np.diff(t))
med = np.median(np.asarray(rows), axis=0) * 1e6
names = ["resolve", "split", "pack chunk 1", "submit chunk 1", "pack chunk 2", "submit chunk 2", "collect (wait)"]
print("split at", split, {n: round(float(v), 1) for n, v in zip(names, med)}, "total", round(float(med.sum()), 1))
