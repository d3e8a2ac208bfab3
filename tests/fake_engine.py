"""A stand-in for ``queasars_b200.engine.Engine`` that evaluates with the NumPy oracle -- TEST INFRASTRUCTURE: lets the CPU
suite drive the primitives' and evaluators' host logic (plan / Hamiltonian caches, device split, coalescing queue, pub
containers, locks) end to end without a GPU.  Never used by the product."""
import threading

import numpy as np

from oracle import qiskit_semantics as oq
from queasars_b200 import gate_list as gl
from queasars_b200.engine import HamiltonianHandle, PlanHandle, operator_terms, rewritten


def _state(gates, params):
    n = gates.n_qubits
    st = np.zeros(1 << n, dtype=complex)
    st[0] = 1.0
    for op in gates.ops:
        m = np.array(op.matrix(list(params)))
        if op.control < 0:
            st = oq.apply_matrix(st, n, m, [op.target])
        else:
            cm = np.kron(np.eye(2), np.diag([1, 0])) + np.kron(m, np.diag([0, 1]))
            st = oq.apply_matrix(st, n, cm, [op.control, op.target])
    return st


class FakeEngine:
    workspace_bytes = 140 << 30

    def __init__(self, device=0, dtype="complex128"):
        self.device, self.dtype = device, dtype
        self.launch_count = 0
        self._lock = threading.Lock()
        self._plans, self._hams, self._next = {}, {}, 1

    def _new_id(self):
        with self._lock:
            self._next += 1
            return self._next

    def compile(self, gates, dtype=None, from_zero_state=True, cache=True, defer=True, drop_final_phases=False):
        if defer:
            gates = rewritten(gates, drop_final_phases)
        pid = self._new_id()
        self._plans[pid] = gates
        return PlanHandle(pid, gates.n_qubits, gates.n_params, len(gates.ops), 1, 1, 0)

    def compile_with_prefix_reuse(self, gates, dtype=None, min_prefix_ops=4, drop_final_phases=False):
        return self.compile(gates, drop_final_phases=drop_final_phases)

    def hamiltonian(self, operator, build_table=None):
        n, x, z, c = operator_terms(operator)
        hid = self._new_id()
        self._hams[hid] = (n, [int(v) for v in x], [int(v) for v in z], [complex(v) for v in c])
        diag = not any(int(v) for v in x)
        return HamiltonianHandle(hid, n, diag, np.asarray(z), np.asarray(c).real, len(c))

    def expectation(self, plans, params, ham):
        n, xs, zs, cs = self._hams[ham.ham_id]
        out = []
        for plan, vals in zip(plans, params):
            vals = np.asarray(vals, dtype=np.float64).reshape(-1)
            if vals.size != plan.n_params:
                raise ValueError(f"circuit has {plan.n_params} parameters but {vals.size} values were given")
            st = _state(self._plans[plan.plan_id], vals)
            idx = np.arange(st.size, dtype=np.uint64)
            total = 0.0
            for x, z, c in zip(xs, zs, cs):
                sign = 1.0 - 2.0 * oq._parity(idx & np.uint64(z))
                ny = bin(x & z).count("1")
                total += (c * (1j**ny) * np.sum(np.conj(st[(idx ^ np.uint64(x)).astype(np.int64)]) * sign * st)).real
            out.append(total)
            self.launch_count += 1
        return np.asarray(out)

    def expectation_submit(self, plans, params, ham):
        self._queued = self.expectation(plans, params, ham)
        return len(plans)

    def expectation_collect(self, count):
        out, self._queued = self._queued, None
        assert len(out) == count
        return out

    def sample(self, plans, params, shots, uniforms):
        out = np.empty((len(plans), shots), dtype=np.int64)
        for i, (plan, vals) in enumerate(zip(plans, params)):
            st = _state(self._plans[plan.plan_id], np.asarray(vals, dtype=np.float64).reshape(-1))
            out[i] = oq.sample_indices(st, shots, uniforms=np.asarray(uniforms)[i])
            self.launch_count += 1
        return out

    def diag_energies(self, ham, states):
        n, xs, zs, cs = self._hams[ham.ham_id]
        return np.asarray([oq.diagonal_energy(int(s), [(z, c.real) for z, c in zip(zs, cs)]) for s in np.asarray(states).reshape(-1)])
