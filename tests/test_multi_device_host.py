"""Host logic of the multi-GPU primitives (no GPU needed): how one submission is split over a device set, that results come
back in submission order, that rows of one circuit are dealt out when there are fewer circuits than devices, that a circuit
keeps its home device (plan / cached prefix state) while that does not unbalance the split, and what pickling preserves.
Reference behaviour being mapped: ONE primitive shared by ``population_size`` threads or pickled into dask workers
(/root/reference/queasars/minimum_eigensolvers/evqe/evqe.py:38-44, 232-236; selection.py:75-82)."""
import pickle
import threading

import numpy as np
import pytest

from queasars_b200 import primitives as pr


class FakeEngine:
    def __init__(self, device):
        self.device = device
        self.calls = []


def fake_primitive(n_devices):
    prim = pr.B200EstimatorV2(devices=list(range(n_devices)), coalesce=False)
    prim._engines_obj = [FakeEngine(d) for d in range(n_devices)]
    return prim


def entries(n):
    return [{"home": None, "gates": None} for _ in range(n)]


def test_split_balances_and_covers_every_entry():
    prim = fake_primitive(4)
    ens = entries(32)
    costs = [float(60 + (7 * i) % 23) for i in range(32)]
    slots = prim._assign(ens, costs)
    assert len(slots) == 32 and set(slots) == {0, 1, 2, 3}
    load = [sum(c for c, s in zip(costs, slots) if s == d) for d in range(4)]
    assert max(load) - min(load) <= max(costs)  # LPT bound
    # second submission of the same circuits: everyone stays on its home device
    assert prim._assign(ens, costs) == slots


def test_rows_of_one_circuit_are_dealt_out():
    prim = fake_primitive(4)
    one = entries(1)[0]
    slots = prim._assign([one] * 40, [1.0] * 40)
    counts = [slots.count(d) for d in range(4)]
    assert sorted(counts) == [10, 10, 10, 10]


def test_single_circuit_calls_rotate_over_the_devices():
    prim = fake_primitive(4)
    seen = [prim._assign(entries(1), [1.0])[0] for _ in range(8)]
    assert sorted(set(seen)) == [0, 1, 2, 3]


def test_results_come_back_in_submission_order():
    prim = fake_primitive(3)
    resolved = [(i % 3, f"plan{i}", [float(i)]) for i in range(11)]
    threads = set()

    def call(slot, plans, params):
        threads.add(threading.current_thread().name)
        return [(slot, p, v[0]) for p, v in zip(plans, params)]

    out = prim._run_per_device(resolved, call)
    assert out == [(i % 3, f"plan{i}", float(i)) for i in range(11)]
    assert all(name.startswith("qb-device") for name in threads)

    def failing(slot, plans, params):
        if slot == 1:
            raise RuntimeError("device 1 failed")
        return [0.0] * len(plans)

    with pytest.raises(RuntimeError, match="device 1 failed"):
        prim._run_per_device(resolved, failing)


def test_pickle_keeps_the_device_set_and_rotates_in_other_processes(monkeypatch):
    prim = pr.B200SamplerV2(devices=[0, 1, 2, 3], seed=5, default_shots=77)
    clone = pickle.loads(pickle.dumps(prim))
    assert clone.devices == [0, 1, 2, 3] and clone.seed == 5 and clone.default_shots == 77
    assert clone._device_list() == [0, 1, 2, 3]  # same process: same order
    import os

    real = os.getpid()
    monkeypatch.setattr(os, "getpid", lambda: real + 1)  # "another process"
    k = (real + 1) % 4
    assert clone._device_list() == [0, 1, 2, 3][k:] + [0, 1, 2, 3][:k]
    with pytest.raises(ValueError):
        pr.B200EstimatorV2(devices=[0, 0])
    assert pr.B200EstimatorV2(device=3)._device_list() == [3]


def test_fingerprints_detect_in_place_edits():
    from queasars_b200.circuit import QuantumCircuit
    from queasars_b200.operators import SparsePauliOp

    circ = QuantumCircuit(2)
    circ.h(0)
    before = pr._circuit_fingerprint(circ)
    circ.cx(0, 1)
    assert pr._circuit_fingerprint(circ) != before
    op = SparsePauliOp.from_list([("ZI", 1.0), ("IZ", 0.5)])
    fp = pr._operator_fingerprint(op)
    op._c[1] = 0.25 + 0j
    assert pr._operator_fingerprint(op) != fp


def test_duplicate_pauli_terms_are_merged_in_first_appearance_order():
    """engine.merge_duplicate_terms: the reference's JSSP encoder emits every (x, z) string several times over
    (domain_wall_hamiltonian_encoder.py:189-230: 346 raw terms for 84 distinct strings at 26 qubits)."""
    from queasars_b200.engine import merge_duplicate_terms

    x = [0, 0, 5, 0, 5, 0, 5]
    z = [3, 9, 1, 3, 1, 12, 2]
    c = [1.0, 2.0, 3.0 + 1j, 4.0, 5.0, 6.0, 7.0]
    mx, mz, mc = merge_duplicate_terms(x, z, c)
    assert list(mx) == [0, 0, 5, 0, 5] and list(mz) == [3, 9, 1, 12, 2]
    np.testing.assert_allclose(mc, [5.0, 2.0, 8.0 + 1j, 6.0, 7.0])
    ux, uz, uc = merge_duplicate_terms([1, 2], [0, 0], [1.0, 2.0])
    assert list(ux) == [1, 2] and list(uc) == [1.0, 2.0]


def test_golden_jssp_hamiltonian_merges_to_its_distinct_terms(jssp_golden):
    from queasars_b200.engine import merge_duplicate_terms

    entry = jssp_golden["jssp_26q"]
    z = np.asarray(entry["z_masks"], dtype=np.uint64)
    c = np.asarray(entry["coeffs"], dtype=float)
    zz, cc = np.concatenate([z, z[:40]]), np.concatenate([c, c[:40]])
    x, mz, mc = merge_duplicate_terms(np.zeros_like(zz), zz, cc)
    assert len(mz) == len(set(int(v) for v in z))
    probe = np.asarray(entry["probe_states"], dtype=np.uint64)

    def energy(zz, cc, k):
        return sum(float(np.real(cf)) * (1 - 2 * (bin(int(k) & int(zm)).count("1") & 1)) for zm, cf in zip(zz, cc))

    for k, want in zip(probe, entry["probe_energies"]):
        extra = energy(z[:40], c[:40], k)
        assert energy(mz, mc, k) == pytest.approx(want + extra, rel=1e-12)
