"""Host logic: circuit container + gate-list front end + sweep planner, checked against the oracle through
the NumPy plan interpreter (tests/plan_emulator.py)."""
import math

import numpy as np
import pytest

from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq
from queasars_b200 import gate_list as gl
from queasars_b200 import schedule as sc
from queasars_b200.circuit import CU3Gate, Parameter, QuantumCircuit, circuit_to_gate
from tests.plan_emulator import run_program


def build_circuit(instructions, n):
    """oracle-style instruction list -> queasars_b200 QuantumCircuit (named params become Parameters)."""
    circ = QuantumCircuit(n)
    cache = {}

    def conv(p):
        if isinstance(p, str):
            return cache.setdefault(p, Parameter(p))
        if isinstance(p, tuple):
            return cache.setdefault(p[1], Parameter(p[1])) * p[0] + p[2]
        return p

    for name, qubits, params in instructions:
        params = [conv(p) for p in params]
        if name == "cu3":
            circ.append(CU3Gate(*params), qubits)
        elif name in ("u", "u3"):
            circ.u(*params, qubits[0])
        else:
            circ._std(name, list(qubits), params)
    return circ


def emulate(gates: gl.GateList, values, k=sc.TILE_BITS, r=4, low=4):
    plan = sc.plan_circuit(gates.ops, gates.n_qubits, k, r, low)
    enc = sc.encode_plan(plan, gates.ops)
    state = run_program(enc, plan.n_eff, k, list(values), reg_bits=r)
    assert np.allclose(state[1 << gates.n_qubits :], 0)
    return state[: 1 << gates.n_qubits], plan


@pytest.mark.parametrize("n,layers,seed", [(4, 2, 0), (6, 3, 1), (9, 4, 2), (12, 3, 3), (14, 2, 4)])
def test_evqe_circuits_through_planner(n, layers, seed):
    genome, values = og.random_individual(n, layers, True, seed)
    instr = og.individual_circuit(genome, values)
    circ = build_circuit(instr, n)
    assert [p.name for p in circ.parameters] == oq.parameter_names(instr)
    gates = gl.from_circuit(circ)
    assert gates.n_params == len(values)
    want = oq.statevector(instr, n, values)
    got, plan = emulate(gates, values)
    np.testing.assert_allclose(got, want, atol=1e-13)


@pytest.mark.parametrize("k,r,low", [(6, 4, 2), (7, 4, 3), (8, 4, 4), (9, 4, 3)])
@pytest.mark.parametrize("n,layers,seed", [(8, 3, 10), (10, 4, 11), (11, 2, 12)])
def test_small_tiles_exercise_multi_tile_paths(k, r, low, n, layers, seed):
    genome, values = og.random_individual(n, layers, True, seed)
    instr = og.individual_circuit(genome, values)
    gates = gl.from_circuit(build_circuit(instr, n))
    want = oq.statevector(instr, n, values)
    got, plan = emulate(gates, values, k, r, low)
    np.testing.assert_allclose(got, want, atol=1e-13)
    kinds = {(po.ctrl_kind) for s in plan.sweeps for p in s.passes for po in p.ops}
    assert len(plan.sweeps) >= 1 and kinds  # program is non-trivial


def test_partial_parameterisation_and_direct_genome_path(genome_golden):
    for entry in genome_golden["population_12q_3l_seed11"]:
        genome = tuple(tuple(tuple(g) for g in layer) for layer in entry["layers"])
        values = entry["parameter_values"]
        lid = entry["partial_layers"][0]
        instr = og.individual_circuit(genome, values, {lid})
        layer_values = values[og.layer_value_slice(genome, lid)]
        assert layer_values == entry["partial_layer_values"]
        want = oq.statevector(instr, 12, layer_values)
        gates = gl.from_circuit(build_circuit(instr, 12))
        got, _ = emulate(gates, layer_values)
        np.testing.assert_allclose(got, want, atol=1e-13)
        # fully parameterised circuit with all stored values reproduces the same state (selection path)
        full = og.individual_circuit(genome, values)
        got_full, _ = emulate(gl.from_circuit(build_circuit(full, 12)), values)
        np.testing.assert_allclose(got_full, oq.statevector(full, 12, values), atol=1e-13)


class _FakeGate:
    def __init__(self, qubit_index, n_params, control=None):
        self.qubit_index = qubit_index
        self._n = n_params
        if control is not None:
            self.control_qubit_index = control

    def n_parameters(self):
        return self._n


def _fake_individual(genome, values):
    """duck-typed EVQEIndividual (the real class needs /root/reference)."""

    def mk(n_params):
        return type("G", (_FakeGate,), {"n_parameters": staticmethod(lambda n=n_params: n)})

    layers = []
    for layer in genome:
        gates = []
        for q, gene in enumerate(layer):
            if gene[0] == "rot":
                gates.append(mk(3)(q, 3))
            elif gene[0] == "crot":
                gates.append(mk(3)(q, 3, gene[1]))
            else:
                gates.append(mk(0)(q, 0))
        layers.append(type("L", (), {"gates": tuple(gates)})())
    idx, off = {}, 0
    for i, layer in enumerate(genome):
        npar = og.layer_n_params(layer)
        idx[i] = tuple(range(off, off + npar))
        off += npar
    return type("I", (), {"layers": tuple(layers), "n_qubits": len(genome[0]), "parameter_values": tuple(values), "layer_parameter_indices": idx})()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_genome_direct_path_equals_circuit_path(seed):
    n = 11
    genome, values = og.random_individual(n, 3, True, seed)
    ind = _fake_individual(genome, values)
    for layers in (None, {-1}, {0}, {1, 2}):
        instr = og.individual_circuit(genome, values, layers)
        via_circuit = gl.from_circuit(build_circuit(instr, n))
        direct = gl.from_evqe_individual(ind, layers)
        assert direct.param_names == via_circuit.param_names
        assert direct.ops == [op for op in via_circuit.ops]


def test_transpiled_basis_gates_and_expressions():
    n = 5
    a, b = Parameter("a"), Parameter("b")
    circ = QuantumCircuit(n)
    circ.h(0), circ.sx(1), circ.rz(a, 1), circ.cx(0, 1), circ.rzz(b * 2 + 0.3, 1, 3), circ.ecr(2, 4)
    circ.p(-a, 2), circ.cz(2, 3), circ.swap(0, 4), circ.ry(0.4, 3), circ.rx(b, 0), circ.t(1), circ.sdg(2)
    circ.cp(0.7, 4, 1), circ.crz(a * 0.5, 3, 0), circ.cu(0.1, 0.2, 0.3, 0.4, 1, 2), circ.rzx(0.9, 0, 2), circ.rxx(a, 3, 4)
    circ.y(0), circ.z(1), circ.x(2), circ.s(3), circ.tdg(4), circ.barrier(), circ.id(0)
    instr = []
    for inst in circ.data:
        ps = []
        for p in inst.operation.params:
            if hasattr(p, "parameters") and p.parameters:
                (prm,) = p.parameters
                ps.append((p._terms[prm], prm.name, p._const))
            else:
                ps.append(float(p))
        instr.append((inst.operation.name, tuple(q._index for q in inst.qubits), tuple(ps)))
    values = [0.37, -1.2]
    want = oq.statevector(instr, n, values)
    gates = gl.from_circuit(circ)
    assert gates.param_names == ("a", "b")
    got, _ = emulate(gates, values)
    np.testing.assert_allclose(got, want, atol=1e-13)  # exact including global phase


def test_composite_gate_decompose_and_compose():
    layer = QuantumCircuit(3, name="layer_0")
    t0, t1 = Parameter("layer0_q0_theta"), Parameter("layer0_q2_theta")
    layer.u(t0, 0.1, 0.2, 0)
    layer.append(CU3Gate(t1, 0.3, 0.4), (1, 2))
    outer = QuantumCircuit(3)
    outer.append(circuit_to_gate(layer), range(3))
    flat = outer.decompose()
    assert [i.operation.name for i in flat.data] == ["u", "cu3"]
    g1 = gl.from_circuit(outer)  # expands through .definition
    g2 = gl.from_circuit(flat)
    assert g1.ops == g2.ops and g1.param_names == ("layer0_q0_theta", "layer0_q2_theta")
    init = QuantumCircuit(3)
    init.h(0), init.h(1), init.h(2)
    composed = init.compose(flat, inplace=False)
    assert len(init.data) == 3 and len(composed.data) == 5
    measured = composed.measure_all(inplace=False)
    assert len(gl.from_circuit(measured).ops) == len(gl.from_circuit(composed).ops)
    bound = flat.assign_parameters([0.5, 0.6])
    assert not bound.parameters and gl.from_circuit(bound).n_params == 0


def test_planner_invariants():
    genome, values = og.random_individual(20, 6, True, 3)
    gates = gl.from_circuit(build_circuit(og.individual_circuit(genome, values), 20))
    plan = sc.plan_circuit(gates.ops, 20)
    seen = []
    for sw in plan.sweeps:
        assert len(sw.tile_qubits) == sc.TILE_BITS and sw.tile_qubits[:4] == [0, 1, 2, 3]
        if sc.ALLOW_LOW_EDGE_PASSES:  # no pass exists only to re-lay the tile out for the HBM access
            assert all(ps.ops for ps in sw.passes) or len(sw.passes) == 1
        else:  # first / last pass keep the low tile bits on the lanes (fully coalesced HBM accesses)
            assert not any(b < 4 for b in sw.passes[0].reg_bits)
            assert not any(b < 4 for b in sw.passes[-1].reg_bits)
        for ps in sw.passes:
            assert len(set(ps.reg_bits)) == 4
            seen += [po.op_index for po in ps.ops]
    absorbed = [i for i in plan.init_ops if i >= 0]
    init_ops, remaining = sc.split_product_prefix(gates.ops, 20)
    assert sorted(seen) == remaining and sorted(absorbed) == sorted(i for i in init_ops if i >= 0)
    dropped = set(range(len(gates.ops))) - set(seen) - set(absorbed)
    assert all(gates.ops[i].control >= 0 for i in dropped)  # only controlled gates on a |0> control vanish
    assert len(absorbed) >= 5 and dropped
    full = sc.plan_circuit(gates.ops, 20, product_prefix=False)
    assert sorted(po.op_index for sw in full.sweeps for ps in sw.passes for po in ps.ops) == list(range(len(gates.ops)))
    # a 20-qubit layer needs >= 2 sweeps (16 non-low qubits, 8 per tile); stay close to that bound
    assert len(plan.sweeps) <= 2 * 6 + 2
    plan12 = sc.plan_circuit(gates.ops, 20, tile_bits=12)
    assert len(plan12.sweeps) <= len(plan.sweeps)


def test_pipeline_split_point_fills_whole_waves():
    """Host logic of the pipelined submission (engine.pipeline_split_point; used on the NumPy conversion path only): 32
    twenty-qubit evaluations on 148 SMs (64 sweep CTAs per state at 8 tiles per CTA, 592 resident) are cut into 6 + 26 (the
    second chunk fills 2.81 waves); small lists and large states are not split."""
    from queasars_b200.engine import pipeline_split_point

    assert pipeline_split_point(32, 20, 11, 148) == 6
    assert pipeline_split_point(7, 20, 11, 148) is None
    assert pipeline_split_point(32, 26, 11, 148) is None
    for n in (8, 16, 24, 40, 64):
        c = pipeline_split_point(n, 18, 11, 148)
        assert n // 5 <= c <= n // 2 or c == 2


def test_diagonal_energy_bitstring_evaluator_contract():
    """DiagonalEnergyBitstringEvaluator keeps the BitstringEvaluator contract (bitstring_evaluation.py:7-61 of the reference:
    length / charset validation, one float per string) and agrees with the oracle's diagonal energy; duplicate masks merge."""
    import pickle

    from queasars_b200 import BitstringEvaluatorException, DiagonalEnergyBitstringEvaluator

    z, c = [1, 3, 8, 1, 5], [0.5, -1.0, 2.0, 0.25, 0.125]
    ev = DiagonalEnergyBitstringEvaluator(4, z, c)
    assert ev.input_length == 4
    for state in range(16):
        want = oq.diagonal_energy(state, list(zip(z, c)))
        assert ev.evaluate_bitstring(format(state, "04b")) == pytest.approx(want, abs=1e-15)
    np.testing.assert_allclose(ev.evaluate_states(np.arange(16, dtype=np.uint64)), [oq.diagonal_energy(s, list(zip(z, c))) for s in range(16)], atol=1e-15)
    assert len(ev.diagonal_terms[0]) == 4  # the two terms on mask 1 were merged
    with pytest.raises(BitstringEvaluatorException):
        ev.evaluate_bitstring("01")
    with pytest.raises(BitstringEvaluatorException):
        ev.evaluate_bitstring("01a1")
    clone = pickle.loads(pickle.dumps(ev))
    assert clone.evaluate_bitstring("1010") == ev.evaluate_bitstring("1010")


# ------------------------------------------------------------------------------------ phase deferral (gate_list.defer_phases)
def _apply_ops(ops, n, params):
    state = np.zeros(1 << n, dtype=complex)
    state[0] = 1.0
    for op in ops:
        m = np.array(op.matrix(params))
        if op.control < 0:
            state = oq.apply_matrix(state, n, m, [op.target])
        else:
            cm = np.kron(np.eye(2), np.diag([1, 0])) + np.kron(m, np.diag([0, 1]))
            state = oq.apply_matrix(state, n, cm, [op.control, op.target])
    return state


def test_phase_deferral_is_exact_on_random_op_lists():
    """u = D(phi) R_Y(theta) D(lam): deferring the trailing D(phi) of every uncontrolled gate (absorbed by the next gate on the
    qubit, passed through controlled gates, re-applied at the end) leaves the state unchanged; dropping the final phases leaves
    |psi_k|^2 unchanged; every uncontrolled dense gate ends up with phi == 0 (a real first column); angles stay affine in at
    most two parameters."""
    import random

    rng = random.Random(3)
    n, n_params = 5, 12

    def angle():
        return gl.Angle(rng.randrange(n_params), rng.uniform(-2, 2), rng.uniform(-3, 3)) if rng.random() < 0.7 else gl.const(rng.uniform(-3, 3))

    for _ in range(40):
        ops = []
        for _ in range(25):
            kind = rng.random()
            t = rng.randrange(n)
            c = rng.choice([q for q in range(n) if q != t])
            gamma = gl.ZERO if rng.random() < 0.8 else angle()
            if kind < 0.45:
                ops.append(gl.KernelOp(gl.DENSE, t, -1, gamma, angle(), angle(), angle()))
            elif kind < 0.8:
                ops.append(gl.KernelOp(gl.DENSE, t, c, gamma, angle(), angle(), angle()))
            elif kind < 0.9:
                ops.append(gl.KernelOp(gl.DIAG, t, -1, angle(), gl.ZERO, gl.ZERO, angle()))
            else:
                ops.append(gl.KernelOp(gl.DIAG, t, c, angle(), gl.ZERO, gl.ZERO, angle()))
        params = [rng.uniform(0, 6.28) for _ in range(n_params)]
        want = _apply_ops(ops, n, params)
        kept = gl.defer_phases(ops)
        dropped = gl.defer_phases(ops, drop_final=True)
        np.testing.assert_allclose(_apply_ops(kept, n, params), want, atol=1e-13)
        np.testing.assert_allclose(np.abs(_apply_ops(dropped, n, params)) ** 2, np.abs(want) ** 2, atol=1e-13)
        assert all(op.phi.is_zero for op in dropped if op.kind == gl.DENSE and op.control < 0)
        assert len(dropped) <= len(ops) <= len(kept)


@pytest.mark.parametrize("k,r,low", [(6, 4, 2), (8, 4, 4), (7, 3, 3), (sc.TILE_BITS, 4, 4)])
@pytest.mark.parametrize("n,layers,seed", [(6, 4, 20), (9, 5, 21), (12, 4, 22)])
def test_deferred_plans_keep_the_probabilities(k, r, low, n, layers, seed):
    """What the evaluators plan for diagonal observables / sampling (engine.rewritten with drop_final_phases): the planned and
    encoded program (two-term affine angles, qb_op_angles.slot2 / coeff2) must reproduce the oracle's |psi_k|^2 in the plan
    emulator; without the flag the gate list is left as written."""
    from queasars_b200.engine import rewritten

    genome, values = og.random_individual(n, layers, True, seed)
    instr = og.individual_circuit(genome, values)
    gates = gl.from_circuit(build_circuit(instr, n))
    assert rewritten(gates, drop_final_phases=False) is gates
    deferred = rewritten(gates, drop_final_phases=True)
    assert any(a.slot2 >= 0 for op in deferred.ops for a in op.angles)  # some gate absorbed a parameterised phase
    want = np.abs(oq.statevector(instr, n, values)) ** 2
    got, plan = emulate(deferred, values, k, r, low)
    np.testing.assert_allclose(np.abs(got) ** 2, want, atol=1e-13)
    from queasars_b200.gate_list import dfma_per_amplitude

    assert sum(dfma_per_amplitude(op) for op in deferred.ops) < sum(dfma_per_amplitude(op) for op in gates.ops)
