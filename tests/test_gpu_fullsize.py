"""Full-size GPU parity: the CUDA path against the C/OpenMP oracle (oracle/c/statevector.c, pinned to the NumPy restatement
in tests/test_c_oracle.py) on ENTANGLED EVQE states at BASELINE.json's sizes --

  C2  20 qubits x 32 individuals x 6 layers, random diagonal Ising, through ``B200OperatorCircuitEvaluator``   (1e-10)
  C3  24-qubit transverse-field Ising Pauli sum (23 ZZ + 24 X), fp64                                            (1e-10)
  C4  26-qubit JSSP QUBO, 10 000 shots: sampled indices vs cumsum -> searchsorted(right), CVaR value
                                                                    (<= 1 boundary flip per 10^4 shots, to an adjacent state)

and every compiled sweep-kernel variant at one full size each: 2^12-amplitude tiles, 2^3 amplitudes per thread, complex64
(1e-4), and the 64-bit-index kernels that > 31 local qubits (the shards of C5) run, forced on at 20 / 24 qubits.
Reference call sites: circuit_evaluation.py:200-215 (estimator route), :29-59 (sampler route), expectation_calculation.py:35-69.
"""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq
from queasars_b200 import gate_list as gl
from queasars_b200.operators import SparsePauliOp
from tests.test_frontend_planner import build_circuit
from tests.test_gpu_parity import random_ising, rel_err, tfim

pytestmark = pytest.mark.gpu


def evqe_case(n, layers, seed):
    genome, values = og.random_individual(n, layers, True, seed)
    instr = og.individual_circuit(genome, values)
    return instr, list(values), build_circuit(instr, n)


@pytest.fixture(scope="module")
def engine():
    from queasars_b200.engine import Engine

    return Engine(device=0, dtype="complex128")


@pytest.fixture(scope="module")
def ising20():
    terms = random_ising(20)
    table = c_oracle.diag_table(20, oq.diag_terms_from_labels(terms))
    return terms, table


# ------------------------------------------------------------------------------------ C2
def test_c2_population_of_32_through_the_evaluator(ising20):
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator

    terms, table = ising20
    n = 20
    cases = [evqe_case(n, 6, 1000 + s) for s in range(32)]
    evaluator = B200OperatorCircuitEvaluator(B200EstimatorV2(device=0), 0.0, SparsePauliOp.from_list(terms))
    got = evaluator.evaluate_circuits([c for _, _, c in cases], [v for _, v, _ in cases])
    assert len(got) == 32
    state = np.empty(1 << n, dtype=np.complex128)
    for g, (instr, values, _) in zip(got, cases):
        want, _ = c_oracle.evaluate(instr, n, values, table, state)
        assert rel_err(g, want) < 1e-10


# ------------------------------------------------------------------------------------ C3
def test_c3_24q_tfim_entangled_state(engine):
    n = 24
    terms = tfim(n)
    ham = engine.hamiltonian(SparsePauliOp.from_list(terms))
    state = np.empty(1 << n, dtype=np.complex128)
    plans, params, want = [], [], []
    for seed in (3, 4):
        instr, values, circ = evqe_case(n, 6, seed)
        plans.append(engine.compile(gl.from_circuit(circ)))
        params.append(values)
        c_oracle.evaluate(instr, n, values, None, state)
        want.append(c_oracle.pauli_sum(state, n, terms))
    got = engine.expectation(plans, params, ham)
    for g, w in zip(got, want):
        assert rel_err(g, w) < 1e-10
    # the amplitudes themselves, for the last circuit (the oracle's state is still in `state`)
    sv = engine.statevector(plans[-1], params[-1])
    assert np.max(np.abs(sv - state)) < 1e-13
    # mixed X/Y/Z strings with complex coefficients on the same entangled state (generic two-read kernel + tile kernel)
    rng = np.random.default_rng(24)
    mixed = [("".join(rng.choice(list("IXYZ"), size=n, p=[0.7, 0.1, 0.1, 0.1])), complex(rng.normal(), rng.normal())) for _ in range(4)]
    ham2 = engine.hamiltonian(SparsePauliOp.from_list(mixed))
    assert rel_err(engine.expectation([plans[-1]], [params[-1]], ham2)[0], c_oracle.pauli_sum(state, n, mixed)) < 1e-10


# ------------------------------------------------------------------------------------ C4
def test_c4_26q_jssp_sampler_entangled_state(engine, jssp_golden):
    from queasars_b200 import B200OperatorSamplerCircuitEvaluator, B200SamplerV2

    entry = jssp_golden["jssp_26q"]
    n, shots, seed = 26, 10000, 17
    diag_terms = list(zip(entry["z_masks"], entry["coeffs"]))
    op = SparsePauliOp._raw(n, [0] * len(diag_terms), entry["z_masks"], entry["coeffs"])
    instr, values, circ = evqe_case(n, 4, 26)
    state = np.empty(1 << n, dtype=np.complex128)
    c_oracle.evaluate(instr, n, values, None, state)
    uniforms = np.random.default_rng(seed).random(shots)
    want_idx = c_oracle.sample_indices(state, n, uniforms)
    plan = engine.compile(gl.from_circuit(circ))
    got_idx = engine.sample([plan], [values], shots, uniforms.reshape(1, -1))[0]
    flips = np.nonzero(got_idx != want_idx)[0]
    assert len(flips) <= 1, f"{len(flips)} of {shots} sampled indices differ from cumsum -> searchsorted(right)"
    probs = state.real**2 + state.imag**2
    for i in flips:  # a uniform within rounding of a CDF boundary: the two draws are CDF neighbours (no probability mass between them)
        lo, hi = sorted((int(got_idx[i]), int(want_idx[i])))
        assert probs[got_idx[i]] > 0 and float(np.sum(probs[lo + 1 : hi])) < 1e-10
    # the evaluator route (B200SamplerV2(seed) draws default_rng(seed).random(shots) itself): mean and CVaR of the sampled energies
    sampler = B200SamplerV2(device=0, seed=seed)
    for alpha in (1.0, 0.5):
        ev = B200OperatorSamplerCircuitEvaluator(sampler, shots, op, alpha=alpha)
        got = ev.evaluate_circuits([circ], [values])[0]
        dist_gpu = oq.quasi_distribution(oq.counts_from_indices(got_idx, n), shots)
        assert got == pytest.approx(oq.expectation_with_operator(dist_gpu, diag_terms, alpha), rel=1e-10, abs=1e-9)
        if not len(flips):
            dist = oq.quasi_distribution(oq.counts_from_indices(want_idx, n), shots)
            assert got == pytest.approx(oq.expectation_with_operator(dist, diag_terms, alpha), rel=1e-10, abs=1e-9)


# ------------------------------------------------------------------------------------ kernel variants at full size
def _variant_engine(kind):
    from queasars_b200.engine import Engine

    if kind == "tile12":
        return Engine(device=0, tile_bits=12)
    if kind == "reg3":
        return Engine(device=0, reg_bits=3)
    if kind == "idx64":
        eng = Engine(device=0)
        eng.set_index_width(64)
        return eng
    if kind == "c64":
        return Engine(device=0, dtype="complex64")
    if kind == "c64_idx64":
        eng = Engine(device=0, dtype="complex64")
        eng.set_index_width(64)
        return eng
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["tile12", "reg3", "idx64", "c64", "c64_idx64"])
def test_sweep_kernel_variants_20q_vs_c_oracle(kind, ising20):
    terms, table = ising20
    n = 20
    eng = _variant_engine(kind)
    single = kind.startswith("c64")
    ham = eng.hamiltonian(SparsePauliOp.from_list(terms))
    ham_x = eng.hamiltonian(SparsePauliOp.from_list(tfim(n)))
    state = np.empty(1 << n, dtype=np.complex128)
    plans, plans_prob, params, want, want_x = [], [], [], [], []
    for seed in (11, 12, 13):
        instr, values, circ = evqe_case(n, 5, seed)
        plans.append(eng.compile(gl.from_circuit(circ)))
        # what the evaluators compile for diagonal observables / sampling: trailing phases deferred (REAL10 gate bodies)
        plans_prob.append(eng.compile(gl.from_circuit(circ), drop_final_phases=True))
        params.append(values)
        value, _ = c_oracle.evaluate(instr, n, values, table, state)
        want.append(value)
        want_x.append(c_oracle.pauli_sum(state, n, tfim(n)))
    tol = 1e-4 if single else 1e-10
    for g, w in zip(eng.expectation(plans, params, ham), want):
        assert rel_err(g, w) < tol
    for g, w in zip(eng.expectation(plans_prob, params, ham), want):
        assert rel_err(g, w) < tol
    for g, w in zip(eng.expectation(plans, params, ham_x), want_x):
        assert rel_err(g, w) < tol
    sv = eng.statevector(plans[-1], params[-1])
    assert np.max(np.abs(sv - state)) < (2e-6 if single else 1e-13)
    # sampler on the same variant: identical uniforms -> identical indices (complex64 probabilities differ at 1e-7 relative:
    # a handful of boundary flips to neighbouring states are expected there)
    shots = 4096
    uniforms = np.random.default_rng(5).random(shots)
    got_idx = eng.sample([plans_prob[-1]], [params[-1]], shots, uniforms.reshape(1, -1))[0]
    want_idx = c_oracle.sample_indices(state, n, uniforms)
    if single:
        assert np.count_nonzero(np.abs(got_idx - want_idx) > 64) <= shots // 100
    else:
        assert np.count_nonzero(got_idx != want_idx) <= 1
    eng.close()


def test_idx64_kernels_24q_tfim(engine):
    """The uint64-index sweep + expectation path (what a 32-qubit shard of C5 runs) on a 24-qubit entangled state, against both
    the C oracle and the 32-bit-index engine."""
    n = 24
    eng = _variant_engine("idx64")
    terms = tfim(n)
    instr, values, circ = evqe_case(n, 4, 9)
    state = np.empty(1 << n, dtype=np.complex128)
    c_oracle.evaluate(instr, n, values, None, state)
    want = c_oracle.pauli_sum(state, n, terms)
    got64 = eng.expectation([eng.compile(gl.from_circuit(circ))], [values], eng.hamiltonian(SparsePauliOp.from_list(terms)))[0]
    got32 = engine.expectation([engine.compile(gl.from_circuit(circ))], [values], engine.hamiltonian(SparsePauliOp.from_list(terms)))[0]
    assert rel_err(got64, want) < 1e-10
    assert got64 == got32  # same program, same arithmetic: bit-identical
    eng.close()
