"""ONE process driving a statevector sharded over several GPUs (``LocalShardedStatevector``: what the evaluators use for circuits
too wide for one GPU -- BASELINE config C5 behind ``evaluate_circuits``) against the oracle at sizes it can check.  On a one-GPU
box the shards are "virtual" (several shards on the same device: same kernels, same peer stores, same host logic); with more
GPUs ``devices="all"`` puts one shard on each.  The evaluator route is forced with ``shard_min_qubits``."""
import math

import numpy as np
import pytest

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


def shard_engines(count):
    """``count`` shards over the visible GPUs (round robin): one per GPU when there are enough, virtual shards otherwise."""
    from queasars_b200 import _native
    from queasars_b200.primitives import get_engine

    n_dev = _native.device_count()
    return [get_engine(r % n_dev) for r in range(count)]


@pytest.mark.parametrize("n,shards,layers", [(14, 2, 3), (16, 4, 3), (17, 8, 2)])
def test_local_sharded_state_matches_oracle(n, shards, layers):
    from queasars_b200.sharded import LocalShardedStatevector

    instr, values, circ = evqe_case(n, layers, 40 + n)
    sv = LocalShardedStatevector(n, shard_engines(shards), min_local=8)
    sv.run(gl.from_circuit(circ), values)
    want = oq.statevector(instr, n, values)
    assert sv.swaps_done >= 1  # the circuit targets the top qubits: at least one global swap happened
    assert np.max(np.abs(sv.gather_logical() - want)) < 1e-13
    assert abs(sv.norm_squared() - 1.0) < 1e-12
    rng = np.random.default_rng(n)
    z = [int(v) for v in rng.integers(0, 1 << n, size=9)]
    c = [float(v) for v in rng.normal(size=9)]
    assert rel_err(sv.diagonal_expectation(z, c), float(np.dot(np.abs(want) ** 2, oq.diagonal_table(n, list(zip(z, c)))))) < 1e-10
    terms = tfim(n) + [("Y" + "I" * (n - 2) + "X", 0.3 - 0.2j)]  # X / Y on the top qubit: flips a rank bit -> one more swap
    assert rel_err(sv.expectation(SparsePauliOp.from_list(terms)), oq.estimator_expectation(want, terms)) < 1e-10
    # sampling: per-shot inverse CDF in PHYSICAL enumeration order, mapped back to logical indices
    shots = 4000
    uniforms = np.random.default_rng(3).random(shots)
    got = sv.sample(shots, uniforms=uniforms)
    physical = sv.gather_physical()
    idx_phys = oq.sample_indices(physical, shots, uniforms=uniforms)
    logical = np.zeros_like(idx_phys)
    for q in range(n):
        logical |= ((idx_phys >> sv.position[q]) & 1) << q
    assert np.count_nonzero(got != logical) <= 1
    assert sv.describe()["swap_path"].startswith("swap_p2p_kernel")
    # reset: a second circuit on the same buffers
    instr2, values2, circ2 = evqe_case(n, 2, 90 + n)
    sv.reset()
    sv.run(gl.from_circuit(circ2), values2)
    assert np.max(np.abs(sv.gather_logical() - oq.statevector(instr2, n, values2))) < 1e-13
    sv.close()


def test_evaluators_route_wide_circuits_to_the_sharded_state():
    """``evaluate_circuits`` on an operator wider than ``shard_min_qubits``: the primitive shards ONE state over its device set
    (here: 4 shards) instead of batching -- same values as the single-GPU engine and the oracle."""
    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator, B200OperatorSamplerCircuitEvaluator, B200SamplerV2

    n = 15
    terms = tfim(n) + random_ising(n, 4)[:20]
    op = SparsePauliOp.from_list(terms)
    cases = [evqe_case(n, 3, 60 + s) for s in range(3)]
    circuits, params = [c for _, _, c in cases], [v for _, v, _ in cases]
    est = B200EstimatorV2(devices="all", coalesce=False, shard_min_qubits=14)
    est._engines_obj = shard_engines(4)  # 4 shards whatever the number of visible GPUs
    got = B200OperatorCircuitEvaluator(est, 0.0, op).evaluate_circuits(circuits, params)
    ref = B200OperatorCircuitEvaluator(B200EstimatorV2(device=0, coalesce=False), 0.0, op).evaluate_circuits(circuits, params)
    for g, r, (instr, values, _) in zip(got, ref, cases):
        assert rel_err(g, oq.estimator_expectation(oq.statevector(instr, n, values), terms)) < 1e-10
        assert rel_err(g, r) < 1e-12
    assert est._sharded_obj[n].world == 4 and est._sharded_obj[n].swaps_done >= 1
    # sampler route: distribution of the sharded draws against the exact probabilities, and the CVaR evaluator on top
    diag = SparsePauliOp.from_list(random_ising(n, 6))
    smp = B200SamplerV2(devices="all", seed=5, coalesce=False, shard_min_qubits=14)
    smp._engines_obj = shard_engines(4)
    shots = 20000
    idx = smp.sample_indices([circuits[0]], [params[0]], shots)[0]
    probs = np.abs(oq.statevector(cases[0][0], n, params[0])) ** 2
    emp = np.bincount(idx, minlength=1 << n) / shots
    bound = 1.5 * 0.5 * np.sum(np.sqrt(2 * probs * (1 - probs) / (math.pi * shots)))
    assert 0.5 * np.abs(emp - probs).sum() <= bound
    val = B200OperatorSamplerCircuitEvaluator(smp, shots, diag, alpha=1.0).evaluate_circuits([circuits[0]], [params[0]])[0]
    exact = float(np.dot(probs, oq.diagonal_table(n, oq.diag_terms_from_labels(random_ising(n, 6)))))
    energies = oq.diagonal_table(n, oq.diag_terms_from_labels(random_ising(n, 6)))
    sigma = math.sqrt(float(np.dot(probs, (energies - exact) ** 2)))
    assert abs(val - exact) < 5 * sigma / math.sqrt(shots)


def test_too_wide_without_a_device_set_raises_clearly():
    from queasars_b200 import B200EstimatorV2

    est = B200EstimatorV2(device=0)
    assert not est._needs_sharding(30) and est._needs_sharding(36)
    with pytest.raises(ValueError, match="does not fit one GPU"):
        est._sharded_state(36)
