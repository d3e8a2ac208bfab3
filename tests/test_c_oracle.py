"""The C/OpenMP oracle (CPU baseline) agrees with the NumPy oracle it restates."""
import numpy as np

from oracle import c_oracle
from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq


def test_c_oracle_matches_numpy_oracle():
    rng = np.random.default_rng(0)
    for n, layers, seed in [(3, 2, 0), (8, 4, 1), (13, 3, 2)]:
        genome, values = og.random_individual(n, layers, True, seed)
        instr = og.individual_circuit(genome, values)
        terms = [(int(rng.integers(0, 1 << n)), float(rng.normal())) for _ in range(9)]
        table_np = oq.diagonal_table(n, terms)
        table_c = c_oracle.diag_table(n, terms)
        np.testing.assert_allclose(table_c, table_np, rtol=0, atol=1e-12)
        value, state = c_oracle.evaluate(instr, n, values, table_c)
        want = oq.statevector(instr, n, values)
        np.testing.assert_allclose(state, want, atol=1e-13)
        assert abs(value - float(np.dot(np.abs(want) ** 2, table_np))) < 1e-11


def _tfim_terms(n, h=0.5):
    terms = []
    for i in range(n - 1):
        lab = ["I"] * n
        lab[n - 1 - i] = lab[n - 2 - i] = "Z"
        terms.append(("".join(lab), -1.0))
    for i in range(n):
        lab = ["I"] * n
        lab[n - 1 - i] = "X"
        terms.append(("".join(lab), -h))
    return terms


def test_c_pauli_sum_matches_numpy_oracle():
    """oracle_pauli_sum (one pass per term, x/z masks, i^{nY}) against oq.estimator_expectation, incl. Y strings and complex
    coefficients -- this is what the full-size (24-qubit) GPU parity test trusts."""
    rng = np.random.default_rng(5)
    for n, layers, seed in [(3, 2, 0), (9, 3, 1), (14, 3, 2)]:
        genome, values = og.random_individual(n, layers, True, seed)
        instr = og.individual_circuit(genome, values)
        state = oq.statevector(instr, n, values)
        terms = _tfim_terms(n)
        for _ in range(5):
            terms.append(("".join(rng.choice(list("IXYZ"), size=n)), complex(rng.normal(), rng.normal())))
        want = oq.estimator_expectation(state, terms)
        got = c_oracle.pauli_sum(state, n, terms)
        assert abs(got - want) <= 1e-12 * max(1.0, abs(want))


def test_c_sampler_is_numpy_cumsum_searchsorted():
    """oracle_sample = probs.cumsum(); cdf /= cdf[-1]; searchsorted(side='right') -- index for index identical to the NumPy
    restatement (sequential cumsum on both sides), also for uniforms at 0, just below 1 and exactly on CDF entries."""
    rng = np.random.default_rng(6)
    for n, layers, seed in [(4, 2, 3), (11, 3, 4), (16, 3, 5)]:
        genome, values = og.random_individual(n, layers, True, seed)
        state = oq.statevector(og.individual_circuit(genome, values), n, values)
        probs = state.real**2 + state.imag**2
        cdf = probs.cumsum()
        cdf /= cdf[-1]
        uniforms = np.concatenate([rng.random(5000), [0.0, np.nextafter(1.0, 0.0)], cdf[rng.integers(0, cdf.size - 1, 50)]])
        want = oq.sample_indices(state, uniforms.size, uniforms=uniforms)
        got = c_oracle.sample_indices(state, n, uniforms)
        assert np.array_equal(got, want)
