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
