"""CPU oracle for the EVQE circuit-evaluation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``queasars_b200/`` may import this package; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs do.

Parity status: **parity unpinned at the Estimator/Sampler primitive boundary** -- the arithmetic the
reference delegates to (qiskit 2.4.2, qiskit-aer 0.17.2, qiskit-algorithms 0.4.0, pinned in
/root/reference/poetry.lock) is not vendored and not installable here, and the reference's own tests
pin no numeric expectation value or sampled distribution.  What *is* pinned (tests/golden/): the
diagonal-energy semantics and minima of the reference's JSSP encoder, the genome random generation
and the parameter naming, all produced by running the reference's own pure-Python modules in this
container (tests/golden/make_golden.py).
"""
