"""Host logic of queasars_b200.optimizers (the Optimizer calling convention of mutation.py:77-84) on analytic objectives."""
import numpy as np

from queasars_b200 import qiskit_compat
from queasars_b200.optimizers import NFT, SPSA


def test_nft_minimises_separable_sinusoid():
    qiskit_compat.install()
    shifts = np.array([0.3, -1.1, 2.0, 0.7])
    calls = []

    def fun(x):
        x = np.asarray(x).reshape(-1, 4)
        calls.append(len(x))
        vals = np.sum(np.cos(x - shifts), axis=1)
        return vals[0] if len(vals) == 1 else vals

    res = NFT(maxfev=60).minimize(fun=fun, x0=np.zeros(4), bounds=[(None, None)] * 4)
    assert res.fun < -3.999 and res.nfev == sum(calls)
    grouped = NFT(maxfev=60)
    grouped.set_max_evals_grouped(2)
    calls.clear()
    res2 = grouped.minimize(fun=fun, x0=np.zeros(4), bounds=None)
    assert res2.fun < -3.999 and max(calls) == 2  # the +-pi/2 pair arrives as one batch of two


def test_spsa_decreases_and_honours_termination_checker():
    qiskit_compat.install()
    from qiskit_algorithms.utils import algorithm_globals

    algorithm_globals.random_seed = 7
    target = np.linspace(-1, 1, 6)

    def fun(x):
        x = np.asarray(x).reshape(-1, 6)
        vals = np.sum((x - target) ** 2, axis=1)
        return vals[0] if len(vals) == 1 else vals

    seen = []

    def checker(nfev, x, fx, step, accepted):
        seen.append(nfev)
        return len(seen) >= 25

    opt = SPSA(maxiter=200, learning_rate=0.1, perturbation=0.1, termination_checker=checker)
    opt.set_max_evals_grouped(2)
    res = opt.minimize(fun=fun, x0=np.zeros(6), bounds=None)
    assert res.nit == 25 and res.nfev == 2 * 25 + 1
    assert res.fun < fun(np.zeros(6))
