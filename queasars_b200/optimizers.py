"""Derivative-free optimizers with the ``qiskit_algorithms.optimizers.Optimizer`` calling convention the
EVQE mutation operators use: ``optimizer.minimize(fun=..., x0=..., bounds=...) -> result`` with
``result.x`` / ``result.nfev`` (/root/reference/queasars/minimum_eigensolvers/evqe/evolutionary_algorithm/
mutation.py:77-84), ``copy.deepcopy``-able, and a ``termination_checker(nfev, x, fx, stepsize, accepted)``
hook compatible with /root/reference/queasars/utility/spsa_termination.py.

Both optimizers support *grouped evaluation* (``set_max_evals_grouped``): the +/- perturbation pair of an
SPSA step, or the two shifted points of an NFT step, are handed to ``fun`` as ONE concatenated vector so the
evaluator receives a batch (the caller reshapes with ``reshape(-1, n_parameters)``: mutation.py:63-75).
That turns the sequential objective chain into 2-wide batches for the GPU.

Algorithms: SPSA after Spall (1998) with constant gain sequences, Bernoulli +/-1 perturbations, optional
trust region, last-iterate averaging; NFT after Nakanishi, Fujii, Todo (2020): exact one-parameter
sinusoidal minimisation in round-robin order with periodic re-evaluation of the anchor value.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Callable, Optional, Sequence

import numpy as np


class OptimizerResult:
    def __init__(self):
        self.x = None
        self.fun = None
        self.jac = None
        self.nfev = 0
        self.njev = None
        self.nit = 0


class Optimizer(ABC):
    def __init__(self):
        self._max_evals_grouped = 1

    def set_max_evals_grouped(self, limit: Optional[int]) -> None:
        self._max_evals_grouped = max(1, int(limit or 1))

    @abstractmethod
    def minimize(self, fun: Callable, x0, jac=None, bounds=None) -> OptimizerResult:
        ...

    def _evaluate_points(self, fun: Callable, points: Sequence[np.ndarray]) -> list[float]:
        """Evaluate several points, grouped into concatenated calls of at most ``_max_evals_grouped``."""
        out: list[float] = []
        group = self._max_evals_grouped
        i = 0
        while i < len(points):
            chunk = points[i : i + group]
            if len(chunk) == 1:
                out.append(float(fun(chunk[0])))
            else:
                vals = np.atleast_1d(np.asarray(fun(np.concatenate(chunk)), dtype=np.float64))
                out.extend(float(v) for v in vals)
            i += group
        return out


def _rng():
    try:  # honour the seed the reference sets (mutation.py:54-55) whichever module provides it
        from qiskit_algorithms.utils import algorithm_globals

        return algorithm_globals.random
    except Exception:  # pragma: no cover - stand-in not installed and no qiskit_algorithms
        return np.random.default_rng()


class SPSA(Optimizer):
    def __init__(
        self,
        maxiter: int = 100,
        blocking: bool = False,
        allowed_increase: Optional[float] = None,
        trust_region: bool = False,
        learning_rate: Optional[float] = None,
        perturbation: Optional[float] = None,
        last_avg: int = 1,
        resamplings: int = 1,
        callback: Optional[Callable] = None,
        termination_checker: Optional[Callable] = None,
    ):
        super().__init__()
        self.maxiter = maxiter
        self.blocking = blocking
        self.allowed_increase = allowed_increase
        self.trust_region = trust_region
        self.learning_rate = 0.2 if learning_rate is None else float(learning_rate)
        self.perturbation = 0.2 if perturbation is None else float(perturbation)
        self.last_avg = max(1, int(last_avg))
        self.resamplings = max(1, int(resamplings))
        self.callback = callback
        self.termination_checker = termination_checker

    def minimize(self, fun, x0, jac=None, bounds=None) -> OptimizerResult:
        rng = _rng()
        x = np.asarray(x0, dtype=np.float64).copy()
        dim = x.size
        nfev = 0
        eps, eta = self.perturbation, self.learning_rate
        fx = None
        if self.blocking:
            fx = float(fun(x))
            nfev += 1
            allowed = self.allowed_increase if self.allowed_increase is not None else 0.0
        tail: list[np.ndarray] = []
        nit = 0
        for k in range(1, self.maxiter + 1):
            nit = k
            grad = np.zeros(dim)
            estimate = 0.0
            for _ in range(self.resamplings):
                delta = 1.0 - 2.0 * rng.binomial(1, 0.5, size=dim)
                plus, minus = self._evaluate_points(fun, [x + eps * delta, x - eps * delta])
                nfev += 2
                grad += (plus - minus) / (2.0 * eps) * delta
                estimate += 0.5 * (plus + minus)
            grad /= self.resamplings
            estimate /= self.resamplings
            update = grad
            if self.trust_region:
                norm = float(np.linalg.norm(update))
                if norm > 1.0:
                    update = update / norm
            update = eta * update
            x_next = x - update
            fx_next = None
            accepted = True
            if self.blocking:
                fx_next = float(fun(x_next))
                nfev += 1
                if fx + allowed <= fx_next:
                    accepted = False
                else:
                    fx = fx_next
            if accepted:
                x = x_next
            if self.callback is not None:
                self.callback(nfev, x_next, estimate if fx_next is None else fx_next, float(np.linalg.norm(update)), accepted)
            if accepted:
                tail.append(x_next)
                tail = tail[-self.last_avg :]
            if self.termination_checker is not None:
                check = estimate if fx_next is None else fx_next
                if self.termination_checker(nfev, x_next, check, float(np.linalg.norm(update)), accepted):
                    break
        if self.last_avg > 1 and tail:
            x = np.mean(np.asarray(tail), axis=0)
        result = OptimizerResult()
        result.x = x
        result.fun = float(fun(x))
        result.nfev = nfev + 1
        result.nit = nit
        return result


class NFT(Optimizer):
    def __init__(self, maxiter: Optional[int] = None, maxfev: int = 1024, disp: bool = False, reset_interval: int = 32):
        super().__init__()
        self.maxiter = maxiter
        self.maxfev = maxfev
        self.reset_interval = reset_interval

    def minimize(self, fun, x0, jac=None, bounds=None) -> OptimizerResult:
        x = np.asarray(x0, dtype=np.float64).copy()
        n = x.size
        maxiter = self.maxiter if self.maxiter is not None else (self.maxfev if self.maxfev is not None else n * 2)
        tiny = 1e-32
        nfev = 0
        nit = 0
        z0 = float(fun(x))
        nfev += 1
        while n > 0:
            idx = nit % n
            if self.reset_interval > 0 and nit > 0 and nit % self.reset_interval == 0:
                z0 = float(fun(x))
                nfev += 1
            up, down = x.copy(), x.copy()
            up[idx] += np.pi / 2
            down[idx] -= np.pi / 2
            z1, z3 = self._evaluate_points(fun, [up, down])
            nfev += 2
            z2 = z1 + z3 - z0
            offset = (z1 + z3) / 2.0
            denom = (z0 - z2) + tiny * float(z0 == z2)
            amp = np.sqrt((z0 - z2) ** 2 + (z1 - z3) ** 2) / 2.0
            shift = np.arctan((z1 - z3) / denom) + x[idx]
            shift += 0.5 * np.pi + 0.5 * np.pi * np.sign(denom)
            x[idx] = shift
            z0 = offset - amp
            nit += 1
            if self.maxfev is not None and nfev >= self.maxfev:
                break
            if nit >= maxiter:
                break
        result = OptimizerResult()
        result.x = x
        result.fun = float(fun(x))
        result.nfev = nfev + 1
        result.nit = nit
        return result
