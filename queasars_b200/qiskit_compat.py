"""Stand-in ``qiskit`` / ``qiskit_algorithms`` / ``dask`` modules for environments without them.

``install()`` registers lightweight module objects in ``sys.modules`` -- ONLY for top-level packages that
are not importable -- exposing exactly the names the reference package imports
(/root/reference/queasars/**: ``qiskit.circuit``, ``qiskit.circuit.library.CU3Gate``,
``qiskit.converters.circuit_to_gate``, ``qiskit.primitives[.base|.containers.*|.primitive_job]``,
``qiskit.quantum_info[.operators[.base_operator]]``, ``qiskit.result``, ``qiskit.transpiler[...]``,
``qiskit.qpy``, ``qiskit_algorithms[...]``, ``dask.distributed``, ``dask.utils``).  With it the unmodified
reference package can be imported from a source checkout and driven end-to-end by the B200 primitives /
evaluators.  It is *not* a simulator: no arithmetic of the hot path lives here.
"""
from __future__ import annotations

import importlib.util
import sys
import threading
import types
from abc import ABC, abstractmethod
from concurrent.futures import Future, wait as _futures_wait
from typing import Any, Dict, List, Optional, TypeVar, Union

import numpy as np

from . import circuit as _circuit
from . import containers as _containers
from . import operators as _operators


_T = TypeVar("_T")
ListOrDict = Union[List[Optional[_T]], Dict[str, _T]]  # qiskit_algorithms.list_or_dict.ListOrDict


def _missing(name: str) -> bool:
    if name in sys.modules:
        return False
    try:
        return importlib.util.find_spec(name) is None
    except (ImportError, ValueError):
        return True


def _module(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    mod.__b200_stub__ = True
    sys.modules[name] = mod
    parent, _, child = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], child, mod)
    return mod


# ------------------------------------------------------------------------------------------ qiskit
class BaseOperator:  # qiskit.quantum_info.operators.base_operator.BaseOperator
    pass


class BaseEstimatorV2(ABC):
    @abstractmethod
    def run(self, pubs, *, precision=None):
        ...


class BaseSamplerV2(ABC):
    @abstractmethod
    def run(self, pubs, *, shots=None):
        ...


class PassManager:
    """Identity pass manager (``generate_preset_pass_manager(optimization_level=0)`` with no backend
    leaves ``u``/``cu3``/``id`` untouched upstream; base/evolving_ansatz_minimum_eigensolver.py:163-164)."""

    def __init__(self, passes=None):
        self.passes = passes or []

    def run(self, circuits, **_kw):
        return circuits


def generate_preset_pass_manager(optimization_level: int = 0, backend=None, **_kw) -> PassManager:
    return PassManager()


def sampled_expectation_value(dist, oper) -> float:
    """qiskit.result.sampled_expectation_value for a diagonal operator (host-side; small dists only)."""
    from .expectation import diagonal_energies

    if not isinstance(oper, _operators.SparsePauliOp):
        raise TypeError("stub sampled_expectation_value needs a SparsePauliOp")
    x, z, c = oper.masks()
    if np.any(x != 0):
        raise ValueError("Operator string contains non-diagonal terms")
    states = np.fromiter((int(k, 2) if isinstance(k, str) else int(k) for k in dist.keys()), dtype=np.uint64)
    probs = np.fromiter(dist.values(), dtype=np.float64)
    if getattr(dist, "shots", None) is None and not isinstance(dist, _containers.QuasiDistribution):
        probs = probs / probs.sum()
    return float(np.dot(probs, diagonal_energies(states, z, c.real)))


def _evaluate_sparsepauli(state: int, observable) -> complex:
    """qiskit_algorithms.minimum_eigensolvers.diagonal_estimator._evaluate_sparsepauli."""
    _, z, c = observable.masks()
    total = 0.0 + 0.0j
    for zm, coeff in zip(z.tolist(), c.tolist()):
        total += coeff * (1 - 2 * (bin(int(state) & int(zm)).count("1") & 1))
    return total


def _qpy_unavailable(*_a, **_k):
    raise NotImplementedError("qiskit.qpy is not available in the B200 qiskit stand-in")


# ------------------------------------------------------------------------------- qiskit_algorithms
class MinimumEigensolverResult:
    def __init__(self):
        self.eigenvalue = None
        self.aux_operators_evaluated = None


class MinimumEigensolver(ABC):
    @abstractmethod
    def compute_minimum_eigenvalue(self, operator, aux_operators=None):
        return MinimumEigensolverResult()

    @classmethod
    def supports_aux_operators(cls) -> bool:
        return False


class _AlgorithmGlobals:
    def __init__(self):
        self._seed: Optional[int] = None
        self._rng: Optional[np.random.Generator] = None

    @property
    def random_seed(self):
        return self._seed

    @random_seed.setter
    def random_seed(self, seed):
        self._seed = seed
        self._rng = None

    @property
    def random(self) -> np.random.Generator:
        if self._rng is None:
            self._rng = np.random.default_rng(self._seed)
        return self._rng


algorithm_globals = _AlgorithmGlobals()


# -------------------------------------------------------------------------------------------- dask
class SerializableLock:
    """dask.utils.SerializableLock: a lock that pickles (by token) and re-links inside one process."""

    _locks: dict = {}
    _guard = threading.Lock()

    def __init__(self, token: Optional[str] = None):
        self.token = token or f"lock-{id(self):x}"
        with SerializableLock._guard:
            self.lock = SerializableLock._locks.setdefault(self.token, threading.Lock())

    def acquire(self, *args, **kwargs):
        return self.lock.acquire(*args, **kwargs)

    def release(self):
        return self.lock.release()

    def locked(self):
        return self.lock.locked()

    def __enter__(self):
        self.lock.acquire()
        return self

    def __exit__(self, *exc):
        self.lock.release()

    def __getstate__(self):
        return self.token

    def __setstate__(self, token):
        self.__init__(token)


class Client:
    """dask.distributed.Client placeholder: only used by the reference for isinstance() dispatch."""

    def __init__(self, *_a, **_k):
        raise NotImplementedError("dask is not installed; use a ThreadPoolExecutor as parallel_executor")


def dask_wait(futures, *a, **k):
    return _futures_wait(futures, *a, **k)


def install(force: bool = False) -> list[str]:
    """Register the stand-ins for every missing top-level package; returns the names installed."""
    installed = []
    if force or _missing("qiskit"):
        _module("qiskit", __version__="0.0-b200-standin")
        _module(
            "qiskit.circuit",
            QuantumCircuit=_circuit.QuantumCircuit,
            Parameter=_circuit.Parameter,
            ParameterExpression=_circuit.ParameterExpression,
            Gate=_circuit.Gate,
            Instruction=_circuit.Instruction,
            Qubit=_circuit.Qubit,
        )
        _module("qiskit.circuit.library", CU3Gate=_circuit.CU3Gate)
        _module("qiskit.converters", circuit_to_gate=_circuit.circuit_to_gate)
        prim = dict(
            BaseEstimatorV2=BaseEstimatorV2,
            BaseSamplerV2=BaseSamplerV2,
            EstimatorPubLike=Any,
            SamplerPubLike=Any,
            PrimitiveResult=_containers.PrimitiveResult,
            PubResult=_containers.PubResult,
            SamplerPubResult=_containers.SamplerPubResult,
            BasePrimitiveJob=_containers.BasePrimitiveJob,
            PrimitiveJob=_containers.PrimitiveJob,
            DataBin=_containers.DataBin,
            BitArray=_containers.ShotRegister,
        )
        _module("qiskit.primitives", **prim)
        _module("qiskit.primitives.base", BaseEstimatorV2=BaseEstimatorV2, BaseSamplerV2=BaseSamplerV2)
        _module("qiskit.primitives.containers", **prim)
        _module("qiskit.primitives.containers.estimator_pub", EstimatorPub=_containers.EstimatorPub)
        _module("qiskit.primitives.containers.sampler_pub", SamplerPub=_containers.SamplerPub)
        _module(
            "qiskit.primitives.primitive_job",
            PrimitiveJob=_containers.PrimitiveJob,
            BasePrimitiveJob=_containers.BasePrimitiveJob,
        )
        _module("qiskit.quantum_info", SparsePauliOp=_operators.SparsePauliOp)
        _module("qiskit.quantum_info.operators", SparsePauliOp=_operators.SparsePauliOp)
        _module("qiskit.quantum_info.operators.base_operator", BaseOperator=BaseOperator)
        _module(
            "qiskit.result",
            QuasiDistribution=_containers.QuasiDistribution,
            ProbDistribution=_containers.ProbDistribution,
            sampled_expectation_value=sampled_expectation_value,
        )
        _module("qiskit.transpiler", PassManager=PassManager)
        _module("qiskit.transpiler.preset_passmanagers", generate_preset_pass_manager=generate_preset_pass_manager)
        _module("qiskit.qpy", dump=_qpy_unavailable, load=_qpy_unavailable)
        installed.append("qiskit")
    if force or _missing("qiskit_algorithms"):
        from . import optimizers as _opt

        _module("qiskit_algorithms", MinimumEigensolver=MinimumEigensolver, MinimumEigensolverResult=MinimumEigensolverResult)
        _module(
            "qiskit_algorithms.minimum_eigensolvers",
            MinimumEigensolver=MinimumEigensolver,
            MinimumEigensolverResult=MinimumEigensolverResult,
        )
        _module("qiskit_algorithms.minimum_eigensolvers.diagonal_estimator", _evaluate_sparsepauli=_evaluate_sparsepauli)
        _module(
            "qiskit_algorithms.optimizers",
            Optimizer=_opt.Optimizer,
            OptimizerResult=_opt.OptimizerResult,
            SPSA=_opt.SPSA,
            NFT=_opt.NFT,
        )
        _module("qiskit_algorithms.utils", algorithm_globals=algorithm_globals)
        _module("qiskit_algorithms.list_or_dict", ListOrDict=ListOrDict)
        installed.append("qiskit_algorithms")
    if force or _missing("dask"):
        _module("dask")
        _module("dask.distributed", Client=Client, Future=Future, wait=dask_wait)
        _module("dask.utils", SerializableLock=SerializableLock)
        installed.append("dask")
    return installed
