"""B200 Estimator / Sampler primitives implementing the Qiskit V2 primitive *contract* the reference consumes

    estimator.run(pubs, *, precision=None).result()[i].data.evs          (circuit_evaluation.py:210-215)
    sampler.run(pubs, *, shots=None).result()[i].data["meas"].get_counts()  (circuit_evaluation.py:54-55)

so the unmodified reference evaluators and wrappers (``TranspilingEstimatorV2``, ``(Batching)Mutex*``,
``Configured*V2``; /root/reference/queasars/circuit_evaluation/*.py) can be handed these objects.  Pubs may
be tuples, ``EstimatorPub`` / ``SamplerPub`` objects (real Qiskit's or the stand-ins) or generators.

Semantics follow [upstream] ``StatevectorEstimator`` / ``StatevectorSampler`` (qiskit 2.4.2):
  * estimator: ``evs = Re <psi|H|psi>``; with ``precision != 0`` Gaussian noise
    ``default_rng(seed).normal(evs, precision)`` from a fresh generator per pub, ``stds = precision``
  * sampler: ``Generator.choice`` = ``cumsum(|psi|^2)`` CDF + ``searchsorted(uniforms, side='right')`` with
    ``uniforms = default_rng(seed).random(shots)``, a fresh generator per pub when ``seed`` is an int.
Only zero-dimensional pubs (one parameter vector per pub) are used by QUEASARS; 2-D ``parameter_values`` are
accepted and produce array-shaped ``evs`` / one register per row.
"""
from __future__ import annotations

import threading
import weakref
from typing import Iterable, Optional, Sequence

import numpy as np

from . import containers as ct
from .batching import CoalescingQueue
from .engine import Engine, HamiltonianHandle, PlanHandle
from .gate_list import from_circuit_or_none, from_circuit

_engines: dict = {}
_engines_lock = threading.Lock()


def get_engine(device: int = 0, dtype: str = "complex128") -> Engine:
    """Process-wide engine per (device, dtype); created lazily so configuration objects stay picklable."""
    key = (int(device), str(np.dtype(dtype).name))
    with _engines_lock:
        eng = _engines.get(key)
        if eng is None:
            eng = _engines[key] = Engine(device=device, dtype=dtype)
        return eng


class _CircuitCache:
    """circuit object -> compiled plan, keyed by identity (the optimizer loop re-submits the *same* circuit
    object on every objective call: evqe/evolutionary_algorithm/mutation.py:63-75)."""

    def __init__(self, engine: Engine):
        self._engine = engine
        self._lock = threading.Lock()
        self._by_id: dict = {}

    def plan_for(self, circuit) -> Optional[PlanHandle]:
        key = id(circuit)
        with self._lock:
            hit = self._by_id.get(key)
            if hit is not None and hit[0]() is circuit:
                return hit[1]
        gates = from_circuit_or_none(circuit)
        # identity-cached circuits are the ones an optimizer loop re-submits: worth caching their constant prefix state
        plan = None if gates is None else self._engine.compile_with_prefix_reuse(gates)
        try:
            ref = weakref.ref(circuit, lambda _r, k=key: self._by_id.pop(k, None))
        except TypeError:  # not weak-referenceable: do not cache by identity
            return plan
        with self._lock:
            self._by_id[key] = (ref, plan)
        return plan

    def bound_plan(self, circuit, values) -> PlanHandle:
        """Slow path for circuits whose angles are not affine in single parameters: bind on the host."""
        bound = circuit.assign_parameters(list(values)) if len(values) else circuit
        return self._engine.compile(from_circuit(bound))


class _B200Primitive:
    def __init__(self, device: int = 0, dtype: str = "complex128", seed=None, coalesce: bool = True):
        self.device, self.dtype, self.seed, self.coalesce = int(device), str(np.dtype(dtype).name), seed, bool(coalesce)
        self._init_runtime()

    def _init_runtime(self):
        self._engine_obj: Optional[Engine] = None
        self._cache_obj: Optional[_CircuitCache] = None
        self._ham_cache: dict = {}
        self._queue_obj: Optional[CoalescingQueue] = None
        self._lock = threading.Lock()

    # picklable: CUDA handles are per process and re-created lazily (the dask route of the reference pickles
    # evaluators + primitives into worker processes: evqe.py:38-44)
    def __getstate__(self):
        return {"device": self.device, "dtype": self.dtype, "seed": self.seed, "coalesce": self.coalesce}

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._init_runtime()

    @property
    def engine(self) -> Engine:
        if self._engine_obj is None:
            self._engine_obj = get_engine(self.device, self.dtype)
        return self._engine_obj

    @property
    def _cache(self) -> _CircuitCache:
        with self._lock:
            if self._cache_obj is None:
                self._cache_obj = _CircuitCache(self.engine)
            return self._cache_obj

    @property
    def _queue(self) -> CoalescingQueue:
        with self._lock:
            if self._queue_obj is None:
                self._queue_obj = CoalescingQueue(self._execute)
            return self._queue_obj

    def hamiltonian_for(self, operator, build_table=None) -> HamiltonianHandle:
        key = (id(operator), build_table)
        with self._lock:
            hit = self._ham_cache.get(key)
            if hit is not None and hit[0] is operator:
                return hit[1]
        handle = self.engine.hamiltonian(operator, build_table=build_table)
        with self._lock:
            if len(self._ham_cache) > 64:
                self._ham_cache.clear()
            self._ham_cache[key] = (operator, handle)
        return handle

    def _resolve(self, circuit, values) -> tuple[PlanHandle, Sequence[float]]:
        """(plan, parameter values as handed in): the float64 conversion happens chunk by chunk inside the engine's
        pipelined submission, overlapped with the GPU work of the previous chunk."""
        plan = self._cache.plan_for(circuit)
        if plan is None:
            values = np.asarray(values if values is not None else (), dtype=np.float64).reshape(-1)
            return self._cache.bound_plan(circuit, values), np.zeros(0)
        return plan, (values if values is not None else ())

    def _submit(self, key, payload):
        if self.coalesce:
            return self._queue.submit(key, payload)
        return self._execute(key, [payload])[0]

    def _execute(self, key, payloads):  # pragma: no cover - overridden
        raise NotImplementedError


def _param_rows(values) -> tuple[np.ndarray, tuple]:
    """pub parameter_values -> (2-D array rows x n_params, result shape)."""
    if values is None:
        return np.zeros((1, 0)), ()
    if hasattr(values, "as_array"):  # qiskit BindingsArray
        values = values.as_array()
    arr = np.asarray(values, dtype=np.float64)
    if arr.ndim <= 1:
        return arr.reshape(1, -1), ()
    return arr.reshape(-1, arr.shape[-1]), arr.shape[:-1]


class B200EstimatorV2(_B200Primitive):
    """EstimatorV2-contract primitive backed by the CUDA statevector engine."""

    def __init__(self, device: int = 0, dtype: str = "complex128", seed=None, default_precision: float = 0.0, coalesce: bool = True):
        super().__init__(device, dtype, seed, coalesce)
        self.default_precision = default_precision

    def __getstate__(self):
        return {**super().__getstate__(), "default_precision": self.default_precision}

    def expectation_values(self, circuits, parameter_values, operator) -> np.ndarray:
        """Fast path used by ``B200OperatorCircuitEvaluator``: exact <H> of (circuit_i, params_i)."""
        ham = self.hamiltonian_for(operator)
        resolved = [self._resolve(c, v) for c, v in zip(circuits, parameter_values)]
        return np.asarray(self._submit(("exp", ham.ham_id), (ham, resolved)))

    def _execute(self, key, payloads):
        ham = payloads[0][0]
        plans, params, sizes = [], [], []
        for _, resolved in payloads:
            sizes.append(len(resolved))
            for plan, vals in resolved:
                plans.append(plan)
                params.append(vals)
        flat = self.engine.expectation(plans, params, ham)
        out, pos = [], 0
        for n in sizes:
            out.append(flat[pos : pos + n])
            pos += n
        return out

    def _noisy(self, evs: np.ndarray, precision: float) -> np.ndarray:
        if not precision:
            return evs
        rng = np.random.default_rng(self.seed)  # fresh generator per pub, like upstream
        return rng.normal(evs, precision)

    def run(self, pubs: Iterable, *, precision: Optional[float] = None):
        try:
            coerced = [ct.EstimatorPub.coerce(pub, precision) for pub in pubs]
            results = []
            for pub in coerced:
                prec = pub.precision if pub.precision is not None else self.default_precision
                rows, shape = _param_rows(pub.parameter_values)
                evs = self.expectation_values([pub.circuit] * len(rows), list(rows), _single_observable(pub.observables))
                evs = self._noisy(np.asarray(evs, dtype=np.float64), float(prec or 0.0)).reshape(shape)
                stds = np.full(shape, float(prec or 0.0))
                results.append(
                    ct.PubResult(ct.DataBin(evs=evs, stds=stds, shape=shape), metadata={"target_precision": prec, "circuit_metadata": {}})
                )
            return ct.FinishedJob(ct.PrimitiveResult(results, metadata={"version": 2, "backend": "queasars_b200"}))
        except Exception as exc:  # surfaced by job.result(), like a PrimitiveJob would
            return ct.FinishedJob(error=exc)


def _single_observable(observables):
    """QUEASARS always passes one operator per pub; unwrap qiskit's ObservablesArray / dict form."""
    obs = observables
    if hasattr(obs, "tolist") and not hasattr(obs, "to_list"):  # ObservablesArray
        obs = obs.tolist()
    if isinstance(obs, (list, tuple)):
        if len(obs) != 1:
            raise NotImplementedError("B200EstimatorV2 supports one observable per pub")
        obs = obs[0]
    if isinstance(obs, dict):  # {pauli label: coeff}
        from .operators import SparsePauliOp

        return SparsePauliOp.from_list(list(obs.items()))
    return obs


class B200SamplerV2(_B200Primitive):
    """SamplerV2-contract primitive: shots drawn on the GPU from the exact statevector distribution."""

    def __init__(self, device: int = 0, dtype: str = "complex128", seed=None, default_shots: int = 1024, coalesce: bool = True):
        super().__init__(device, dtype, seed, coalesce)
        self.default_shots = default_shots

    def __getstate__(self):
        return {**super().__getstate__(), "default_shots": self.default_shots}

    def sample_indices(self, circuits, parameter_values, shots: int) -> np.ndarray:
        """Fast path used by the B200 sampler evaluators: int64 array [len(circuits), shots] of basis states."""
        resolved = [self._resolve(c, v) for c, v in zip(circuits, parameter_values)]
        return np.asarray(self._submit(("smp", int(shots)), resolved))

    def _execute(self, key, payloads):
        shots = key[1]
        plans, params, sizes = [], [], []
        for resolved in payloads:
            sizes.append(len(resolved))
            for plan, vals in resolved:
                plans.append(plan)
                params.append(vals)
        # a fresh default_rng(seed) per pub when seed is an int (or None); a shared Generator is consumed in order
        if isinstance(self.seed, np.random.Generator):
            uniforms = np.stack([self.seed.random(shots) for _ in plans]) if plans else np.zeros((0, shots))
        else:
            uniforms = np.stack([np.random.default_rng(self.seed).random(shots) for _ in plans]) if plans else np.zeros((0, shots))
        flat = self.engine.sample(plans, params, shots, uniforms)
        out, pos = [], 0
        for n in sizes:
            out.append(flat[pos : pos + n])
            pos += n
        return out

    def run(self, pubs: Iterable, *, shots: Optional[int] = None):
        try:
            coerced = [ct.SamplerPub.coerce(pub, shots) for pub in pubs]
            results = []
            for pub in coerced:
                n_shots = int(pub.shots if pub.shots is not None else self.default_shots)
                rows, shape = _param_rows(pub.parameter_values)
                if shape != ():
                    raise NotImplementedError("B200SamplerV2 supports one parameter vector per pub")
                idx = self.sample_indices([pub.circuit], [rows[0]], n_shots)[0]
                reg = ct.ShotRegister(idx, int(pub.circuit.num_qubits))
                results.append(ct.SamplerPubResult(ct.DataBin(meas=reg, shape=()), metadata={"shots": n_shots, "circuit_metadata": {}}))
            return ct.FinishedJob(ct.PrimitiveResult(results, metadata={"version": 2, "backend": "queasars_b200"}))
        except Exception as exc:
            return ct.FinishedJob(error=exc)
