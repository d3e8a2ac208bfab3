"""One statevector sharded over several GPUs by its top ("global") qubits -- BASELINE config C5
(35 qubits complex128 = 512 GiB over 8 x B200).  New capability: the reference has no single-state scaling
mechanism; the same `evaluate` semantics as the single-GPU engine are kept.

Layout: world = 2^g ranks, rank r holds the 2^(n-g) amplitudes whose top g index bits equal r.  A *physical*
bit position p < n-g is local, p >= n-g is a rank bit.  Logical qubits are tracked through a permutation
(`position[q]`), so a swap is pure relabelling plus data movement:

  * gates whose target sits on a local position run the normal sweep kernel on the shard
    (``qb_apply_plan_device`` with ``index_offset = rank << n_local``: a control or diagonal target on a rank bit is
    a per-shard predicate, no communication);
  * before a gate that *targets* a rank bit, all g rank bits are exchanged with g local positions whose qubits are
    needed latest.  On GPUs the exchange is ONE kernel per rank (``qb_swap_global_p2p``): it streams the shard once and
    stores every amplitude straight into the peer that owns it afterwards, at its final index, through peer-mapped
    buffers (torch symmetric memory over NVLink 5 / NVSwitch) -- the all-to-all is fused into the permutation, there is
    no pack pass, no staging copy and no unpack pass.  Fallback (``QB_SWAP=nccl``, CPU/gloo tests, no peer mapping):
    pack -> ``all_to_all_single`` -> unpack.  Two buffers of shard size are used in ping-pong either way.

The local arithmetic is delegated to a *backend* object (``apply(state, gate_ops, params, n_local, index_offset)``,
``expectation(state, masks..., index_offset)``); the product default drives the CUDA engine.  Tests inject a NumPy
backend to exercise the multi-rank host logic on CPU with the gloo backend.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Sequence

import numpy as np

from .gate_list import DENSE, GateList, KernelOp


def _remap(op: KernelOp, position: Sequence[int]) -> KernelOp:
    return KernelOp(op.kind, position[op.target], -1 if op.control < 0 else position[op.control], op.gamma, op.theta, op.phi, op.lam)


class CudaShardBackend:
    """Runs the shard-local work on the CUDA engine (torch tensors only provide the device memory)."""

    def __init__(self, engine):
        self.engine = engine

    @staticmethod
    def _join_torch_stream(state):
        # the engine launches on its own stream: everything torch queued on this buffer (pack / unpack copies,
        # the NCCL all-to-all) has to be complete before the kernels touch it.  The native calls synchronise
        # their stream before returning, which orders the other direction.
        import torch

        torch.cuda.current_stream(state.device).synchronize()

    def apply(self, state, ops: list[KernelOp], params: np.ndarray, n_local: int, n_params: int, index_offset: int, init_zero: bool):
        # controls may sit on rank bits (>= n_local), so the single-register product-state prefix does not apply
        plan = self.engine.compile(GateList(n_local, ops, n_params, ()), from_zero_state=False)
        self._join_torch_stream(state)
        self.engine.apply_plan_device(plan, params, state.data_ptr(), init_zero, index_offset)

    def diagonal_expectation(self, state, z_masks: np.ndarray, coeffs: np.ndarray, n_total: int, n_local: int, index_offset: int) -> float:
        from .operators import SparsePauliOp

        op = SparsePauliOp._raw(n_total, [0] * len(z_masks), [int(z) for z in z_masks], [float(c) for c in coeffs])
        ham = self.engine.hamiltonian(op, build_table=False)
        self._join_torch_stream(state)
        return self.engine.expectation_device(ham, "complex128", n_local, state.data_ptr(), index_offset)

    def pauli_expectation(self, state, x_masks, z_masks, coeffs, n_total: int, n_local: int, index_offset: int) -> float:
        """Shard-local part of Re <psi| sum_t c_t P_t |psi> for terms whose X part only flips local qubits (z may reach
        rank bits: they enter through index_offset)."""
        from .operators import SparsePauliOp

        op = SparsePauliOp._raw(n_total, [int(x) for x in x_masks], [int(z) for z in z_masks], [complex(c) for c in coeffs])
        ham = self.engine.hamiltonian(op, build_table=False)
        self._join_torch_stream(state)
        return self.engine.expectation_device(ham, "complex128", n_local, state.data_ptr(), index_offset)

    def sample(self, state, uniforms: np.ndarray, n_local: int) -> np.ndarray:
        self._join_torch_stream(state)
        return self.engine.sample_device("complex128", n_local, state.data_ptr(), uniforms)

    def swap_p2p(self, state, peer_ptrs: Sequence[int], n_local: int, rank: int, local_positions: Sequence[int]) -> None:
        """Fused swap + all-to-all into the peers' buffers; returns when this rank's stores have been issued and completed."""
        self._join_torch_stream(state)
        self.engine.swap_global_p2p("complex128", n_local, state.data_ptr(), peer_ptrs, rank, local_positions)
        self.engine.synchronize()


class ShardedStatevector:
    def __init__(self, n_qubits: int, backend=None, group=None, device=None, min_local: int = 12):
        import torch
        import torch.distributed as dist

        self._torch, self._dist = torch, dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        g = int(round(math.log2(self.world)))
        if 1 << g != self.world:
            raise ValueError("the number of ranks must be a power of two")
        self.n_qubits, self.n_global, self.n_local = n_qubits, g, n_qubits - g
        if self.n_local < max(min_local, 2 * g):
            raise ValueError(f"{n_qubits} qubits over {self.world} ranks leaves only {self.n_local} local qubits")
        if backend is None:
            from .primitives import get_engine

            dev_index = torch.cuda.current_device() if device is None else torch.device(device).index
            backend = CudaShardBackend(get_engine(dev_index, "complex128"))
            device = torch.device("cuda", dev_index)
        self.backend = backend
        self.device = torch.device("cpu") if device is None else torch.device(device)
        size = 1 << self.n_local
        self._peer_ptrs = None  # data_ptr of a local buffer -> that buffer's address on every rank, as mapped into this process
        if self.world > 1 and self.device.type == "cuda" and hasattr(backend, "swap_p2p") and os.environ.get("QB_SWAP", "p2p") != "nccl":
            try:
                import torch.distributed._symmetric_memory as symm

                group_name = (group if group is not None else dist.group.WORLD).group_name
                bufs = [symm.empty(size, dtype=torch.complex128, device=self.device) for _ in range(2)]
                handles = [symm.rendezvous(b, group_name) for b in bufs]
                self._peer_ptrs = {b.data_ptr(): [int(p) for p in h.buffer_ptrs] for b, h in zip(bufs, handles)}
                self._symm_keepalive = (bufs, handles)
                self.state, self.spare = bufs
                self.state.zero_()
            except Exception as exc:  # no peer mapping on this system: NCCL all-to-all path
                self._peer_ptrs = None
                self.swap_fallback_reason = repr(exc)
        if self._peer_ptrs is None:
            self.state = torch.zeros(size, dtype=torch.complex128, device=self.device)
            self.spare = torch.empty(size, dtype=torch.complex128, device=self.device)
        self.position = list(range(n_qubits))  # logical qubit -> physical bit position
        self.swaps_done = 0
        self.bytes_sent = 0
        self._fresh = True

    @property
    def index_offset(self) -> int:
        return self.rank << self.n_local

    # ------------------------------------------------------------------ global <-> local swap
    def _swap_all_global(self, local_positions: Sequence[int]) -> None:
        """Exchange the g rank bits with the local bit positions ``local_positions`` (ascending list of length g).
        Afterwards rank bit j holds what was at local_positions[j] and vice versa."""
        g, nl = self.n_global, self.n_local
        if g == 0:
            return
        torch, dist = self._torch, self._dist
        lp = list(local_positions)
        assert len(lp) == g and len(set(lp)) == g and all(0 <= p < nl for p in lp)
        if self._peer_ptrs is not None:
            # one kernel: every amplitude goes straight to the rank that owns it afterwards, at its final index
            self.backend.swap_p2p(self.state, self._peer_ptrs[self.spare.data_ptr()], nl, self.rank, lp)
            dist.barrier(group=self.group)  # every rank's stores have landed before anyone reads its new shard
            self.state, self.spare = self.spare, self.state
        else:
            # tensor axis a <-> bit nl-1-a ; bring the chosen bits to the front, most significant first = rank bit g-1
            front_bits = [lp[j] for j in range(g - 1, -1, -1)]
            rest_bits = [b for b in range(nl - 1, -1, -1) if b not in lp]
            perm = [nl - 1 - b for b in front_bits + rest_bits]
            packed = self.spare.view((2,) * nl)
            packed.copy_(self.state.view((2,) * nl).permute(perm))  # pack
            send = self.spare.view(self.world, -1)
            recv = self.state.view(self.world, -1)
            dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)  # chunk i <-> rank i
            # received chunk j came from rank j and carries its elements with (chosen local bits) == my rank;
            # its position j now plays the role of the chosen local bits -> undo the packing permutation
            inverse = [0] * nl
            for axis, src in enumerate(perm):
                inverse[src] = axis
            self.spare.view((2,) * nl).copy_(self.state.view((2,) * nl).permute(inverse))  # unpack
            self.state, self.spare = self.spare, self.state
        for q in range(self.n_qubits):
            p = self.position[q]
            if p >= nl:
                self.position[q] = lp[p - nl]
            elif p in lp:
                self.position[q] = nl + lp.index(p)
        self.swaps_done += 1
        self.bytes_sent += (self.world - 1) * (self.state.numel() // self.world) * 16

    # ------------------------------------------------------------------ circuit execution
    def run(self, gates: GateList, params: Sequence[float]) -> None:
        """Apply the circuit to |0...0> (first call) or to the current state."""
        if gates.n_qubits != self.n_qubits:
            raise ValueError("circuit and sharded state act on different numbers of qubits")
        params = np.ascontiguousarray(np.asarray(params, dtype=np.float64).reshape(-1))
        remaining = list(gates.ops)
        nl = self.n_local
        while remaining or self._fresh:
            # Everything that can run with the current placement: ops whose dense target is local and that do not
            # depend on a deferred op (same reordering rule as the sweep planner: two ops commute when they act
            # diagonally on every qubit they share).  Ops targeting a rank bit are deferred and served by ONE swap.
            pend_dense: set[int] = set()
            pend_any: set[int] = set()
            segment, deferred = [], []
            for op in remaining:
                t, c = op.target, op.control
                dense = op.kind == DENSE
                blocked = (t in pend_any) if dense else (t in pend_dense)
                if c >= 0 and c in pend_dense:
                    blocked = True
                if not blocked and not (dense and self.position[t] >= nl):
                    segment.append(_remap(op, self.position))
                    continue
                deferred.append(op)
                if dense:
                    pend_dense.add(t)
                pend_any.add(t)
                if c >= 0:
                    pend_any.add(c)
            if segment or self._fresh:
                self.backend.apply(self.state, segment, params, nl, gates.n_params, self.index_offset, self._fresh)
                self._fresh = False
            remaining = deferred
            if remaining:
                self._swap_all_global(self._choose_local_positions(remaining))

    def _choose_local_positions(self, upcoming: Sequence[KernelOp]) -> list[int]:
        """Local positions to give up: those whose qubits are targeted latest (or never) by the upcoming ops."""
        nl = self.n_local
        first_use = {}
        for t, op in enumerate(upcoming):
            if op.kind == DENSE and op.target not in first_use:
                first_use[op.target] = t
        if all(self.position[q] < nl for q in first_use):  # nothing upcoming targets a rank bit
            return list(range(nl - self.n_global, nl))
        logical_at = {self.position[q]: q for q in range(self.n_qubits)}
        candidates = sorted(range(nl), key=lambda p: (-first_use.get(logical_at[p], 1 << 30), -p))
        return sorted(candidates[: self.n_global])

    # ------------------------------------------------------------------ observables
    def diagonal_expectation(self, z_masks: Sequence[int], coeffs: Sequence[float]) -> float:
        """<psi| sum_t c_t Z^{z_t} |psi> ; every rank returns the global value (one all-reduce of a double)."""
        phys = []
        for z in z_masks:
            m = 0
            for q in range(self.n_qubits):
                if (int(z) >> q) & 1:
                    m |= 1 << self.position[q]
            phys.append(m)
        part = self.backend.diagonal_expectation(
            self.state, np.asarray(phys, dtype=np.uint64), np.asarray(coeffs, dtype=np.float64), self.n_qubits, self.n_local, self.index_offset
        )
        if self.world == 1:
            return float(part)
        t = self._torch.tensor([part], dtype=self._torch.float64, device=self.device)
        self._dist.all_reduce(t, group=self.group)
        return float(t.item())

    def norm_squared(self) -> float:
        return self.diagonal_expectation([0], [1.0])

    def _physical_mask(self, mask: int) -> int:
        m = 0
        for q in range(self.n_qubits):
            if (int(mask) >> q) & 1:
                m |= 1 << self.position[q]
        return m

    def expectation(self, operator) -> float:
        """Re <psi|H|psi> for a general Pauli sum (SparsePauliOp-like: ``masks()`` -> x, z, coeffs as in operators.py).
        Terms whose X part flips only local qubits are evaluated shard-locally (Z factors on rank bits are per-shard
        signs); for the others ONE global swap brings the flipped rank bits down first, giving up local positions that
        no remaining term flips.  Every rank returns the global value."""
        from .engine import operator_terms

        n, xs, zs, cs = operator_terms(operator)
        if n != self.n_qubits:
            raise ValueError("operator and sharded state act on different numbers of qubits")
        todo = [(int(x), int(z), complex(c)) for x, z, c in zip(xs, zs, cs)]
        total = 0.0
        rank_bits = ((1 << self.n_global) - 1) << self.n_local
        for attempt in range(3):
            now = [t for t in todo if not (self._physical_mask(t[0]) & rank_bits)]
            todo = [t for t in todo if self._physical_mask(t[0]) & rank_bits]
            if now:
                total += self.backend.pauli_expectation(
                    self.state, [self._physical_mask(t[0]) for t in now], [self._physical_mask(t[1]) for t in now], [t[2] for t in now],
                    self.n_qubits, self.n_local, self.index_offset,
                )
            if not todo:
                break
            flipped = 0
            for x, _, _ in todo:
                flipped |= self._physical_mask(x)
            free = [p for p in range(self.n_local - 1, -1, -1) if not (flipped >> p) & 1]
            if attempt == 2 or len(free) < self.n_global:
                raise NotImplementedError("a Pauli term flips more qubits than fit one shard next to the rank bits")
            self._swap_all_global(sorted(free[: self.n_global]))
        if self.world == 1:
            return float(total)
        t = self._torch.tensor([total], dtype=self._torch.float64, device=self.device)
        self._dist.all_reduce(t, group=self.group)
        return float(t.item())

    def sample(self, shots: int, seed=None, uniforms: Optional[np.ndarray] = None) -> np.ndarray:
        """``shots`` basis-state indices (logical qubit order) drawn from |psi|^2; every rank returns all of them.

        [upstream] Statevector.sample_memory semantics per shot (inverse-CDF of one uniform), evaluated in two levels: the
        all-gathered shard masses pick the shard of every shot, the owning rank then inverts its own CDF with the re-scaled
        uniform (qb_sample_device); the physical index (rank bits | local index) is mapped back through the qubit
        permutation.  The enumeration order of the CDF is the physical one, so for given uniforms the draws are a different
        -- equally distributed -- realisation than the single-GPU sampler's."""
        torch, dist = self._torch, self._dist
        if uniforms is None:
            uniforms = np.random.default_rng(seed).random(shots)
        u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1)
        shots = u.size
        mass = self.backend.diagonal_expectation(self.state, np.zeros(1, dtype=np.uint64), np.ones(1), self.n_qubits, self.n_local, self.index_offset)
        if self.world > 1:
            masses = torch.zeros(self.world, dtype=torch.float64, device=self.device)
            masses[self.rank] = mass
            dist.all_reduce(masses, group=self.group)
            masses = masses.cpu().numpy()
        else:
            masses = np.array([mass])
        edges = np.concatenate([[0.0], np.cumsum(masses)])
        scaled = u * edges[-1]
        owner = np.minimum(np.searchsorted(edges[1:], scaled, side="right"), self.world - 1)
        mine = np.nonzero(owner == self.rank)[0]
        physical = np.zeros(shots, dtype=np.int64)
        if mine.size:
            local_u = np.clip((scaled[mine] - edges[self.rank]) / masses[self.rank], 0.0, np.nextafter(1.0, 0.0))
            physical[mine] = self.backend.sample(self.state, local_u, self.n_local) | (self.rank << self.n_local)
        if self.world > 1:
            t = torch.from_numpy(physical).to(self.device)
            dist.all_reduce(t, group=self.group)  # the ranks fill disjoint shots
            physical = t.cpu().numpy()
        logical = np.zeros_like(physical)
        for q in range(self.n_qubits):
            logical |= ((physical >> self.position[q]) & 1) << q
        return logical

    def gather_physical(self) -> np.ndarray:
        """Full statevector in *physical* bit order (rank bits on top) on every rank (testing at small sizes only)."""
        torch, dist = self._torch, self._dist
        if self.world == 1:
            return self.state.cpu().numpy()
        parts = [torch.empty_like(self.state) for _ in range(self.world)]
        dist.all_gather(parts, self.state, group=self.group)
        return torch.cat(parts).cpu().numpy()

    def gather_logical(self) -> Optional[np.ndarray]:
        """Full statevector in *logical* qubit order on every rank (testing at small sizes only)."""
        torch, dist = self._torch, self._dist
        if self.world > 1:
            parts = [torch.empty_like(self.state) for _ in range(self.world)]
            dist.all_gather(parts, self.state, group=self.group)
            full = torch.cat(parts).cpu().numpy()
        else:
            full = self.state.cpu().numpy()
        n = self.n_qubits
        # physical axis a <-> physical bit n-1-a ; logical bit q lives at physical position[q]
        perm = [n - 1 - self.position[q] for q in range(n - 1, -1, -1)]
        return np.ascontiguousarray(full.reshape((2,) * n).transpose(perm)).reshape(-1)
