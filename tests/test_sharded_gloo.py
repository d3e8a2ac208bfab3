"""Multi-rank host logic of queasars_b200.sharded (global-qubit swaps, position tracking, segmenting, all-reduce)
on CPU: world_size 2 and 4 over the gloo backend with a NumPy shard backend standing in for the CUDA engine."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import evqe_genome as og
from oracle import qiskit_semantics as oq
from queasars_b200 import gate_list as gl
from queasars_b200.gate_list import DENSE
from tests.test_frontend_planner import build_circuit


class NumpyShardBackend:
    """Checker backend: shard-local gate application with the oracle's tensordot (test infrastructure)."""

    def apply(self, state, ops, params, n_local, n_params, index_offset, init_zero):
        arr = state.numpy()
        if init_zero:
            arr[:] = 0
            if index_offset == 0:
                arr[0] = 1
        idx = np.arange(arr.size, dtype=np.int64)
        for op in ops:
            m = np.array(op.matrix(list(params)))
            ctrl_local = op.control >= 0 and op.control < n_local
            if op.control >= n_local and not (index_offset >> op.control) & 1:
                continue
            if op.kind == DENSE:
                assert op.target < n_local
                new = oq.apply_matrix(arr.copy(), n_local, m, (op.target,))
                if ctrl_local:
                    mask = ((idx >> op.control) & 1).astype(bool)
                    arr[mask] = new[mask]
                else:
                    arr[:] = new
            else:
                bit = ((idx >> op.target) & 1) if op.target < n_local else np.full_like(idx, (index_offset >> op.target) & 1)
                factor = np.where(bit == 1, m[1, 1], m[0, 0])
                if ctrl_local:
                    factor = np.where(((idx >> op.control) & 1) == 1, factor, 1.0)
                arr *= factor

    def diagonal_expectation(self, state, z_masks, coeffs, n_total, n_local, index_offset):
        arr = state.numpy()
        idx = np.arange(arr.size, dtype=np.uint64) | np.uint64(index_offset)
        probs = arr.real**2 + arr.imag**2
        total = 0.0
        for z, c in zip(z_masks, coeffs):
            v = idx & np.uint64(z)
            for s in (32, 16, 8, 4, 2, 1):
                v ^= v >> np.uint64(s)
            total += c * float(np.dot(probs, 1.0 - 2.0 * (v & np.uint64(1)).astype(np.float64)))
        return total


    def pauli_expectation(self, state, x_masks, z_masks, coeffs, n_total, n_local, index_offset):
        """sum_t Re[c_t <psi|P_t|psi>] restricted to this shard's rows, P|k> = i^{#Y} (-1)^{pc(k&z)} |k^x> (oracle/qiskit_semantics.py)."""
        arr = state.numpy()
        idx = np.arange(arr.size, dtype=np.uint64)
        total = 0.0
        for x, z, c in zip(x_masks, z_masks, coeffs):
            assert int(x) >> n_local == 0
            v = (idx | np.uint64(index_offset)) & np.uint64(z)
            for s in (32, 16, 8, 4, 2, 1):
                v ^= v >> np.uint64(s)
            sign = 1.0 - 2.0 * (v & np.uint64(1)).astype(np.float64)
            phase = 1j ** bin(int(x) & int(z)).count("1")
            total += float((c * phase * np.sum(np.conj(arr[idx ^ np.uint64(x)]) * sign * arr)).real)
        return total

    def sample(self, state, uniforms, n_local):
        arr = state.numpy()
        return oq.sample_indices(arr / np.linalg.norm(arr), len(uniforms), uniforms=np.asarray(uniforms))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, layers, seed, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from queasars_b200.sharded import ShardedStatevector

        genome, values = og.random_individual(n, layers, True, seed)
        instr = og.individual_circuit(genome, values)
        gates = gl.from_circuit(build_circuit(instr, n))
        sv = ShardedStatevector(n, backend=NumpyShardBackend(), min_local=4)
        sv.run(gates, values)
        got = sv.gather_logical()
        want = oq.statevector(instr, n, values)
        rng = np.random.default_rng(5)
        z_masks = [int(rng.integers(0, 1 << n)) for _ in range(6)]
        coeffs = [float(c) for c in rng.normal(size=6)]
        e_got = sv.diagonal_expectation(z_masks, coeffs)
        e_want = float(np.dot(np.abs(want) ** 2, oq.diagonal_table(n, list(zip(z_masks, coeffs)))))
        # general Pauli sum: X / Y factors on every qubit in turn (some sit on rank bits -> one more swap), ZZ couplings
        labels = []
        for q in range(n):
            lab = ["I"] * n
            lab[n - 1 - q] = "X" if q % 2 else "Y"
            lab[n - 1 - (q + 1) % n] = "Z"
            labels.append(("".join(lab), float(rng.normal())))
        labels.append(("Z" * n, 0.25))
        from queasars_b200.operators import SparsePauliOp

        p_got = sv.expectation(SparsePauliOp.from_list(labels))
        p_want = oq.estimator_expectation(want, labels)
        # sampling: the draws must be the inverse-CDF draws in *physical* enumeration order, mapped back to logical indices
        shots = 2000
        uniforms = np.random.default_rng(17).random(shots)
        drawn = sv.sample(shots, uniforms=uniforms)
        full_phys = sv.gather_physical()
        phys = oq.sample_indices(full_phys, shots, uniforms=uniforms)
        expect = np.zeros_like(phys)
        for q in range(n):
            expect |= ((phys >> sv.position[q]) & 1) << q
        mismatches = int(np.count_nonzero(drawn != expect))
        zero_prob = int(np.count_nonzero(np.abs(want[drawn]) ** 2 == 0.0))
        if rank == 0:
            out.put((float(np.max(np.abs(got - want))), abs(e_got - e_want), sv.swaps_done, abs(sv.norm_squared() - 1.0), abs(p_got - p_want), mismatches, zero_prob))
        else:
            sv.norm_squared()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,layers,seed", [(2, 8, 4, 0), (4, 9, 5, 1), (2, 10, 3, 2)])
def test_sharded_state_matches_oracle(world, n, layers, seed):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, layers, seed, out)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    err, e_err, swaps, norm_err, p_err, mismatches, zero_prob = out.get(timeout=10)
    assert err < 1e-12 and e_err < 1e-12 and norm_err < 1e-12 and p_err < 1e-12
    assert swaps >= 1  # the circuits do target global qubits
    assert mismatches <= 2 and zero_prob == 0  # (a re-scaled uniform may land on the other side of a CDF boundary)


def test_single_rank_no_swaps():
    from queasars_b200.sharded import ShardedStatevector

    n = 6
    genome, values = og.random_individual(n, 3, True, 3)
    instr = og.individual_circuit(genome, values)
    sv = ShardedStatevector(n, backend=NumpyShardBackend(), min_local=4)
    sv.run(gl.from_circuit(build_circuit(instr, n)), values)
    assert sv.swaps_done == 0
    np.testing.assert_allclose(sv.gather_logical(), oq.statevector(instr, n, values), atol=1e-13)


@pytest.mark.parametrize("n_local,lp", [(10, [9]), (11, [2, 10]), (12, [0, 5, 11]), (12, [3, 4, 5])])
def test_fused_swap_work_order_is_the_bit_permutation(n_local, lp):
    """NumPy mirror of swap_p2p_kernel's index arithmetic (queasars_b200/csrc/qb_kernels.cuh): element number t of rank r ->
    destination rank d = run number XOR r, source index i with d's bits dropped in at the exchanged positions, destination index
    i with those bits replaced by r's.  Over all ranks this must be exactly the permutation rank bit j <-> local bit lp[j], every
    element moved once, and at any element number the ranks must target pairwise different peers."""
    g, world, size = len(lp), 1 << len(lp), 1 << n_local
    run_bits = min(10, n_local - g)
    t = np.arange(size, dtype=np.int64)
    lpmask = sum(1 << p for p in lp)
    full_src = np.arange(world * size, dtype=np.int64)
    landed = np.full(world * size, -1, dtype=np.int64)
    dest_of_rank = []
    for r in range(world):
        d = ((t >> run_bits) & (world - 1)) ^ r
        i = ((t >> (run_bits + g)) << run_bits) | (t & ((1 << run_bits) - 1))
        for j, p in enumerate(lp):  # ascending: open a gap at p, drop in bit j of d
            i = ((i >> p) << (p + 1)) | (i & ((1 << p) - 1)) | (((d >> j) & 1) << p)
        assert np.array_equal(np.sort(i), t)  # every source element of the shard exactly once
        rbits = sum(((r >> j) & 1) << p for j, p in enumerate(lp))
        dst = (d << n_local) | (i & ~lpmask) | rbits
        assert np.all(landed[dst] == -1)
        landed[dst] = full_src[(r << n_local) | i]
        dest_of_rank.append(d)
    swapped = full_src.copy()
    for j, p in enumerate(lp):
        a, b = (full_src >> (n_local + j)) & 1, (full_src >> p) & 1
        swapped &= ~((1 << (n_local + j)) | (1 << p))
        swapped |= (b << (n_local + j)) | (a << p)
    want = np.empty_like(full_src)
    want[swapped] = full_src
    assert np.array_equal(landed, want)
    stacked = np.stack(dest_of_rank)  # ranks x element numbers
    assert all(len(set(stacked[:, k])) == world for k in range(0, size, max(1, size // 64)))
