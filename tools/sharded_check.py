"""Sharded-statevector check + timing, one process per GPU:
   torchrun --nproc-per-node N tools/sharded_check.py --qubits 26 --layers 2
Rank 0 compares <H> of the sharded run with the single-GPU engine (when the state fits one GPU) and prints
swap timings (NVLink GB/s per direction per GPU)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from queasars_b200 import gate_list as gl  # noqa: E402
from queasars_b200 import genome as gn  # noqa: E402
from queasars_b200.sharded import ShardedStatevector  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--qubits", dest="n", type=int, default=26)
ap.add_argument("--layers", type=int, default=2)
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--analytic", type=int, default=0, help="product-state circuit with analytically known <Z_q>, <Z_p Z_q>")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

op = gn.ising_operator(args.n, seed=3)
_, z, c = op.masks()
z, c = z[: 2 * args.n], c.real[: 2 * args.n]  # n Z terms + n ZZ terms keep the on-the-fly diagonal kernel cheap
analytic_value = None
if args.analytic:
    from queasars_b200.circuit import QuantumCircuit

    thetas = np.random.default_rng(5).uniform(0, np.pi, args.n)
    circ = QuantumCircuit(args.n)
    for rep in range(2):  # two half rotations per qubit: the second one is a real gate on an existing state
        for q in range(args.n):
            circ.ry(float(thetas[q]) / 2, q)
    gates = gl.from_circuit(circ)
    cosines = np.cos(thetas)
    analytic_value = float(sum(cf * np.prod([cosines[q] for q in range(args.n) if (int(zm) >> q) & 1]) for zm, cf in zip(z, c)))

    class _Ind:
        parameter_values = ()

    ind = _Ind()
else:
    ind = gn.Individual.random(args.n, args.layers, True, 11)
    gates = gl.from_evqe_individual(ind)

sv = ShardedStatevector(args.n)
swaps_in_run = None
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
sv.run(gates, ind.parameter_values)
torch.cuda.synchronize()
t_run = time.perf_counter() - t0
swaps_in_run = sv.swaps_done
t0 = time.perf_counter()
value = sv.diagonal_expectation(z, c)
t_exp = time.perf_counter() - t0
norm = sv.norm_squared()
# general Pauli sum (transverse-field Ising: X on every qubit, some of them on rank bits) and sampling
from queasars_b200.operators import SparsePauliOp  # noqa: E402

tfim_terms = []
for q in range(args.n - 1):
    lab = ["I"] * args.n
    lab[args.n - 1 - q] = lab[args.n - 2 - q] = "Z"
    tfim_terms.append(("".join(lab), -1.0))
for q in range(args.n):
    lab = ["I"] * args.n
    lab[args.n - 1 - q] = "X"
    tfim_terms.append(("".join(lab), -0.5))
tfim = SparsePauliOp.from_list(tfim_terms)
# closed form on the product state prod_q RY(theta_q)|0>: <Z_q Z_q+1> = cos(theta_q) cos(theta_q+1), <X_q> = sin(theta_q)
tfim_analytic = None
if args.analytic:
    tfim_analytic = float(-np.sum(np.cos(thetas[:-1]) * np.cos(thetas[1:])) - 0.5 * np.sum(np.sin(thetas)))
swaps_before = sv.swaps_done
torch.cuda.synchronize()
t0 = time.perf_counter()
tfim_value = sv.expectation(tfim)
t_tfim = time.perf_counter() - t0
tfim_swaps = sv.swaps_done - swaps_before
t0 = time.perf_counter()
shots = sv.sample(10000, seed=123)
t_sample = time.perf_counter() - t0
sampled_energy = float(np.mean([sum(cf * (1 - 2 * (bin(int(s) & int(zm)).count("1") & 1)) for zm, cf in zip(z, c)) for s in shots[:2000]]))

# swap-only timing
swap_ms = None
if world > 1:
    lp = list(range(sv.n_local - sv.n_global, sv.n_local))
    for _ in range(2):
        sv._swap_all_global(lp)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        sv._swap_all_global(lp)
    e1.record(); torch.cuda.synchronize()
    swap_ms = e0.elapsed_time(e1) / 4

if rank == 0:
    out = {"n": args.n, "world": world, "n_local": sv.n_local, "gates": len(gates.ops), "swaps": swaps_in_run, "value": value, "analytic_value": analytic_value,
           "analytic_rel_err": None if analytic_value is None else abs(value - analytic_value) / max(1.0, abs(analytic_value)),
           "norm_err": abs(norm - 1.0), "run_s": t_run, "expectation_s": t_exp, "tfim_value": tfim_value, "tfim_s": t_tfim, "tfim_swaps": tfim_swaps,
           "tfim_analytic_value": tfim_analytic, "tfim_analytic_rel_err": None if tfim_analytic is None else abs(tfim_value - tfim_analytic) / max(1.0, abs(tfim_analytic)),
           "sample_10k_s": t_sample, "sampled_energy_2000": sampled_energy, "swap_path": "p2p kernel (peer memory)" if sv._peer_ptrs is not None else "nccl all_to_all"}
    if swap_ms is not None:
        shard_bytes = 16 * (1 << sv.n_local)
        sent = shard_bytes * (world - 1) / world
        out.update(swap_ms=swap_ms, swap_sent_GB=sent / 1e9, swap_GBps_per_dir=sent / (swap_ms * 1e-3) / 1e9,
                   swap_includes="one fused kernel (peer stores) + barrier" if sv._peer_ptrs is not None else "pack + all_to_all_single + unpack")
    if args.check and args.n <= 30:
        from queasars_b200.engine import Engine
        from queasars_b200.operators import SparsePauliOp

        eng = Engine(local)
        ham = eng.hamiltonian(SparsePauliOp._raw(args.n, [0] * len(z), [int(v) for v in z], [float(v) for v in c]), build_table=False)
        plan = eng.compile(gates)
        ref = eng.expectation([plan], [list(ind.parameter_values)], ham)[0]
        out["single_gpu_value"] = float(ref)
        out["rel_err"] = abs(value - ref) / max(1.0, abs(ref))
        tref = eng.expectation([plan], [list(ind.parameter_values)], eng.hamiltonian(tfim))[0]
        out["tfim_rel_err"] = abs(tfim_value - tref) / max(1.0, abs(tref))
        out["sampled_vs_exact_energy"] = [sampled_energy, float(ref)]
    print(json.dumps(out))
sv.close()
if world > 1:
    dist.destroy_process_group()
