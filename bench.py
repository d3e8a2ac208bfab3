#!/usr/bin/env python
"""Benchmark of the EVQE circuit-evaluation hot path (BASELINE.json: "EVQE circuit evals/sec at 20q; gate-apply
HBM GB/s vs peak; at 1/2/4/8 GPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--layers L]

Workload (BASELINE.json configs[1], SURVEY.md section 8d "C2"): a population of 32 random EVQE individuals
(reference generation rules, seed 0 + 1000*rank) on 20 qubits with L parameterised layers, evaluated against a
synthetic random diagonal Ising Hamiltonian (20 Z + 190 ZZ terms, default_rng(1234)) -- one "step" = one batched
``evaluate_circuits`` submission of the whole population, i.e. one EVQE selection pass.
  value   circuit evaluations / s with parameters + plans resident on the device (CUDA events, kernels only)
  e2e     the same through ``B200OperatorCircuitEvaluator.evaluate_circuits`` with host lists in / floats out
  roofline  algorithmic bytes (2 * 16 B * 2^n per swept statevector) / CUDA-event time of the sweep kernel
  gate_apply  the same sweep kernel on ONE 26-qubit state (1 GiB, HBM-resident): the 24-30 q gate-apply figure
Multi-GPU (torchrun): every rank evaluates its own population of 32 (weak scaling, no data-path collective).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_QUBITS = 20
POPULATION = 32
METRIC = "evqe_circuit_evals_per_sec_20q"
UNIT = "circuit_evals/s"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def profiled_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per sweep launch of this workload, from the committed ncu --set full
    capture (profiles/r1_sweep20_traffic.json); None when no capture is present."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_sweep20_traffic.json")) as fh:
            return float(json.load(fh)["mean_traffic_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device_index: int, enabled: bool = True):
        # one sampler per job (rank 0): eight nvidia-smi processes starting at once take longer to come up than a timed region lasts
        self.device_index = device_index
        self.enabled = enabled
        self.proc = None
        self.path = None

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.device_index)],
                stdout=open(self.path, "w"),
                stderr=subprocess.DEVNULL,
            )
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        if not self.enabled:
            return None
        if self.proc is None or not self.path:
            return self.single_sample()
        sm, smax, reasons = [], [], set()
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1])), smax.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            return None
        if not sm:
            return self.single_sample()
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(smax)), "reasons": sorted(reasons), "samples": len(sm)}

    def single_sample(self):
        """Fallback when the looping nvidia-smi produced no line inside the (short) timed region: one query right after it."""
        try:
            out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.device_index)],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            parts = [p.strip() for p in out.split(",")]
            reasons = [n for n, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[5:9]) if v.lower().startswith("active")]
            return {"sm_mhz": float(parts[1]), "sm_max_mhz": float(parts[2]), "reasons": sorted(reasons), "samples": 0, "note": "sampled right after the timed region"}
        except Exception:
            return None


def build_workload(n_qubits: int, layers: int, population: int, seed: int):
    from queasars_b200 import genome as gn

    individuals = gn.random_population(n_qubits, layers, population, True, seed)
    circuits = [ind.to_circuit() for ind in individuals]
    params = [list(ind.parameter_values) for ind in individuals]
    return individuals, circuits, params


# ---------------------------------------------------------------------------------------------- CPU baseline
def cpu_prepare(individuals):
    """Lower every individual to the C oracle's gate arrays (outside the timed region: conservative for the CPU arm,
    the reference re-transpiles and re-binds on every call)."""
    from oracle import c_oracle

    prepared = []
    for ind in individuals:
        instr = []
        for inst in ind.to_circuit().data:
            ps = tuple(p.name if hasattr(p, "name") else float(p) for p in inst.operation.params)
            instr.append((inst.operation.name, tuple(q._index for q in inst.qubits), ps))
        prepared.append(c_oracle.lower(instr, list(ind.parameter_values)))
    return prepared


def cpu_evaluate(prepared, table, n_qubits, threads):
    """Oracle port of the reference path on the host cores: simulate each individual with the C/OpenMP oracle
    (oracle/c/statevector.c: one OpenMP team of all host threads per gate, like a compiled CPU simulator), then <H> from
    the precomputed diagonal table (conservative: the reference's estimator evaluates 210 Pauli terms per call instead)."""
    from oracle import c_oracle

    lib = c_oracle.load()
    state = np.empty(1 << n_qubits, dtype=np.complex128)
    out = []
    for targets, controls, mats in prepared:
        out.append(float(lib.oracle_run_circuit(state.ctypes.data, n_qubits, len(targets), targets.ctypes.data, controls.ctypes.data, mats.ctypes.data, table.ctypes.data)))
    return out


def cpu_table(n_qubits):
    from oracle import c_oracle
    from queasars_b200 import genome as gn

    _, z, c = gn.ising_operator(n_qubits).masks()
    return c_oracle.diag_table(n_qubits, [(int(a), float(b.real)) for a, b in zip(z, c)])


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation is Qiskit (not installable offline), so this arm
    times the oracle port of it on the host cores: each step = a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = host_threads()
    sample = POPULATION
    individuals, _, _ = build_workload(N_QUBITS, args.layers, POPULATION, 0)
    prepared = cpu_prepare(individuals[:sample])
    table = cpu_table(N_QUBITS)
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_evaluate(prepared, table, N_QUBITS, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_evaluate(prepared, table, N_QUBITS, threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    desc = f"{sample} of the {POPULATION} individuals per step, C/OpenMP oracle port ({threads} threads per evaluation), diagonal table prebuilt"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    return {
        "workload": f"C2: EVQE population of {POPULATION} individuals x {args.layers} layers, {N_QUBITS}-qubit random diagonal Ising (210 terms), one selection pass per step",
        "n_qubits": N_QUBITS,
        "population_per_gpu": POPULATION,
        "layers": args.layers,
        "precision": "complex128",
        "parallelism": f"population-parallel x{n_gpus} (one population per GPU, no collective)",
        "l2": "working set 32 x 16 MiB = 512 MiB per step > 126 MB L2; no explicit flush",
    }


# ---------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import torch

    # keep stdout clean for the single JSON line: libraries (e.g. NCCL's version banner) write to fd 1 during set-up
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    from queasars_b200 import B200EstimatorV2, B200OperatorCircuitEvaluator
    from queasars_b200 import genome as gn

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    individuals, circuits, params = build_workload(N_QUBITS, args.layers, POPULATION, 1000 * rank)
    operator = gn.ising_operator(N_QUBITS)
    estimator = B200EstimatorV2(device=local_rank, dtype="complex128", coalesce=False)
    evaluator = B200OperatorCircuitEvaluator(estimator, 0.0, operator)
    engine = estimator.engine
    values = evaluator.evaluate_circuits(circuits, params)  # compiles plans, builds the table
    assert len(values) == POPULATION and all(np.isfinite(values))

    plans = [estimator._cache.plan_for(c) for c in circuits]
    ham = estimator.hamiltonian_for(operator)
    batch = engine.resident_batch(plans, ham)
    h2d = batch.set_params(params)
    stream = torch.cuda.ExternalStream(engine.stream, device=torch.device("cuda", local_rank))
    for _ in range(max(3, args.warmup)):
        batch.run()
    resident = batch.read()
    assert np.allclose(resident, values, rtol=0, atol=1e-12)

    # ---- value: device-resident throughput, CUDA events on the launching stream ----
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    launches0 = engine.launch_count
    with ClockSampler(local_rank, enabled=(rank == 0)) as clocks:
        ev0.record(stream)
        for _ in range(args.steps):
            batch.run()
        ev1.record(stream)
        engine.synchronize()
        torch.cuda.synchronize()
    gpu_launches = engine.launch_count - launches0
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_per_step = ms_total / args.steps
    value = world * POPULATION * args.steps / (ms_total * 1e-3)

    # ---- e2e: public evaluator API, host lists in, floats out (H2D of parameters + D2H of results inside) ----
    for _ in range(max(3, args.warmup)):
        evaluator.evaluate_circuits(circuits, params)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = evaluator.evaluate_circuits(circuits, params)
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * POPULATION * args.steps / e2e_s
    assert np.allclose(out, values, rtol=0, atol=1e-12)

    # ---- roofline of the dominant kernel (sweep_kernel<double>), live CUDA events per launch ----
    peak, peak_kind = measured_peak_gbs()
    stats = batch.stats()

    def sweep_bytes_of(b):
        return b.stats()["sweep_bytes"]

    # algorithmic bytes of one step: every read+write sweep moves 2 * 16 B * 2^n per state, the first sweep of a circuit
    # only writes (it synthesises the product-state start), the last sweep of a circuit only reads (the Hamiltonian is
    # diagonal: <H> is accumulated in the sweep's epilogue and the final state, which nothing reads, is not written back),
    # and that epilogue reads the 8 B * 2^n table once per circuit
    tot_ms, tot_bytes, tot_states = 0.0, 0.0, 0
    for _ in range(3):
        ms, states = batch.run_timed()
        tot_ms += float(ms.sum())
        tot_states += int(states.sum())
        tot_bytes += float(sweep_bytes_of(batch) * (states.sum() - 0.5 * states[0] - 0.5 * POPULATION) + POPULATION * 8 * (1 << N_QUBITS))
    sweep_bytes = sweep_bytes_of(batch)
    achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
    # FP64 side of the roofline: DFMA-class instructions the applied gates need (8 per amplitude for a dense 2x2 gate, 7 when
    # its top-left entry is real, half of that for a controlled one; gates folded into the product-state start or dropped on a |0> control cost nothing) against
    # the sustained DFMA issue rate measured on this pool's B200 with tools/fp64_peak.cu (16.9e12 instr/s).
    from queasars_b200 import gate_list as _gl
    from queasars_b200 import schedule as _sc

    dfma = 0.0
    for ind in individuals:
        ops = _gl.from_evqe_individual(ind).ops
        _, remaining = _sc.split_product_prefix(ops, N_QUBITS)
        # a gate without a global phase (every u / cu3 of an EVQE circuit) has a real top-left entry: 7 instead of 8 per amplitude
        per_amp = lambda op: 7.0 if (op.gamma.slot < 0 and op.gamma.const == 0.0) else 8.0  # noqa: E731
        dfma += sum(per_amp(ops[i]) * (1.0 if ops[i].control < 0 else 0.5) for i in remaining) * float(1 << N_QUBITS)
    fp64_peak = 16.9e12
    fp64_rate = dfma / (ms_per_step * 1e-3)
    roofline = {
        "bound": "hbm",
        "kernel": "qb::sweep_kernel<double>",
        "achieved": achieved,
        "peak": peak,
        "peak_kind": peak_kind,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": profiled_traffic(),
        "algorithmic_bytes_per_launch": tot_bytes / max(1, 3 * stats["sweep_launches"]),
        "bytes_per_statevector_sweep": sweep_bytes,
        "sweeps_per_evaluation": stats["state_sweeps"] / POPULATION,
        "gates_per_sweep": float(np.mean([p.n_ops for p in plans])) / (stats["state_sweeps"] / POPULATION),
        "sweep_share_of_step": (tot_ms / 3) / ms_per_step,
        "fp64": {"dfma_instr_per_s": fp64_rate, "peak_measured": fp64_peak, "frac": fp64_rate / fp64_peak, "unit": "DFMA-class instr/s"},
        "note": "sweeps of this workload fuse ~16 applied fp64 gates each: the binding roof is FP64 issue (see fp64), not HBM; gate_apply reports the HBM-bound regime at 26-30 q",
    }

    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8 * POPULATION, "ms_per_step": 1e3 * e2e_s / args.steps},
        "gpu_launches": int(gpu_launches),
        "roofline": roofline,
        "clocks": clocks.summary(),
    }

    if rank == 0 and world == 1 and not args.skip_extras:
        line["gate_apply"] = gate_apply_probe(engine, estimator, peak, args)
        line["cpu_baseline"] = cpu_baseline_leg(individuals, args, values)
    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(json_fd, 1)
    os.close(json_fd)
    if rank == 0:
        print(json.dumps(line), flush=True)


def gate_apply_probe(engine, estimator, peak, args):
    """The sweep kernel on ONE HBM-resident state of 26 / 28 / 30 qubits (1 / 4 / 16 GiB), two regimes:
      fused_evqe   a random EVQE individual: ~17 gates fused per sweep -> FP64-issue bound by design
      hbm_regime   layers of 7 ``u`` gates on 7 non-low qubits: one read+write sweep per layer -> HBM bound
    GB/s = algorithmic sweep bytes (2 * 16 B * 2^n) / CUDA-event time of each read+write sweep launch; the first sweep
    of a circuit (product-state start, write only) is excluded."""
    from queasars_b200 import gate_list as gl
    from queasars_b200 import genome as gn
    from queasars_b200.circuit import QuantumCircuit

    def measure(plan, params, reps):
        rb = engine.resident_batch([plan], None)
        rb.set_params([params])
        for _ in range(2):
            rb.run()
        tot_ms, all_ms, launches = 0.0, 0.0, 0
        for _ in range(reps):
            ms, _states = rb.run_timed()
            tot_ms += float(ms[1:].sum())
            all_ms += float(ms.sum())
            launches += len(ms) - 1
        bytes_per = rb.stats()["sweep_bytes"]
        rb.close()
        gbs = bytes_per * launches / (tot_ms * 1e-3) / 1e9 if launches else None
        # ms_per_circuit = all sweeps of the circuit (incl. the write-only product-state sweep): what fusing more gates per sweep buys
        return {"sweeps_rw": launches // reps, "ms_per_sweep": tot_ms / max(1, launches), "ms_per_circuit": all_ms / reps, "GBps": gbs,
                "frac_of_measured_hbm": gbs / peak if gbs else None}

    out = {}
    for n, layers in ((26, args.layers), (28, 4), (30, 4)):
        ind = gn.Individual.random(n, layers, True, 7)
        plan = engine.compile(gl.from_evqe_individual(ind))
        fused = measure(plan, list(ind.parameter_values), 3)
        n_u = sum(1 for layer in ind.layers for g in layer.gates if type(g).__name__ == "Rotation")
        n_cu3 = sum(1 for layer in ind.layers for g in layer.gates if type(g).__name__ == "ControlledRotation")
        fused.update(layers=layers, gates=plan.n_ops, gates_per_sweep=plan.n_ops / plan.n_sweeps)
        # what the same gates would move if applied one sweep per gate (SURVEY.md 8d "effective GB/s")
        total_ms = fused["ms_per_sweep"] * fused["sweeps_rw"]
        fused["effective_unfused_GBps"] = (2 * n_u + n_cu3) * 16 * (1 << n) / (total_ms * 1e-3) / 1e9 if total_ms else None
        circ = QuantumCircuit(n)
        for q in range(n):
            circ.u(0.1 + 0.01 * q, 0.2, 0.3, q)  # absorbed into the product-state start
        for layer in range(3):
            for g in range(7):
                circ.u(0.3 + g, 0.2 * layer, 0.1, 4 + ((g * 3 + 7 * layer) % (n - 4)))
        plan = engine.compile(gl.from_circuit(circ))
        hbm = measure(plan, [], 3)
        hbm.update(gates_per_sweep=21 / max(1, plan.n_sweeps - 1))
        out[f"{n}q"] = {"fused_evqe": fused, "hbm_regime": hbm}
    return out


def cpu_baseline_leg(individuals, args, gpu_values):
    threads = host_threads()
    sample = len(individuals)
    table = cpu_table(N_QUBITS)
    prepared = cpu_prepare(individuals[:sample])
    t0 = time.perf_counter()
    vals = cpu_evaluate(prepared, table, N_QUBITS, threads)
    reps = 1
    while time.perf_counter() - t0 < 10.0 and reps < 8:
        cpu_evaluate(prepared, table, N_QUBITS, threads)
        reps += 1
    dt = time.perf_counter() - t0
    err = float(np.max(np.abs(np.asarray(vals) - np.asarray(gpu_values[:sample])) / np.maximum(1.0, np.abs(vals))))
    return {
        "value": sample * reps / dt,
        "unit": UNIT,
        "cores": threads,
        "kind": "port",
        "sample": f"{reps} x {sample} of the {POPULATION} individuals, C/OpenMP oracle port of the Qiskit statevector estimator ({threads} threads per evaluation), diagonal table prebuilt",
        "max_rel_err_gpu_vs_oracle": err,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--skip-extras", action="store_true", help="skip the 26/28-qubit gate-apply probe and the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
