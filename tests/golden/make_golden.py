"""Generate the golden fixtures in tests/golden/ by running the REFERENCE's own pure-Python modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Qiskit is not installable here, so the reference modules are imported with the attribute-compatible
stand-ins from ``queasars_b200.qiskit_compat`` (container types only -- no simulation arithmetic).
What is produced is therefore exactly what the reference's code computes for:
  * jssp_hamiltonians.json -- JSSPDomainWallHamiltonianEncoder.get_problem_hamiltonian() raw term lists
    (domain_wall_hamiltonian_encoder.py:87-104, 189-230) for the notebook instances (4/5/8/12 qubits), the
    reference's unit-test instance, and the 26-qubit C4 instance of SURVEY.md section 8d
  * genomes.json -- EVQEPopulation.random_population / EVQEIndividual.random_individual genomes, parameter
    values, circuit op lists (name, qubits, parameter names) and ``circuit.parameters`` order
  * cvar.json -- outputs of expectation_calculation._get_expectation on seeded inputs
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))

from queasars_b200 import qiskit_compat  # noqa: E402

qiskit_compat.install()
sys.path.insert(0, "/root/reference")

from queasars.job_shop_scheduling.problem_instances import Machine, Operation, Job, JobShopSchedulingProblemInstance  # noqa: E402
from queasars.job_shop_scheduling.domain_wall_hamiltonian_encoder import JSSPDomainWallHamiltonianEncoder  # noqa: E402
from queasars.minimum_eigensolvers.evqe.evolutionary_algorithm.individual import EVQEIndividual  # noqa: E402
from queasars.minimum_eigensolvers.evqe.evolutionary_algorithm.population import EVQEPopulation  # noqa: E402
from queasars.minimum_eigensolvers.evqe.quantum_circuit.quantum_gate import (  # noqa: E402
    IdentityGate,
    RotationGate,
    ControlGate,
    ControlledRotationGate,
)
from queasars.circuit_evaluation.expectation_calculation import _get_expectation  # noqa: E402


def instance(name, machine_names, jobs):
    machines = {m: Machine(name=m) for m in machine_names}
    job_objs = []
    for jname, ops in jobs:
        operations = tuple(
            Operation(name=f"{jname}op{i}", machine=machines[m], processing_duration=d, job_name=jname)
            for i, (m, d) in enumerate(ops)
        )
        job_objs.append(Job(name=jname, operations=operations))
    return JobShopSchedulingProblemInstance(name=name, machines=tuple(machines.values()), jobs=tuple(job_objs))


NOTEBOOK = dict(max_opt_value=100, opt_all_operations_share=0.19, encoding_penalty=319, overlap_constraint_penalty=319, precedence_constraint_penalty=275)

INSTANCES = {
    # examples/evqe_jssp_small_examples.ipynb cells 4+8
    "jssp_4q": (instance("2_jobs_2_machines_seed_121", ["m0", "m1"], [("j0", [("m0", 1), ("m1", 1)]), ("j1", [("m0", 1), ("m1", 1)])]), 3, NOTEBOOK),
    # examples/evqe_jssp_small_examples.ipynb cells 23+27
    "jssp_5q": (instance("2_jobs_2_machines_asymmetric", ["m0", "m1", "m2"], [("j0", [("m0", 1), ("m1", 2)]), ("j1", [("m0", 1), ("m1", 1), ("m2", 1)])]), 4, NOTEBOOK),
    # examples/using_the_ibm_runtime.ipynb cells 2+6
    "jssp_8q": (instance("Simple Instance", ["m0", "m1"], [("j0", [("m0", 2), ("m1", 1)]), ("j1", [("m0", 1), ("m1", 2)])]), 5, NOTEBOOK),
    # examples/evqe_jssp_optimization.ipynb cells 2+6
    "jssp_12q": (instance("2_jobs_3_machines_seed_121", ["m0", "m1", "m2"], [("j0", [("m2", 1), ("m0", 1), ("m1", 2)]), ("j1", [("m2", 2), ("m0", 1), ("m1", 1)])]), 6, NOTEBOOK),
    # test/job_shop_scheduling/problem_instance.py:15-27 with the encoder defaults used by the unit tests
    "jssp_unit_test": (instance("instance", ["m1", "m2"], [("j1", [("m1", 1), ("m2", 1)]), ("j2", [("m2", 1), ("m1", 1)])]), 4, {}),
    # SURVEY.md section 8d, config C4 (26 qubits)
    "jssp_26q": (
        instance(
            "c4_3_jobs_5_machines",
            ["m0", "m1", "m2", "m3", "m4"],
            [
                ("j0", [("m0", 1), ("m1", 1), ("m2", 1), ("m3", 1), ("m4", 2)]),
                ("j1", [("m1", 2), ("m0", 1), ("m3", 2), ("m2", 1)]),
                ("j2", [("m2", 1), ("m4", 2), ("m0", 1), ("m1", 2)]),
            ],
        ),
        8,
        NOTEBOOK,
    ),
}


def energy(state, zs, cs):
    return sum(c * (1 - 2 * (bin(state & z).count("1") & 1)) for z, c in zip(zs, cs))


def jssp_fixtures():
    out = {}
    for key, (inst, limit, kwargs) in INSTANCES.items():
        enc = JSSPDomainWallHamiltonianEncoder(jssp_instance=inst, makespan_limit=limit, **kwargs)
        ham = enc.get_problem_hamiltonian()
        x, z, c = ham.masks()
        assert not x.any()
        n = enc.n_qubits
        zs = [int(v) for v in z]
        cs = [float(v.real) for v in c]
        entry = {"n_qubits": n, "makespan_limit": limit, "z_masks": zs, "coeffs": cs, "n_raw_terms": len(zs)}
        merged = {}
        for zm, cf in zip(zs, cs):
            merged[zm] = merged.get(zm, 0.0) + cf
        entry["n_distinct_terms"] = len(merged)
        if n <= 12:
            energies = [energy(k, zs, cs) for k in range(1 << n)]
            order = sorted(range(1 << n), key=lambda k: (energies[k], k))
            entry["lowest"] = [[format(k, f"0{n}b"), energies[k]] for k in order[:6]]
            entry["energy_sum"] = sum(energies)
            entry["energies"] = energies if n <= 8 else None
        else:
            rng = random.Random(7)
            probe = [rng.getrandbits(n) for _ in range(64)]
            entry["probe_states"] = probe
            entry["probe_energies"] = [energy(k, zs, cs) for k in probe]
        out[key] = entry
    return out


def gene(gate):
    if isinstance(gate, IdentityGate):
        return ["id"]
    if isinstance(gate, RotationGate):
        return ["rot"]
    if isinstance(gate, ControlGate):
        return ["ctrl", gate.controlled_qubit_index]
    if isinstance(gate, ControlledRotationGate):
        return ["crot", gate.control_qubit_index]
    raise TypeError(gate)


def circuit_ops(circuit):
    ops = []
    for inst in circuit.data:
        params = [p.name if hasattr(p, "name") else float(p) for p in inst.operation.params]
        ops.append([inst.operation.name, [circuit.find_bit(q).index for q in inst.qubits], params])
    return ops


def individual_entry(ind, partial_layers=None):
    entry = {
        "n_qubits": ind.n_qubits,
        "layers": [[gene(g) for g in layer.gates] for layer in ind.layers],
        "parameter_values": list(ind.parameter_values),
    }
    full = ind.get_parameterized_quantum_circuit()
    entry["full_ops"] = circuit_ops(full)
    entry["full_parameters"] = [p.name for p in full.parameters]
    if partial_layers is not None:
        part = ind.get_partially_parameterized_quantum_circuit(set(partial_layers))
        entry["partial_layers"] = list(partial_layers)
        entry["partial_ops"] = circuit_ops(part)
        entry["partial_parameters"] = [p.name for p in part.parameters]
        entry["partial_layer_values"] = list(ind.get_layer_parameter_values(partial_layers[0]))
    return entry


def genome_fixtures():
    out = {}
    pop = EVQEPopulation.random_population(n_qubits=4, n_layers=2, n_individuals=10, randomize_parameter_values=True, random_seed=0)
    out["population_4q_2l_seed0"] = [individual_entry(ind, [-1]) for ind in pop.individuals]
    out["individual_4q_2l_seed0"] = individual_entry(EVQEIndividual.random_individual(4, 2, False, 0))
    pop12 = EVQEPopulation.random_population(n_qubits=12, n_layers=3, n_individuals=3, randomize_parameter_values=True, random_seed=11)
    out["population_12q_3l_seed11"] = [individual_entry(ind, [1]) for ind in pop12.individuals]
    ind13 = EVQEIndividual.random_individual(3, 12, True, 5)  # >= 11 layers: layer10 < layer1 < layer2 ordering quirk
    out["individual_3q_12l_seed5"] = individual_entry(ind13, [-1])
    pop20 = EVQEPopulation.random_population(n_qubits=20, n_layers=2, n_individuals=2, randomize_parameter_values=True, random_seed=0)
    out["population_20q_2l_seed0_genes_only"] = [
        {"layers": [[gene(g) for g in layer.gates] for layer in ind.layers], "parameter_values": list(ind.parameter_values)}
        for ind in pop20.individuals
    ]
    return out


def cvar_fixtures():
    rng = random.Random(3)
    cases = []
    for alpha in (1.0, 0.5, 0.25, 0.999999, 0.1):
        for n_states in (1, 3, 17):
            probs = [rng.random() for _ in range(n_states)]
            tot = sum(probs)
            probs = [p / tot for p in probs]
            vals = [round(rng.uniform(-5, 5), 1) for _ in range(n_states)]  # rounded -> ties occur
            state_list = [(i, p, v) for i, (p, v) in enumerate(zip(probs, vals))]
            cases.append({"alpha": alpha, "probs": probs, "values": vals, "expected": _get_expectation(state_list, alpha)})
    return cases


if __name__ == "__main__":
    for name, payload in (("jssp_hamiltonians.json", jssp_fixtures()), ("genomes.json", genome_fixtures()), ("cvar.json", cvar_fixtures())):
        with open(os.path.join(HERE, name), "w") as fh:
            json.dump(payload, fh, indent=1)
        print("wrote", name)
