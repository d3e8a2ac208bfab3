"""Shared driver of the UNMODIFIED reference EVQE loop (``queasars.minimum_eigensolvers.evqe``) for the CPU test (oracle
primitives) and the GPU test (B200 primitives).  The reference package is imported from /root/reference in the build container
and from the git-ignored ``baseline/_ref`` copy on the GPU box (tools/stage_reference.py); its Qiskit imports are served by the
stand-in modules of ``queasars_b200.qiskit_compat`` (Qiskit is not installable offline).  Mirrors the reference's own end-to-end
test: test/minimum_eigensolvers/evqe/solver.py:17-53, test_evqe_algorithm.py:23-38."""
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
CANDIDATES = ("/root/reference", os.path.join(ROOT, "baseline", "_ref"))


def locate_reference():
    for path in CANDIDATES:
        if os.path.isfile(os.path.join(path, "queasars", "__init__.py")):
            return path
    return None


def import_reference(path=None):
    """-> namespace of the reference classes the tests need (imports the unmodified package)."""
    from queasars_b200 import qiskit_compat

    path = locate_reference() if path is None else path
    if path is None:
        raise RuntimeError("reference package not found")
    qiskit_compat.install()
    if path not in sys.path:
        sys.path.insert(0, path)
    from queasars.circuit_evaluation.configured_primitives import ConfiguredEstimatorV2, ConfiguredSamplerV2
    from queasars.job_shop_scheduling.domain_wall_hamiltonian_encoder import JSSPDomainWallHamiltonianEncoder
    from queasars.job_shop_scheduling.problem_instances import Job, JobShopSchedulingProblemInstance, Machine, Operation
    from queasars.minimum_eigensolvers.base.termination_criteria import BestIndividualRelativeChangeTolerance
    from queasars.minimum_eigensolvers.evqe.evqe import EVQEMinimumEigensolver, EVQEMinimumEigensolverConfiguration
    from queasars.utility.spsa_termination import SPSATerminationChecker

    return dict(locals(), path=path)


def test_model_hamiltonian():
    """min x^2 - y^2, x, y in [0, 3] in Ising form (test/minimum_eigensolvers/evqe/model.py:10-23):
    ground state '1100' (x = 0, y = 3), E = -9."""
    from queasars_b200.operators import SparsePauliOp

    return SparsePauliOp.from_list([("IIIZ", -1.5), ("IIZI", -3.0), ("IIZZ", 1.0), ("IZII", 1.5), ("ZIII", 3.0), ("ZZII", -1.0)])


def sample_solver(ref, estimator, sampler, executor, mutex, max_generations=None):
    """create_sample_solver of test/minimum_eigensolvers/evqe/solver.py:17-53 with the given primitives."""
    from qiskit_algorithms.optimizers import NFT

    configuration = ref["EVQEMinimumEigensolverConfiguration"](
        configured_sampler=ref["ConfiguredSamplerV2"](sampler=sampler, shots=1000),
        configured_estimator=ref["ConfiguredEstimatorV2"](estimator=estimator, precision=0.05),
        pass_manager=None,
        optimizer=NFT(maxiter=40),
        optimizer_n_circuit_evaluations=40,
        max_generations=max_generations,
        max_circuit_evaluations=None,
        termination_criterion=None if max_generations else ref["BestIndividualRelativeChangeTolerance"](minimum_relative_change=0.005),
        random_seed=0,
        population_size=10,
        randomize_initial_population_parameters=False,
        speciation_genetic_distance_threshold=3,
        selection_alpha_penalty=0.1,
        selection_beta_penalty=0.1,
        parameter_search_probability=0.24,
        topological_search_probability=0.2,
        layer_removal_probability=0.05,
        parallel_executor=executor,
        mutually_exclusive_primitives=mutex,
    )
    return ref["EVQEMinimumEigensolver"](configuration=configuration)


def jssp_4q(ref):
    """The 2 jobs / 2 machines instance of examples/evqe_jssp_small_examples.ipynb cells 4 and 8 (4 qubits; notebook result:
    best CVaR(0.5) objective 63.5, states 0011 / 1100)."""
    machines = (ref["Machine"](name="m0"), ref["Machine"](name="m1"))
    Op, Job = ref["Operation"], ref["Job"]
    job0 = Job(name="j0", operations=(Op(name="j0op0", machine=machines[0], processing_duration=1, job_name="j0"),
                                      Op(name="j0op1", machine=machines[1], processing_duration=1, job_name="j0")))
    job1 = Job(name="j1", operations=(Op(name="j1op1", machine=machines[0], processing_duration=1, job_name="j1"),
                                      Op(name="j1op2", machine=machines[1], processing_duration=1, job_name="j1")))
    instance = ref["JobShopSchedulingProblemInstance"](name="2_jobs_2_machines_seed_121", machines=machines, jobs=(job0, job1))
    encoder = ref["JSSPDomainWallHamiltonianEncoder"](jssp_instance=instance, makespan_limit=3, max_opt_value=100, opt_all_operations_share=0.19,
                                                      encoding_penalty=319, overlap_constraint_penalty=319, precedence_constraint_penalty=275)
    return encoder, encoder.get_problem_hamiltonian()


JSSP_INSTANCES = {
    # name: (jobs as ((machine, duration), ...), makespan_limit, n_qubits, notebook convergence value = minimum diagonal energy)
    # (second-lowest energies, SURVEY.md section 8c-2: 338.5 / 92.4 / 52.125)
    "4q": (((("m0", 1), ("m1", 1)), (("m0", 1), ("m1", 1))), 3, 4, 63.5),  # evqe_jssp_small_examples.ipynb cells 4, 8, 14
    "5q": (((("m0", 1), ("m1", 2)), (("m0", 1), ("m1", 1), ("m2", 1))), 4, 5, 61.6),  # same notebook, cells 23, 27, 33
    "8q": (((("m0", 2), ("m1", 1)), (("m0", 1), ("m1", 2))), 5, 8, 22.75),  # using_the_ibm_runtime.ipynb cells 2, 6, 8
}


def jssp_instance(ref, which):
    """The small JSSP instances of the reference's example notebooks (BASELINE config C1) -> (encoder, Hamiltonian, expected
    minimum).  Penalties as in every notebook."""
    spec, limit, n_qubits, best = JSSP_INSTANCES[which]
    names = sorted({m for job in spec for m, _ in job})
    machines = {name: ref["Machine"](name=name) for name in names}
    jobs = []
    for j, ops in enumerate(spec):
        jobs.append(ref["Job"](name=f"j{j}", operations=tuple(
            ref["Operation"](name=f"j{j}op{i}", machine=machines[m], processing_duration=d, job_name=f"j{j}") for i, (m, d) in enumerate(ops))))
    instance = ref["JobShopSchedulingProblemInstance"](name=which, machines=tuple(machines[n] for n in names), jobs=tuple(jobs))
    encoder = ref["JSSPDomainWallHamiltonianEncoder"](jssp_instance=instance, makespan_limit=limit, max_opt_value=100, opt_all_operations_share=0.19,
                                                      encoding_penalty=319, overlap_constraint_penalty=319, precedence_constraint_penalty=275)
    assert encoder.n_qubits == n_qubits
    return encoder, encoder.get_problem_hamiltonian(), best


def jssp_solver(ref, sampler, executor, random_seed=0):
    """Sampler-only CVaR(0.5) configuration of examples/evqe_jssp_small_examples.ipynb cell 10."""
    from qiskit_algorithms.optimizers import SPSA

    checker = ref["SPSATerminationChecker"](minimum_relative_change=0.01, allowed_consecutive_violations=2)
    configuration = ref["EVQEMinimumEigensolverConfiguration"](
        configured_sampler=ref["ConfiguredSamplerV2"](sampler=sampler, shots=512),
        configured_estimator=None,
        pass_manager=None,
        distribution_alpha_tail=0.5,
        optimizer=SPSA(maxiter=33, perturbation=0.35, learning_rate=0.43, trust_region=True, last_avg=1, resamplings=1, termination_checker=checker.termination_check),
        optimizer_n_circuit_evaluations=66,
        max_generations=None,
        max_circuit_evaluations=None,
        termination_criterion=ref["BestIndividualRelativeChangeTolerance"](minimum_relative_change=0.01, allowed_consecutive_violations=1),
        random_seed=random_seed,
        population_size=10,
        n_initial_layers=2,
        randomize_initial_population_parameters=True,
        speciation_genetic_distance_threshold=1,
        use_tournament_selection=True,
        tournament_size=2,
        selection_alpha_penalty=0.15,
        selection_beta_penalty=0.02,
        parameter_search_probability=0.39,
        topological_search_probability=0.79,
        layer_removal_probability=0.02,
        parallel_executor=executor,
        mutually_exclusive_primitives=False,
    )
    return ref["EVQEMinimumEigensolver"](configuration=configuration)


JSSP_SECOND_LEVEL = {"4q": 338.5, "5q": 92.4, "8q": 52.125}


def likeliest_bitstring(result) -> str:
    probs = result.eigenstate.binary_probabilities()
    return max(probs, key=probs.get)
