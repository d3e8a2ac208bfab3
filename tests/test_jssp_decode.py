"""``jssp_decode.decode_start_times`` against the reference's own ``translate_result_bitstring`` (run here from the reference
package: /root/reference in the build container, baseline/_ref on the GPU box) for EVERY basis state of the 8-qubit instance of
examples/using_the_ibm_runtime.ipynb and a random sample of the 26-qubit C4 instance."""
import numpy as np
import pytest

from tests import reference_loop

pytestmark = pytest.mark.skipif(reference_loop.locate_reference() is None, reason="reference package not present")


def _instance(ref, spec, limit):
    Machine, Operation, Job = ref["Machine"], ref["Operation"], ref["Job"]
    machines = {name: Machine(name=name) for name in sorted({m for job in spec for m, _ in job})}
    jobs = []
    for j, ops in enumerate(spec):
        jobs.append(Job(name=f"j{j}", operations=tuple(Operation(name=f"j{j}op{i}", machine=machines[m], processing_duration=d, job_name=f"j{j}") for i, (m, d) in enumerate(ops))))
    inst = ref["JobShopSchedulingProblemInstance"](name="inst", machines=tuple(machines.values()), jobs=tuple(jobs))
    return ref["JSSPDomainWallHamiltonianEncoder"](jssp_instance=inst, makespan_limit=limit, max_opt_value=100, opt_all_operations_share=0.19,
                                                   encoding_penalty=319, overlap_constraint_penalty=319, precedence_constraint_penalty=275)


def _reference_times(encoder, state, n):
    result = encoder.translate_result_bitstring(format(int(state), f"0{n}b"))
    row = []
    for job in encoder.jssp_instance.jobs:
        for op in result.schedule[job]:
            row.append(getattr(op, "start_time", None) if getattr(op, "is_scheduled", True) and hasattr(op, "start_time") else None)
    return [-1 if v is None else int(v) for v in row]


@pytest.mark.parametrize("spec,limit,exhaustive", [
    ([[("m0", 1), ("m1", 1)], [("m0", 1), ("m1", 1)]], 3, True),
    ([[("m0", 2), ("m1", 1)], [("m1", 1), ("m0", 2)]], 5, True),
    ([[("m0", 1), ("m1", 1), ("m2", 1), ("m3", 1), ("m4", 2)], [("m1", 2), ("m0", 1), ("m3", 2), ("m2", 1)], [("m2", 1), ("m4", 2), ("m0", 1), ("m1", 2)]], 8, False),
])
def test_vectorised_decode_matches_reference(spec, limit, exhaustive):
    from queasars_b200 import jssp_decode

    ref = reference_loop.import_reference()
    encoder = _instance(ref, spec, limit)
    n = encoder.n_qubits
    layout = jssp_decode.layout_of(encoder)
    assert sum(s.width for s in layout) == n
    states = np.arange(1 << n, dtype=np.uint64) if exhaustive else np.random.default_rng(0).integers(0, 1 << n, size=300, dtype=np.uint64)
    if not exhaustive:
        assert n == 26  # the C4 instance of SURVEY.md section 8d
        # make sure valid schedules are among the sample: all-walls-at-zero and a few thermometer codes
        valid = [0]
        for s in layout:
            valid.append(valid[-1] | (((1 << (s.width // 2)) - 1) << s.start))
        states = np.concatenate([states, np.asarray(valid, dtype=np.uint64)])
    got = jssp_decode.decode_start_times(states, layout)
    for row, state in zip(got, states):
        assert list(row) == _reference_times(encoder, state, n)
    assert np.any(np.all(got >= 0, axis=1))
    if max(s.width for s in layout) > 1:  # one-qubit variables have no invalid pattern
        assert np.any(got < 0)
