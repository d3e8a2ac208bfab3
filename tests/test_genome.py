"""queasars_b200.genome (product-side genome types / synthetic populations) pinned against the reference's own
EVQEPopulation / EVQEIndividual output (tests/golden/genomes.json)."""
import pytest

from queasars_b200 import gate_list as gl
from queasars_b200 import genome as gn


def _genes(ind):
    out = []
    for layer in ind.layers:
        row = []
        for g in layer.gates:
            if isinstance(g, gn.Identity):
                row.append(["id"])
            elif isinstance(g, gn.Rotation):
                row.append(["rot"])
            elif isinstance(g, gn.Control):
                row.append(["ctrl", g.controlled_qubit_index])
            else:
                row.append(["crot", g.control_qubit_index])
        out.append(row)
    return out


def _ops(circuit):
    ops = []
    for inst in circuit.data:
        params = [p.name if hasattr(p, "name") else float(p) for p in inst.operation.params]
        ops.append([inst.operation.name, [circuit.find_bit(q).index for q in inst.qubits], params])
    return ops


@pytest.mark.parametrize("key,args", [("population_4q_2l_seed0", (4, 2, 10, True, 0)), ("population_12q_3l_seed11", (12, 3, 3, True, 11))])
def test_population_and_circuits_match_reference(genome_golden, key, args):
    pop = gn.random_population(*args)
    for ind, entry in zip(pop, genome_golden[key]):
        assert _genes(ind) == entry["layers"]
        assert list(ind.parameter_values) == entry["parameter_values"]
        full = ind.to_circuit()
        assert _ops(full) == entry["full_ops"]
        assert [p.name for p in full.parameters] == entry["full_parameters"]
        part = ind.to_circuit(set(entry["partial_layers"]))
        assert _ops(part) == entry["partial_ops"]
        assert [p.name for p in part.parameters] == entry["partial_parameters"]
        assert list(ind.layer_values(entry["partial_layers"][0])) == entry["partial_layer_values"]
        for layers in (None, set(entry["partial_layers"])):
            assert gl.from_evqe_individual(ind, layers).ops == gl.from_circuit(ind.to_circuit(layers)).ops


def test_20q_population_and_many_layer_quirk(genome_golden):
    for ind, entry in zip(gn.random_population(20, 2, 2, True, 0), genome_golden["population_20q_2l_seed0_genes_only"]):
        assert _genes(ind) == entry["layers"] and list(ind.parameter_values) == entry["parameter_values"]
    ind = gn.Individual.random(3, 12, True, 5)
    entry = genome_golden["individual_3q_12l_seed5"]
    assert _genes(ind) == entry["layers"]
    assert [p.name for p in ind.to_circuit().parameters] == entry["full_parameters"]
    direct = gl.from_evqe_individual(ind)
    assert list(direct.param_names) == entry["full_parameters"]


def test_synthetic_operators():
    op = gn.ising_operator(20)
    assert op.num_qubits == 20 and len(op) == 210 and op.is_diagonal()
    t = gn.tfim_operator(24)
    assert len(t) == 47 and not t.is_diagonal()
