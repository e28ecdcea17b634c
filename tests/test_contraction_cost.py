"""The reference's cost-model KATs (tnc/src/contractionpath/contraction_cost.rs:366-460 and its doc-tests).
These formulas are the FLOP accounting of the benchmark and the objective of the partition refiner."""
from tnc_b200.contractionpath import ContractionPath, path
from tnc_b200.contractionpath.contraction_cost import (communication_path_cost, communication_path_op_costs,
                                                       compute_memory_requirements, contract_cost_tensors,
                                                       contract_op_cost_tensors, contract_path_cost,
                                                       contract_size_tensors, contract_size_tensors_exact)
from tnc_b200.tensornetwork import Tensor

BD = {0: 5, 1: 2, 2: 6, 3: 8, 4: 1, 5: 3, 6: 4, 7: 3, 8: 2, 9: 2}


def setup_simple():   # :326-334
    return [Tensor.new_from_map([4, 3, 2], BD), Tensor.new_from_map([0, 1, 3, 2], BD), Tensor.new_from_map([4, 5, 6], BD)]


def setup_complex():  # :336-363
    t2 = Tensor.new_composite([Tensor.new_from_map([5, 6, 8], BD), Tensor.new_from_map([7, 8, 9], BD)])
    return [Tensor.new_composite(setup_simple()), t2]


def setup_parallel():  # :365-374 (numbered :354-364 in the file)
    return setup_simple() + [Tensor.new_from_map([5, 6], BD)]


def test_doc_examples():
    bd = {0: 5, 1: 7, 2: 9, 3: 11, 4: 13}
    t1, t2 = Tensor.new_from_map([0, 1, 2], bd), Tensor.new_from_map([2, 3, 4], bd)
    assert contract_cost_tensors(t1, t2) == 350350.0          # :15-25
    assert contract_op_cost_tensors(t1, t2) == 45045.0        # :37-47
    assert contract_size_tensors(t1, t2) == 6607.0            # :59-69
    bd = {0: 5, 1: 7, 2: 9, 3: 11}
    assert contract_size_tensors_exact(Tensor.new_from_map([0, 1, 2], bd), Tensor.new_from_map([3, 2], bd)) == 12784.0  # :84-94


def test_contract_path_cost():   # :376-386
    tn = setup_simple()
    assert contract_path_cost(tn, path((0, 1), (0, 2)), False) == (4540.0, 538.0)
    assert contract_path_cost(tn, path((0, 2), (0, 1)), False) == (49296.0, 1176.0)


def test_contract_complex_path_cost():   # :388-399
    p = path((0, 1), nested={0: [(0, 1), (0, 2)], 1: [(0, 1)]})
    assert contract_path_cost(setup_complex(), p, False) == (11188.0, 538.0)
    assert contract_path_cost(setup_complex(), p, True) == (1464.0, 538.0)   # :413-424


def test_contract_path_cost_only_ops():   # :401-411
    tn = setup_simple()
    assert contract_path_cost(tn, path((0, 1), (0, 2)), True) == (600.0, 538.0)
    assert contract_path_cost(tn, path((0, 2), (0, 1)), True) == (6336.0, 1176.0)
    assert compute_memory_requirements(tn, path((0, 2), (0, 1))) == 1176.0


def test_communication_path_cost():   # :426-460
    tn = setup_parallel()
    assert communication_path_cost(tn, [(0, 1), (2, 3), (0, 2)], True, True, None) == (490.0, 538.0)
    assert communication_path_cost(tn, [(0, 1), (2, 3), (0, 1)], False, True, None) == (7564.0, 538.0)
    tc = [20.0, 30.0, 80.0, 10.0]
    assert communication_path_cost(tn, [(0, 1), (2, 3), (0, 2)], True, True, tc) == (520.0, 538.0)
    assert communication_path_cost(tn, [(0, 1), (2, 3), (0, 1)], False, True, tc) == (7594.0, 538.0)
    (par, ser), mem = communication_path_op_costs(tn, [(0, 1), (2, 3), (0, 2)], True, tc)
    assert par == 520.0 and ser >= par and mem == 538.0
    assert communication_path_cost(tn[:1], [], True, True, [7.0]) == (7.0, 7.0)   # single input :239-241


def test_bench_flop_accounting_matches_reference_formula():
    """8MNK (tncb_plan_info) vs contract_cost_tensors = (8K-2)*MN: they differ by exactly 2 per output element."""
    a, b = Tensor([0, 1, 2], [4, 8, 16]), Tensor([2, 3], [16, 32])
    M, N, K = 32, 32, 16
    assert contract_cost_tensors(a, b) == 8.0 * M * N * K - 2.0 * M * N
