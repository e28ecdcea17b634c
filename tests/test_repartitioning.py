"""Partition evaluation / refinement (planning, metadata only): mirrors of
tnc::contractionpath::repartitioning::compute_solution and the simulated-annealing balancer."""
import numpy as np

from tnc_b200.builders import random_circuit
from tnc_b200.contractionpath import validate_path
from tnc_b200.contractionpath.repartitioning import balance_partitions, communication_path_op_costs, compute_solution
from tnc_b200.tensornetwork import Tensor
from tnc_b200.tensornetwork.partitioning import find_partitioning, partition_tensor_network


def test_partition_tensor_network_kat():
    """tensornetwork/partitioning.rs:165-175 semantics: one composite per id, in order of first appearance."""
    ts = [Tensor.new_from_const([i, i + 1], 2) for i in range(6)]
    ptn = partition_tensor_network(Tensor.new_composite(ts), [2, 1, 2, 0, 0, 1])
    assert [[t.legs for t in c.tensors] for c in ptn.tensors] == [[[0, 1], [2, 3]], [[1, 2], [5, 6]], [[3, 4], [4, 5]]]


def test_communication_costs():
    """contraction_cost.rs:196-281: critical path takes max of the operand latencies, serial the sum."""
    a, b, c = Tensor([0, 1], [2, 3]), Tensor([1, 2], [3, 4]), Tensor([2, 0], [4, 2])
    (par, ser), mem = communication_path_op_costs([a, b, c], [(0, 1), (0, 2)], [10.0, 20.0, 5.0])
    assert par == (2 * 4) + max((2 * 3 * 4) + max(10.0, 20.0), 5.0)
    assert ser == (2 * 4) + ((2 * 3 * 4) + 10.0 + 20.0) + 5.0
    assert mem == max(6 + 12 + 8, 8 + 8 + 1)


def test_compute_solution_and_sa_are_deterministic_and_monotone():
    tn = random_circuit(12, 6, 0.5, 0.5, np.random.default_rng(5))
    init = find_partitioning(tn, 3, seed=1)
    assert sorted(set(init)) == [0, 1, 2] and len(init) == len(tn.tensors)
    ptn, path, par, ser = compute_solution(tn, init)
    assert len(ptn.tensors) == 3 and set(path.nested) == {0, 1, 2} and len(path.toplevel) == 2
    assert validate_path(path) and 0 < par <= ser
    best1, s1 = balance_partitions(tn, 3, init, steps=160, seed=7)
    best2, s2 = balance_partitions(tn, 3, init, steps=160, seed=7)
    assert best1 == best2 and s1 == s2                 # step budget + seed -> reproducible
    assert s1 <= par                                   # never worse than the start
    assert compute_solution(tn, best1)[2] == s1


def test_two_networks_back_to_back_do_not_share_a_cache():
    """the local-path cache was a module global keyed by child indices only, so a second network
    was scored with the first one's legs.  Scores must equal a fresh evaluation regardless of call history."""
    tn1 = random_circuit(10, 5, 0.5, 0.5, np.random.default_rng(1))
    tn2 = random_circuit(10, 7, 0.5, 0.5, np.random.default_rng(2))
    n = min(len(tn1.tensors), len(tn2.tensors))
    part = [i % 2 for i in range(n)]
    p1 = part + [0] * (len(tn1.tensors) - n)
    p2 = part + [0] * (len(tn2.tensors) - n)
    alone2 = compute_solution(tn2, p2)[2:]
    compute_solution(tn1, p1)
    b1, s1 = balance_partitions(tn1, 2, p1, steps=40, seed=3)
    assert compute_solution(tn2, p2)[2:] == alone2
    b2, s2 = balance_partitions(tn2, 2, p2, steps=40, seed=3)
    assert compute_solution(tn2, b2)[2] == s2 and compute_solution(tn1, b1)[2] == s1
