"""Greedy path finder restatement vs the reference's five cotengrust KATs
(tnc/src/contractionpath/paths/cotengrust.rs:241-305): exact SSA paths, replace paths, flops, sizes."""
from tnc_b200.contractionpath import path
from tnc_b200.contractionpath.paths import Cotengrust, OptMethod
from tnc_b200.tensornetwork import Tensor


def T(legs, bd):
    return Tensor.new_from_map(legs, bd)


def run(tn):
    opt = Cotengrust(tn, OptMethod.Greedy)
    opt.find_path()
    return opt


def test_greedy_simple():
    bd = {0: 5, 1: 2, 2: 6, 3: 8, 4: 1, 5: 3, 6: 4}
    opt = run(Tensor.new_composite([T([4, 3, 2], bd), T([0, 1, 3, 2], bd), T([4, 5, 6], bd)]))
    assert opt.get_best_flops() == 600.0 and opt.get_best_size() == 538.0
    assert opt.get_best_path() == path((0, 1), (3, 2))
    assert opt.get_best_replace_path() == path((0, 1), (0, 2))


def test_greedy_simple_inner():
    bd = {0: 5, 1: 2, 2: 6, 3: 8, 4: 1, 5: 3, 6: 4}
    opt = run(Tensor.new_composite([T([4, 3, 2], bd), T([4, 3, 2], bd), T([0, 1, 5], bd), T([1, 6], bd)]))
    assert opt.get_best_flops() == 228.0 and opt.get_best_size() == 121.0
    assert opt.get_best_path() == path((0, 1), (2, 3), (4, 5))
    assert opt.get_best_replace_path() == path((0, 1), (2, 3), (0, 2))


def test_greedy_simple_outer():
    bd = {0: 3, 1: 2, 2: 2}
    opt = run(Tensor.new_composite([T([0], bd), T([1], bd), T([2], bd)]))
    assert opt.get_best_flops() == 16.0 and opt.get_best_size() == 19.0
    assert opt.get_best_path() == path((2, 1), (0, 3))
    assert opt.get_best_replace_path() == path((2, 1), (0, 2))


def test_greedy_complex_outer():
    bd = {0: 5, 1: 4}
    opt = run(Tensor.new_composite([T([0], bd), T([0], bd), T([1], bd), T([1], bd)]))
    assert opt.get_best_flops() == 10.0 and opt.get_best_size() == 11.0
    assert opt.get_best_path() == path((0, 1), (2, 3), (5, 4))
    assert opt.get_best_replace_path() == path((0, 1), (2, 3), (2, 0))


def test_greedy_complex():
    bd = {0: 27, 1: 18, 2: 12, 3: 15, 4: 5, 5: 3, 6: 18, 7: 22, 8: 45, 9: 65, 10: 5, 11: 17}
    tn = Tensor.new_composite([T([4, 3, 2], bd), T([0, 1, 3, 2], bd), T([4, 5, 6], bd), T([6, 8, 9], bd),
                               T([10, 8, 9], bd), T([5, 1, 0], bd)])
    opt = run(tn)
    assert opt.get_best_flops() == 529815.0 and opt.get_best_size() == 89478.0
    assert opt.get_best_path() == path((1, 5), (3, 4), (6, 0), (7, 2), (9, 8))
    assert opt.get_best_replace_path() == path((1, 5), (3, 4), (1, 0), (3, 2), (3, 1))


def test_nested_find_path():
    bd = {0: 2, 1: 3, 2: 4, 3: 2, 4: 5}
    a = Tensor.new_composite([T([0, 1], bd), T([1, 2], bd)])
    b = Tensor.new_composite([T([2, 3], bd), T([3, 4], bd), T([4, 0], bd)])
    opt = run(Tensor.new_composite([a, b]))
    p = opt.get_best_replace_path()
    assert set(p.nested) == {0, 1} and p.toplevel == [(0, 1)]
    assert len(p.nested[0].toplevel) == 1 and len(p.nested[1].toplevel) == 2
