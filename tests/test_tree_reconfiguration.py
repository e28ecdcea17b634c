"""TreeReconfigure (tnc/src/contractionpath/paths/tree_reconfiguration.rs) and the native subtree reconfiguration /
slicing scores behind it (csrc/reconf.cpp, host only): the reference's two known-answer tests, optimality of the subset DP
against an independent brute force, cost bookkeeping against the Python cost functions, sliced amplitudes on the oracle."""
import itertools
import math

import numpy as np
import pytest

from oracle import tnc_oracle as orc
from tnc_b200.builders import random_circuit
from tnc_b200.contractionpath import ContractionPath, path, ssa_replace_ordering
from tnc_b200.contractionpath.contraction_cost import gpu_time_tensors
from tnc_b200.contractionpath.paths import TreeReconfigure, reconfigure_ssa_path, slice_and_reconfigure
from tnc_b200.contractionpath.paths.cotengrust import _ssa_path_cost, optimize_greedy
from tnc_b200.contractionpath.paths.tree_reconfiguration import leg_scores
from tnc_b200.contractionpath.slicing import SlicedNetwork, path_time
from tnc_b200.tensornetwork import Tensor


def T(legs, bd):
    return Tensor.new_from_map(legs, bd)


def setup_simple():       # tree_reconfiguration.rs:98-106
    bd = {0: 5, 1: 2, 2: 6, 3: 8, 4: 1, 5: 3, 6: 4}
    return Tensor.new_composite([T([4, 3, 2], bd), T([0, 1, 3, 2], bd), T([4, 5, 6], bd)])


def setup_complex():      # tree_reconfiguration.rs:108-131
    bd = {0: 27, 1: 18, 2: 12, 3: 15, 4: 5, 5: 3, 6: 18, 7: 22, 8: 45, 9: 65, 10: 5, 11: 17}
    return Tensor.new_composite([T(l, bd) for l in ([4, 3, 2], [0, 1, 3, 2], [4, 5, 6], [6, 8, 9], [10, 8, 9], [5, 1, 0])])


def test_tree_contract_order_simple():       # tree_reconfiguration.rs:134-143
    opt = TreeReconfigure(setup_simple(), 8, "flops")
    opt.find_path()
    assert opt.best_flops == 600.0 and opt.best_size == 538.0
    assert opt.get_best_path() == path((0, 1), (2, 3))
    assert opt.get_best_replace_path() == path((0, 1), (2, 0))


def test_tree_contract_order_complex():      # tree_reconfiguration.rs:146-164
    opt = TreeReconfigure(setup_complex(), 8, "flops")
    opt.find_path()
    assert opt.best_flops == 332685.0 and opt.best_size == 89478.0
    assert opt.best_path == path((1, 5), (0, 6), (2, 7), (3, 8), (4, 9))
    assert opt.get_best_replace_path() == path((1, 5), (0, 1), (2, 0), (3, 2), (4, 3))


def test_only_flops_is_supported():          # tree_reconfiguration.rs:25-29
    with pytest.raises(ValueError):
        TreeReconfigure(setup_simple(), 8, "size")


def random_network(rng, n, extra):
    """n tensors on a random connected graph (+ `extra` more edges), dims 2..5; every leg joins two tensors."""
    edges = [(i, int(rng.integers(0, i))) for i in range(1, n)]
    while len(edges) < n - 1 + extra:
        a, b = (int(x) for x in rng.integers(0, n, 2))
        if a != b:
            edges.append((a, b))
    inputs = [[] for _ in range(n)]
    size = {}
    for l, (a, b) in enumerate(edges):
        inputs[a].append(l); inputs[b].append(l); size[l] = float(rng.integers(2, 6))
    return inputs, size


def brute_force(inputs, size):
    """cheapest sum of prod dims(legs(a) | legs(b)) over ALL binary trees (independent subset recursion in Python)."""
    n = len(inputs)
    legs = {1 << i: frozenset(t) for i, t in enumerate(inputs)}
    best = {1 << i: 0.0 for i in range(n)}
    for S in range(1, 1 << n):
        if S in best:
            continue
        members = [i for i in range(n) if S >> i & 1]
        low = 1 << members[0]
        c_best = math.inf
        for r in range(0, len(members)):
            for sub in itertools.combinations(members[1:], r):
                S1 = low | sum(1 << i for i in sub)
                S2 = S ^ S1
                if not S2:
                    continue
                c = best[S1] + best[S2] + math.prod(size[l] for l in legs[S1] | legs[S2])
                if c < c_best:
                    c_best, legs[S] = c, legs[S1] ^ legs[S2]
        best[S] = c_best
    return best[(1 << n) - 1]


@pytest.mark.parametrize("seed", range(6))
def test_subset_dp_is_optimal_on_small_networks(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(4, 9))
    inputs, size = random_network(rng, n, int(rng.integers(0, 5)))
    ssa = optimize_greedy(inputs, [], size)
    new, flops, peak, obj = reconfigure_ssa_path(inputs, size, ssa, subtree_size=8, max_sweeps=4)
    assert sorted(x for p in new for x in p) == sorted(set(range(2 * n - 2)))          # every id consumed exactly once
    assert flops == pytest.approx(_ssa_path_cost(inputs, [], size, new, "flops"), rel=1e-12)
    assert peak == pytest.approx(max(_ssa_path_cost(inputs, [], size, new, "size"), max(math.prod(size[l] for l in t) for t in inputs)), rel=1e-12)
    assert flops == pytest.approx(brute_force(inputs, size), rel=1e-12)
    assert obj == pytest.approx(flops, rel=1e-12)


def test_reconfiguration_improves_a_circuit_path_and_keeps_it_valid():
    tn = random_circuit(14, 8, 0.5, 0.5, np.random.default_rng(5))
    inputs = [list(t.legs) for t in tn.tensors]
    size = {l: float(d) for t in tn.tensors for l, d in t.edges()}
    ssa = optimize_greedy(inputs, [], size)
    f0 = _ssa_path_cost(inputs, [], size, ssa, "flops")
    for objective in ("flops", "time"):
        new, flops, peak, obj = reconfigure_ssa_path(inputs, size, ssa, subtree_size=9, max_sweeps=8, objective=objective)
        assert flops == pytest.approx(_ssa_path_cost(inputs, [], size, new, "flops"), rel=1e-12)
        rp = ssa_replace_ordering(ContractionPath.simple(new))
        if objective == "flops":
            assert flops <= f0
        else:   # the C++ time model is the Python one (contraction_cost.gpu_time_tensors)
            meta = [(t.legs, t.bond_dims) for t in tn.tensors]
            assert obj == pytest.approx(path_time(meta, rp, []), rel=1e-9)
            assert obj <= path_time(meta, ssa_replace_ordering(ContractionPath.simple(ssa)), []) * (1 + 1e-12)
        # same amplitude as the greedy path on the oracle
        otn = orc.OTensor(children=[orc.OTensor(list(t.legs), list(t.bond_dims), ("gate",) + tuple(t.tensordata.gate) if t.tensordata.kind == "gate" else np.asarray(t.tensordata.matrix)) for t in tn.tensors])
        a_new = complex(orc.contract_tensor_network(otn, orc.OPath(list(rp.toplevel), {})).data)
        otn = orc.OTensor(children=[orc.OTensor(list(t.legs), list(t.bond_dims), ("gate",) + tuple(t.tensordata.gate) if t.tensordata.kind == "gate" else np.asarray(t.tensordata.matrix)) for t in tn.tensors])
        a_old = complex(orc.contract_tensor_network(otn, orc.OPath(list(ssa_replace_ordering(ContractionPath.simple(ssa)).toplevel), {})).data)
        assert abs(a_new - a_old) <= 1e-12 * abs(a_old)


def test_leg_scores_match_a_recount():
    tn = random_circuit(10, 6, 0.5, 0.5, np.random.default_rng(2))
    inputs = [list(t.legs) for t in tn.tensors]
    size = {l: float(d) for t in tn.tensors for l, d in t.edges()}
    ssa = optimize_greedy(inputs, [], size)
    scores, cost, peak = leg_scores(inputs, size, ssa)
    assert cost == pytest.approx(_ssa_path_cost(inputs, [], size, ssa, "flops"), rel=1e-12)
    for l in list(scores)[::7]:
        cut = [[x for x in t if x != l] for t in inputs]
        assert scores[l][0] == pytest.approx(_ssa_path_cost(cut, [], size, ssa, "flops"), rel=1e-12)
        expect_peak = max(_ssa_path_cost(cut, [], size, ssa, "size"), max(math.prod(size[x] for x in t) for t in cut))
        assert scores[l][1] == pytest.approx(expect_peak, rel=1e-12)


def test_slice_and_reconfigure_bounds_the_width_and_keeps_the_amplitude():
    tn = random_circuit(12, 8, 0.5, 0.5, np.random.default_rng(9))
    inputs = [list(t.legs) for t in tn.tensors]
    size = {l: float(d) for t in tn.tensors for l, d in t.edges()}
    ssa = optimize_greedy(inputs, [], size)
    base = reconfigure_ssa_path(inputs, size, ssa, 8, 8)
    target = base[2] / 8.0
    sliced, new, flops, peak, _ = slice_and_reconfigure(inputs, size, ssa, target, 8, 4)
    assert peak <= target and 1 <= len(sliced) <= 12 and len(set(sliced)) == len(sliced)
    rp = ssa_replace_ordering(ContractionPath.simple(new))

    def leaf(t):
        return orc.OTensor(list(t.legs), list(t.bond_dims), ("gate",) + tuple(t.tensordata.gate) if t.tensordata.kind == "gate" else np.asarray(t.tensordata.matrix))
    flat = complex(orc.contract_tensor_network(orc.OTensor(children=[leaf(t) for t in tn.tensors]),
                                               orc.OPath(list(ssa_replace_ordering(ContractionPath.simple(ssa)).toplevel), {})).data)
    sn = SlicedNetwork(tn, sliced)
    total = 0j
    for asg in sn.assignments:
        s = sn.slice(asg)
        total += complex(orc.contract_tensor_network(orc.OTensor(children=[leaf(t) for t in s.tensors]), orc.OPath(list(rp.toplevel), {})).data)
    assert abs(total - flat) <= 1e-12 * abs(flat)


def test_argument_checks():
    inputs, size = random_network(np.random.default_rng(0), 5, 2)
    ssa = optimize_greedy(inputs, [], size)
    with pytest.raises(Exception):
        reconfigure_ssa_path(inputs, size, ssa, subtree_size=40)
    with pytest.raises(Exception):
        reconfigure_ssa_path(inputs, size, [(0, 1), (0, 2), (5, 3), (6, 4)])        # id 0 used twice
    with pytest.raises(ValueError):
        reconfigure_ssa_path([[0, 1], [0, 2], [0, 3]], {0: 2.0, 1: 2.0, 2: 2.0, 3: 2.0}, [(0, 1), (2, 3)])   # a leg in three tensors
