"""Planning helpers of the partitioned path (metadata only): tree_cut / flatten_nested (contractionpath/tree_partition.py),
the device time model, and the committed bench plan (bench_inputs/c4_partitions.json)."""
import os
import sys

import numpy as np
import pytest

from oracle import tnc_oracle as orc
from tnc_b200.builders import random_circuit
from tnc_b200.contractionpath import validate_path
from tnc_b200.contractionpath.contraction_cost import (contract_path_cost, gpu_fanin_time_tensors, gpu_time_tensors)
from tnc_b200.contractionpath.paths import Cotengrust
from tnc_b200.contractionpath.tree_partition import flatten_nested, tree_cut
from tnc_b200.tensornetwork import Tensor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def greedy(tn):
    opt = Cotengrust(tn); opt.find_path()
    return opt.get_best_replace_path()


def to_oracle(t):
    if t.is_composite():
        return orc.OTensor(children=[to_oracle(c) for c in t.tensors])
    td = t.tensordata
    return orc.OTensor(list(t.legs), list(t.bond_dims), ("gate", td.gate[0], td.gate[1], td.gate[2]) if td.kind == "gate" else np.asarray(td.matrix))


def to_opath(p):
    return orc.OPath(list(p.toplevel), {i: to_opath(q) for i, q in p.nested.items()})


@pytest.mark.parametrize("parts", [2, 3, 5])
def test_tree_cut_is_a_valid_partitioned_contraction(parts):
    tn = random_circuit(9, 5, 0.5, 0.5, np.random.default_rng(parts))
    path = greedy(tn)
    pv, ptn, npath, crit, total = tree_cut(tn, path, parts)
    assert len(ptn.tensors) == parts == len(set(pv)) and sorted(npath.nested) == list(range(parts))
    assert pv[0] == 0 and all(pv.index(k) < pv.index(k + 1) for k in range(parts - 1))      # ids in order of first appearance
    assert len(npath.toplevel) == parts - 1 and validate_path(npath)
    assert sum(len(p.toplevel) for p in npath.nested.values()) + parts - 1 == len(path.toplevel)
    assert total == contract_path_cost(tn.tensors, path, False)[0] == contract_path_cost(ptn.tensors, npath, False)[0]
    assert 0 < crit <= total
    flat_amp = complex(orc.contract_tensor_network(to_oracle(tn), to_opath(path)).data)
    part_amp = complex(orc.contract_tensor_network(to_oracle(ptn), to_opath(npath)).data)
    assert abs(flat_amp - part_amp) <= 1e-12 * abs(flat_amp) + 1e-18
    back = flatten_nested(ptn, npath, pv)
    assert validate_path(back) and len(back.toplevel) == len(path.toplevel)
    again = complex(orc.contract_tensor_network(to_oracle(tn), to_opath(back)).data)
    assert abs(flat_amp - again) <= 1e-12 * abs(flat_amp) + 1e-18


def test_gpu_time_model_is_two_roofed():
    big = gpu_time_tensors(Tensor([0, 1], [4096, 4096]), Tensor([1, 2], [4096, 4096]))         # compute bound, K1' rate
    assert abs(big - (8 * 4096 ** 3 / (160e12 * 4096 / 4696) + 40.0 * 2 * 4096 ** 2 / 5e12 + 5e-6)) < 1e-9
    far = gpu_time_tensors(Tensor([0, 1], [1 << 23, 128]), Tensor([0, 2], [1 << 23, 128]))      # K = 2^23: too long for K1' -> FP64 rate
    assert abs(far - (8 * 128 * 128 * 2.0 ** 23 / (34e12 * 2.0 ** 23 / (2.0 ** 23 + 24)) + 5e-6)) < 1e-9
    thin = gpu_time_tensors(Tensor([0, 1], [1 << 22, 2]), Tensor([1, 2], [2, 2]))                 # bandwidth bound
    assert abs(thin - (16.0 * ((1 << 23) + 4 + (1 << 23)) / 5e12 + 5e-6)) < 1e-12
    a, b = Tensor([0, 1], [1024, 1024]), Tensor([1, 2], [1024, 1024])
    assert gpu_fanin_time_tensors(a, b) == gpu_time_tensors(a, b) + 16.0 * 1024 * 1024 / 6e11 + 30e-6


def test_committed_bench_plan_matches_the_network():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import plan_partitions as pp
    tn = pp.build_network()
    for n in (2, 4, 8):
        got = pp.load(tn, n)
        assert got is not None, "bench_inputs/c4_partitions.json is stale: run tools/plan_partitions.py"
        ptn, path, facts = got
        assert len(ptn.tensors) == n and sorted(path.nested) == list(range(n)) and len(path.toplevel) == n - 1
        assert sum(len(c.tensors) for c in ptn.tensors) == len(tn.tensors) == 489 and validate_path(path)
        assert facts["partition_sizes"] == [len(c.tensors) for c in ptn.tensors]


@pytest.mark.parametrize("name", ["sycamore53_d12.json", "sycamore53_d12_alt.json"])
def test_committed_config5_paths_match_the_network(name):
    """bench.py's config5 object replays bench_inputs/sycamore53_d12.json on sycamore_circuit(53, 12), seed 1: the committed
    replace-left path must be valid for exactly that network (1053 tensors, every slot consumed once, scalar at the end), its
    sliced legs must be summed legs, and the recorded flop count / peak size must be what the path gives."""
    import json
    from tnc_b200.builders import sycamore_circuit
    from tnc_b200.contractionpath import ContractionPath
    from tnc_b200.contractionpath.slicing import path_cost
    p = os.path.join(ROOT, "bench_inputs", name)
    if not os.path.exists(p):
        pytest.skip(name + " not committed")
    d = json.load(open(p))
    w = d["network"].split()
    assert (w[0], w[1], w[3], w[5]) == ("sycamore", "53q", "12", "1")
    tn = sycamore_circuit(53, 12, np.random.default_rng(1)).into_amplitude_network("0" * 53)[0]
    assert len(tn.tensors) == 1053
    path = ContractionPath.simple([tuple(x) for x in d["toplevel"]])
    assert len(path.toplevel) == 1052 and validate_path(path)
    alive = [True] * 1053
    for i, j in path.toplevel:
        assert alive[i] and alive[j] and i != j
        alive[j] = False
    assert sum(alive) == 1
    meta = [(t.legs, t.bond_dims) for t in tn.tensors]
    count = {}
    for legs, _ in meta:
        for l in legs:
            count[l] = count.get(l, 0) + 1
    assert all(count.get(l) == 2 for l in d["sliced_legs"]), "sliced legs must be bonds between two tensors"
    assert 2 ** len(d["sliced_legs"]) == d["n_slices"]
    flops, peak, peak_legs = path_cost(meta, path, d["sliced_legs"])
    assert peak <= 2.0 ** 30 and peak == d["peak_elements"]
    assert abs(flops / 8.0 - d["flops_mnk_per_slice"]) <= 1e-9 * d["flops_mnk_per_slice"]
    # the whole network contracts to a scalar along the path
    ts = [dict(zip(l, dd)) for l, dd in meta]
    for i, j in path.toplevel:
        a, b = ts[i], ts[j]
        ts[i] = {**{l: x for l, x in b.items() if l not in a}, **{l: x for l, x in a.items() if l not in b}}
        ts[j] = None
    assert [t for t in ts if t is not None] == [{}]
