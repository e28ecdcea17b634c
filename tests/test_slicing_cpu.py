"""Slicing metadata + sliced-network construction, checked against the oracle on the CPU
(the GPU execution of slices is covered by tests/test_gpu_networks.py::test_sliced_equals_flat)."""
import numpy as np

from oracle import tnc_oracle as orc
from tnc_b200.builders import random_circuit
from tnc_b200.contractionpath.paths import Cotengrust
from tnc_b200.contractionpath.slicing import SlicedNetwork, find_slices, path_cost
from tnc_b200.dist.communication import Communication, contracted_legs, fanin_schedule
from tnc_b200.tensornetwork.partitioning import find_partitioning, partition_tensor_network


def _to_oracle(t):
    if t.is_composite():
        return orc.OTensor(children=[_to_oracle(c) for c in t.tensors])
    td = t.tensordata
    d = ("gate", td.gate[0], td.gate[1], td.gate[2]) if td.kind == "gate" else (np.asarray(td.matrix) if td.kind == "matrix" else None)
    return orc.OTensor(list(t.legs), list(t.bond_dims), d)


def _to_opath(p):
    return orc.OPath(list(p.toplevel), {i: _to_opath(q) for i, q in p.nested.items()})


def test_sum_over_slices_equals_unsliced_on_the_oracle():
    tn = random_circuit(10, 6, 0.5, 0.5, np.random.default_rng(22))
    opt = Cotengrust(tn); opt.find_path(); p = opt.get_best_replace_path()
    ref = complex(orc.contract_tensor_network(_to_oracle(tn), _to_opath(p)).data)
    legs = find_slices(tn, p, min_slices=4)
    sn = SlicedNetwork(tn, legs)
    assert len(sn.assignments) >= 4 and len(set(legs)) == len(legs)
    tot = sum(complex(orc.contract_tensor_network(_to_oracle(sn.slice(a)), _to_opath(p)).data) for a in sn.assignments)
    assert abs(tot - ref) <= 1e-13 * max(1.0, abs(ref))
    for a in sn.assignments[:2]:      # sliced legs are gone from every leaf
        assert all(not (set(t.legs) & set(legs)) for t in sn.slice(a).tensors)


def test_slicing_reduces_peak_and_keeps_work_bounded():
    tn = random_circuit(20, 8, 0.5, 0.5, np.random.default_rng(4))
    opt = Cotengrust(tn); opt.find_path(); p = opt.get_best_replace_path()
    meta = [(t.legs, t.bond_dims) for t in tn.tensors]
    f0, pk0, _ = path_cost(meta, p)
    legs = find_slices(tn, p, min_slices=8)
    f, pk, _ = path_cost(meta, p, legs)
    assert pk <= pk0 / 2 and f * 2 ** len(legs) <= 4 * f0
    legs2 = find_slices(tn, p, min_slices=1, max_peak_elements=pk0 / 16)
    assert path_cost(meta, p, legs2)[1] <= pk0 / 16


def test_fanin_metadata_matches_oracle_result_legs():
    """The receiver of a fan-in message learns the sender's leg ORDER from a metadata replay of the
    sender's local path (only raw data travels); it must equal what the contraction really produces."""
    tn = random_circuit(12, 6, 0.5, 0.5, np.random.default_rng(5))
    ptn = partition_tensor_network(tn, find_partitioning(tn, 3, seed=1))
    opt = Cotengrust(ptn); opt.find_path(); p = opt.get_best_replace_path()
    ext = {}
    for k in sorted(p.nested):
        res = orc.contract_tensor_network(_to_oracle(ptn.tensor(k)), _to_opath(p.nested[k]))
        l, d = contracted_legs(ptn.tensor(k), p.nested[k])
        assert (l, d) == (res.legs, res.dims)
        ext[k] = (l, d)
    comm = Communication({0: 0, 1: 1, 2: 2}, ext)
    ev = fanin_schedule(comm, p.toplevel)
    assert len(ev) == 2 and ev[-1]["out_legs"] == []
    full = orc.contract_tensor_network(_to_oracle(ptn), _to_opath(p))
    assert full.legs == []
