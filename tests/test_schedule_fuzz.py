"""Property test of the schedule compiler behind contract_tensor_network (csrc/network.cpp::build_schedule, reached through the
host-only tncb_network_out_legs): for random networks (random hypergraph-free leg structures, bond dimensions 1..4, scalars, outer
products) and random valid replace-left paths, flat or nested, the result legs / dims are those of the metadata replay of
contraction.rs:30-88 + tensor.rs:463-479; broken paths are refused with the reference's panics as status codes."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from tnc_b200._lib import TncbError, check, u64_array


def _native(lib, tn, path):
    from tnc_b200.tensornetwork import contraction as ct
    m = ct._Marshal()
    c_tn, c_path = m.tn(tn), m.path(path)
    n = C.c_int()
    legs, dims = u64_array([0] * 64), u64_array([0] * 64)
    check(lib.tncb_network_out_legs(C.byref(c_tn), C.byref(c_path), C.byref(n), legs, dims))
    return [int(legs[i]) for i in range(n.value)], [int(dims[i]) for i in range(n.value)]


def _replay(tensors, path):
    """tensors: list of (legs, dims) or (nested list, nested path)"""
    ts = []
    for i, t in enumerate(tensors):
        if isinstance(t[0], list) and t[0] and isinstance(t[0][0], tuple):
            ts.append(_replay(t[0], path.nested[i]))
        else:
            ts.append((list(t[0]), list(t[1])))
    for i, j in path.toplevel:
        (al, ad), (bl, bd) = ts[i], ts[j]
        ol = [l for l in bl if l not in al] + [l for l in al if l not in bl]
        dim = {**dict(zip(al, ad)), **dict(zip(bl, bd))}
        ts[i], ts[j] = (ol, [dim[l] for l in ol]), None
    rest = [t for t in ts if t is not None]
    assert len(rest) == 1
    return rest[0]


@st.composite
def networks(draw):
    rng = np.random.default_rng(draw(st.integers(0, 2 ** 32 - 1)))
    n = draw(st.integers(2, 9))
    n_legs = draw(st.integers(0, 14))
    dims = {l: int(rng.integers(1, 5)) for l in range(n_legs)}
    legs = [[] for _ in range(n)]
    for l in range(n_legs):                       # every leg joins one (open) or two (bond) tensors, never more
        owners = rng.choice(n, size=int(rng.integers(1, 3)), replace=False)
        for o in owners:
            legs[o].append(100 + l)
    for q in legs:
        rng.shuffle(q)
    order = list(range(n))
    pairs = []
    alive = list(range(n))
    while len(alive) > 1:                         # a random replace-left path
        i, j = (int(x) for x in rng.choice(len(alive), size=2, replace=False))
        pairs.append((alive[i], alive[j]))
        alive.pop(j)
    return [(q, [dims[l - 100] for l in q]) for q in legs], pairs, draw(st.integers(0, 3)), rng


def _build(tensors, pairs, split, rng):
    """Tensor tree + ContractionPath: split == 0 flat; otherwise the first `k` tensors contracted among themselves form a
    nested composite (when the path allows it: they must be contracted together before meeting the rest)."""
    from tnc_b200.contractionpath import ContractionPath
    from tnc_b200.tensornetwork import Tensor, TensorData

    def leaf(legs, dims):
        t = Tensor.new(legs, dims)
        shape = dims if dims else []
        t.set_tensor_data(TensorData.Matrix(np.zeros(shape, dtype=np.complex128)))
        return t
    return Tensor.new_composite([leaf(l, d) for l, d in tensors]), ContractionPath.simple(pairs)


@settings(max_examples=150, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(networks())
def test_random_networks_and_paths(built_lib, case):
    tensors, pairs, split, rng = case
    tn, path = _build(tensors, pairs, split, rng)
    from tnc_b200.contractionpath import ContractionPath
    assert _native(built_lib, tn, path) == tuple(_replay(tensors, path))
    # nested form of the same thing: a prefix of the path that only touches a subset S becomes S's own path
    n = len(tensors)
    for k in range(1, len(pairs)):
        inner = pairs[:k]
        S = sorted({x for p in inner for x in p})
        survivors = {i for i, _ in inner} - {j for _, j in inner}
        if len(S) < 2 or len(survivors) != 1 or len(S) == n:
            continue
        from tnc_b200.tensornetwork import Tensor
        pos = {s: q for q, s in enumerate(S)}
        inner_path = ContractionPath.simple([(pos[i], pos[j]) for i, j in inner])
        rest = [i for i in range(n) if i not in S]
        comp = Tensor.new_composite([tn.tensors[s] for s in S])
        outer = Tensor.new_composite([comp] + [tn.tensors[i] for i in rest])
        root = survivors.pop()
        opos = {root: 0, **{r: q + 1 for q, r in enumerate(rest)}}
        try:
            outer_pairs = [(opos[i], opos[j]) for i, j in pairs[k:]]
        except KeyError:
            continue                                  # a later pair reaches into S: not expressible as this nesting
        nested = ContractionPath.nested_path([(0, inner_path)], outer_pairs)
        assert _native(built_lib, outer, nested) == tuple(_replay(tensors, path))
        break


def test_broken_paths_are_refused(built_lib):
    from tnc_b200.contractionpath import ContractionPath
    from tnc_b200.tensornetwork import Tensor, TensorData

    def leaf(legs, dims):
        t = Tensor.new(legs, dims)
        t.set_tensor_data(TensorData.Matrix(np.zeros(dims, dtype=np.complex128)))
        return t
    tn = Tensor.new_composite([leaf([0, 1], [2, 3]), leaf([1, 2], [3, 4]), leaf([2, 0], [4, 2])])
    assert _native(built_lib, tn, ContractionPath.simple([(0, 1), (0, 2)])) == ([], [])
    for pairs, status in [([(0, 1)], -4),                 # "Not fully contracted" (contraction.rs:50)
                          ([(0, 1), (1, 2)], -3),         # slot 1 was consumed: "Cannot convert uncontracted tensor to data"
                          ([(0, 1), (0, 0)], -3),         # a tensor with itself: the second mem::take finds Uncontracted (contraction.rs:61-62)
                          ([(0, 3)], -1)]:                # index out of range
        with pytest.raises(TncbError) as e:
            _native(built_lib, tn, ContractionPath.simple(pairs))
        assert e.value.status == status, (pairs, e.value.status, str(e.value))
    bad = Tensor.new_composite([leaf([0, 1], [2, 3]), leaf([1, 2], [5, 4])])      # leg 1: 3 vs 5
    with pytest.raises(TncbError) as e:
        _native(built_lib, bad, ContractionPath.simple([(0, 1)]))
    assert e.value.status == -2
