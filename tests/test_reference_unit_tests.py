"""The reference's own unit tests of the host-side types on the path, replayed against the Python mirror (no GPU):
tensor.rs:589-728 (Tensor), mpi/mpi_types.rs:86-131 (RankTensorMapping), builders/random_circuit.rs:281-412 (the two
structure KATs of the observable circuits, independent of the random stream), builders/circuit_builder.rs:357-369
(permutation_between)."""
import numpy as np
import pytest

from tnc_b200.tensornetwork import Tensor, TensorData


# ---------------------------------------------------------------- tensor.rs:589-728
def same(a: Tensor, b: Tensor) -> bool:
    return a.legs == b.legs and a.bond_dims == b.bond_dims and a.tensordata.kind == b.tensordata.kind and \
        len(a.tensors) == len(b.tensors) and all(same(x, y) for x, y in zip(a.tensors, b.tensors))


def test_empty_tensor():
    t = Tensor()
    assert t.tensors == [] and t.legs == [] and t.bond_dims == [] and t.is_empty()


def test_new():
    t = Tensor.new([2, 4, 5], [4, 2, 6])
    assert t.legs == [2, 4, 5] and t.bond_dims == [4, 2, 6] and t.tensor_data().kind == "uncontracted"


def test_new_from_map():
    t = Tensor.new_from_map([2, 4, 5], {1: 1, 2: 4, 3: 7, 4: 2, 5: 6})
    assert t.legs == [2, 4, 5] and t.bond_dims == [4, 2, 6] and t.tensor_data().kind == "uncontracted"


def test_new_from_const():
    t = Tensor.new_from_const([9, 2, 5, 1], 3)
    assert t.legs == [9, 2, 5, 1] and t.bond_dims == [3, 3, 3, 3] and t.tensor_data().kind == "uncontracted"


def test_external_tensor():
    bd = {2: 2, 3: 4, 4: 6, 5: 8, 6: 10, 7: 12, 8: 14, 9: 16}
    t12 = Tensor.new_composite([Tensor.new_from_map([2, 3, 4], bd), Tensor.new_from_map([2, 3, 5], bd)])
    t34 = Tensor.new_composite([Tensor.new_from_map([6, 7, 8], bd), Tensor.new_from_map([6, 8, 9], bd)])
    ext = Tensor.new_composite([t12, t34]).external_tensor()
    assert same(ext, Tensor.new_from_map([4, 5, 7, 9], bd))


BD = {2: 17, 3: 1, 4: 11, 8: 3, 9: 20, 7: 7, 10: 14}


def test_push_tensor():
    t = Tensor()
    t.push_tensor(Tensor.new_from_map([8, 4, 9], BD))
    assert len(t.tensors) == 1 and same(t.tensors[0], Tensor.new_from_map([8, 4, 9], BD))
    t.push_tensor(Tensor.new_from_map([7, 10, 2], BD))
    assert [x.legs for x in t.tensors] == [[8, 4, 9], [7, 10, 2]] and [x.bond_dims for x in t.tensors] == [[3, 11, 20], [7, 14, 17]]
    assert t.tensor_data().kind == "uncontracted" and t.legs == []


def test_push_tensor_to_leaf():
    leaf = Tensor.new_from_map([4, 3, 2], BD)
    with pytest.raises(AssertionError, match="Cannot push tensors into a leaf tensor"):
        leaf.push_tensor(Tensor.new_from_map([8, 4, 9], BD))


def test_push_tensors():
    t = Tensor()
    t.push_tensors([Tensor.new_from_map([4, 3, 2], BD), Tensor.new_from_map([8, 4, 9], BD), Tensor.new_from_map([7, 10, 2], BD)])
    assert t.tensor_data().kind == "uncontracted"
    assert [x.legs for x in t.tensors] == [[4, 3, 2], [8, 4, 9], [7, 10, 2]]
    assert [x.bond_dims for x in t.tensors] == [[11, 1, 17], [3, 11, 20], [7, 14, 17]]


def test_push_tensors_to_leaf():
    leaf = Tensor.new_from_map([4, 3, 2], BD)
    with pytest.raises(AssertionError, match="Cannot push tensors into a leaf tensor"):
        leaf.push_tensors([Tensor.new_from_map([8, 4, 9], BD), Tensor.new_from_map([7, 10, 2], BD)])


# ---------------------------------------------------------------- mpi/mpi_types.rs:86-131
def test_tensor_mapping():
    from tnc_b200.dist import RankTensorMapping
    m = RankTensorMapping()
    assert m.tensor(2) is None and m.tensor(3) is None
    m.insert(2, 4)
    assert m.rank(4) == 2 and m.tensor(2) == 4 and m.tensor(3) is None
    m.insert(3, 0)
    assert m.rank(4) == 2 and m.tensor(2) == 4 and m.rank(0) == 3 and m.tensor(3) == 0
    assert len(m) == 2 and not m.is_empty() and list(m) == [(2, 4), (3, 0)]


def test_tensor_mapping_insert_rank_twice():
    from tnc_b200.dist import RankTensorMapping
    m = RankTensorMapping()
    m.insert(2, 4); m.insert(3, 0)
    with pytest.raises(AssertionError, match="Rank 2 is already associated with a tensor"):
        m.insert(2, 5)


def test_tensor_mapping_insert_tensor_twice():
    from tnc_b200.dist import RankTensorMapping
    m = RankTensorMapping()
    m.insert(2, 4); m.insert(3, 5)
    with pytest.raises(AssertionError, match="Tensor 4 is already associated with a rank"):
        m.insert(4, 4)


# ---------------------------------------------------------------- builders/circuit_builder.rs:357-369
@pytest.mark.parametrize("given,target", [([1, 2, 3, 4], [1, 2, 3, 4]), ([1, 2, 3, 4], [4, 3, 2, 1]), ([4, 3, 2, 1], [1, 2, 3, 4]),
                                          ([4, 1, 3, 2], [2, 4, 3, 1]), ([5, 1, 4, 3, 2, 6], [1, 6, 3, 5, 2, 4])])
def test_permutation_between(given, target):
    from tnc_b200.builders import Permutor
    perm = Permutor.permutation_between(given, target)
    assert [given[p] for p in perm] == target
    # ... and it is the axis order of the transpose: a tensor with legs `given` becomes one with legs `target`
    shape = [2 + (l % 3) for l in given]
    a = np.arange(int(np.prod(shape))).reshape(shape)
    assert list(np.transpose(a, perm).shape) == [2 + (l % 3) for l in target]


# ---------------------------------------------------------------- builders/random_circuit.rs:281-412
def test_random_circuit_with_observable():
    from tnc_b200.builders import random_circuit_with_observable
    c = random_circuit_with_observable(4, 3, 1.0, 1.0, 1.0, np.random.default_rng(), "line", 4)   # independent of the rng
    ref = [[0, 1], [2, 3], [4, 5], [6, 7], [8, 9, 0, 2], [1, 3, 10, 11], [12, 13, 9, 4], [11, 5, 14, 15], [16, 17, 13, 6],
           [15, 7, 18, 19], [20, 8], [10, 21], [22, 12], [14, 23], [24, 16], [18, 25], [26, 17], [19, 27], [28, 29, 20, 22],
           [21, 23, 30, 31], [32, 33, 29, 24], [31, 25, 34, 35], [36, 37, 33, 26], [35, 27, 38, 39], [40, 28], [30, 41], [42, 32],
           [34, 43], [44, 36], [38, 45], [46, 37], [39, 47], [40], [41], [42], [43], [44], [45], [46], [47]]
    assert len(c.tensors) == 40
    assert [t.legs for t in c.tensors] == ref and all(t.bond_dims == [2] * len(t.legs) for t in c.tensors)


def test_random_circuit_with_set_observable():
    from tnc_b200.builders import random_circuit_with_set_observable
    c = random_circuit_with_set_observable(4, 3, 1.0, 1.0, [2], np.random.default_rng(), "line", 4)
    ref = [[0, 1], [3, 4, 2, 0], [2, 1, 5, 6], [8, 9, 4, 7], [6, 7, 10, 11], [12, 3], [5, 13], [14, 8], [10, 15], [16, 9], [11, 17],
           [19, 20, 18, 12], [18, 13, 21, 22], [23, 24, 20, 14], [22, 15, 25, 26], [27, 28, 24, 16], [26, 17, 29, 30], [31, 19],
           [21, 32], [33, 23], [25, 34], [35, 27], [29, 36], [37, 28], [30, 38], [31], [32], [33], [34], [35], [36], [37], [38]]
    assert len(c.tensors) == 33
    assert [t.legs for t in c.tensors] == ref and all(t.bond_dims == [2] * len(t.legs) for t in c.tensors)
    # payloads: one Pauli observable, fsim + adjoint fsim pairs, single-qubit gates with their mirror, closing product states
    kinds = [t.tensordata.gate[0] if t.tensordata.kind == "gate" else t.tensordata.kind for t in c.tensors]
    assert kinds[0] in ("x", "y", "z") and kinds[1:5] == ["fsim"] * 4 and kinds[-8:] == ["matrix"] * 8
    assert c.tensors[1].tensordata.gate[2] is False and c.tensors[2].tensordata.gate[2] is True
    np.testing.assert_array_equal(c.tensors[-2].tensordata.matrix, c.tensors[-1].tensordata.matrix)


def test_random_sparse_tensor_data():
    from tnc_b200.builders import random_sparse_tensor_data_with_rng
    rng = np.random.default_rng(3)
    full = random_sparse_tensor_data_with_rng([2], 1.0, rng).matrix
    assert full.shape == (2,) and full.dtype == np.complex128
    half = random_sparse_tensor_data_with_rng([5, 4, 3], None, rng).matrix          # 30 draws into 60 cells
    assert half.shape == (5, 4, 3) and 1 <= np.count_nonzero(half) <= 30
    assert (half.real >= 0).all() and (half.real < 1).all() and (half.imag >= 0).all() and (half.imag < 1).all()
