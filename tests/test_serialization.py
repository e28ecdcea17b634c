"""postcard wire format of Tensor (tnc/src/mpi/serialization.rs): the reference's round-trip test (:84-96), byte-level
checks of the postcard rules (varint, enum index, Vec prefix, f64 LE) on hand-computed vectors, blob padding."""
import struct

import numpy as np

from tnc_b200.dist.serialization import BLOB, deserialize_tensor, serialize, serialize_tensor
from tnc_b200.tensornetwork import Tensor
from tnc_b200.tensornetwork.tensordata import TensorData


def same(a: Tensor, b: Tensor) -> bool:
    if a.legs != b.legs or a.bond_dims != b.bond_dims or len(a.tensors) != len(b.tensors) or a.tensordata.kind != b.tensordata.kind:
        return False
    if a.tensordata.kind == "gate" and a.tensordata.gate != b.tensordata.gate:
        return False
    if a.tensordata.kind == "file" and a.tensordata.file != b.tensordata.file:
        return False
    if a.tensordata.kind == "matrix" and not np.array_equal(np.asarray(a.tensordata.matrix), np.asarray(b.tensordata.matrix)):
        return False
    return all(same(x, y) for x, y in zip(a.tensors, b.tensors))


def test_roundtrip_reference_case():
    """serialization.rs:84-96."""
    bd = {1: 2, 2: 2, 3: 2, 4: 2, 5: 2}
    ta = Tensor.new_composite([Tensor.new_from_map([1, 2, 3], bd), Tensor.new_from_map([2, 3, 4], bd), Tensor.new_from_map([4, 5], bd)])
    blob = serialize_tensor(ta)
    assert len(blob) % BLOB == 0 and len(blob) == BLOB
    assert same(ta, deserialize_tensor(blob))


def test_bytes_follow_postcard_rules():
    leaf = Tensor([3, 300], [2, 70000])
    # tensors: Vec len 0 | legs: len 2, 3, 300 = 0xAC 0x02 | bond_dims: len 2, 2, 70000 = 0xF0 0xA2 0x04 | Uncontracted = 0
    assert serialize(leaf) == bytes([0, 2, 3, 0xAC, 0x02, 2, 2, 0xF0, 0xA2, 0x04, 0])
    g = Tensor([0, 1], [2, 2], tensordata=TensorData.Gate("rx", [0.5], True))
    exp = bytes([0, 2, 0, 1, 2, 2, 2, 2, 2]) + b"rx" + bytes([1]) + struct.pack("<d", 0.5) + bytes([1])
    assert serialize(g) == exp
    f = Tensor([7], [4], tensordata=TensorData.File("a.h5", False))
    assert serialize(f) == bytes([0, 1, 7, 1, 4, 1, 4]) + b"a.h5" + bytes([0])


def test_roundtrip_with_payloads_and_nesting():
    rng = np.random.default_rng(0)
    m = rng.standard_normal((2, 3, 2)) + 1j * rng.standard_normal((2, 3, 2))
    leaves = [Tensor([0, 1, 2], [2, 3, 2], tensordata=TensorData.Matrix(m)),
              Tensor([2, 5], [2, 2], tensordata=TensorData.Gate("u", [0.1, -2.5, 3.0], False)),
              Tensor([5, 9], [2, 2], tensordata=TensorData.File("/data/t.h5", True))]
    tn = Tensor.new_composite([Tensor.new_composite(leaves[:2]), leaves[2], Tensor.new_composite([])])
    tn.legs, tn.bond_dims = [0, 1, 9], [2, 3, 2]
    back = deserialize_tensor(serialize_tensor(tn))
    assert same(tn, back)
    big = Tensor([1], [1 << 40])                      # u64 bond dimension beyond 32 bits
    assert deserialize_tensor(serialize(big)).bond_dims == [1 << 40]
