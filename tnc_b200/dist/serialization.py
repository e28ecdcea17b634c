"""Wire format of the reference's MPI path: postcard(serde(Tensor)) padded to 192-byte blobs
(tnc/src/mpi/serialization.rs:43-79, mpi_types.rs:73-83).

The GPU fan-in never uses this (boundary tensors travel as raw complex128 over NCCL, legs/dims are derived on every
rank); it exists for interoperability with host-side TNC ranks and as the on-disk form of a network description.

postcard 1.x wire rules (published spec, restated): unsigned integers are LEB128 varints, f64 is 8 bytes
little-endian, bool one byte, a string / Vec is varint(len) followed by the elements, a struct is its fields in
declaration order, an enum is varint(variant index) followed by the payload, tuples are their elements in order.
`Tensor` (tensor.rs:21-37) = { tensors: Vec<Tensor>, legs: Vec<usize>, bond_dims: Vec<u64>, tensordata: TensorData },
`TensorData` (tensordata.rs:14-25) = Uncontracted | File((PathBuf, bool)) | Gate((String, Vec<f64>, bool)) | Matrix(tetra::Tensor).

Parity note: the serde layout of `tetra::Tensor` lives in the un-vendored crate tetra 0.1.0 (Cargo.lock:3264-3266) and
the reference holds no byte-level golden vector (only the round-trip test serialization.rs:84-96), so the Matrix payload
is *unpinned*: it is written as { shape: Vec<usize>, data: Vec<(f64, f64)> } (row-major), the natural derive of a
shape + flat-data struct.  Everything else follows the derives in /root/reference."""
from __future__ import annotations

import struct
from typing import List, Tuple

import numpy as np

from ..tensornetwork.tensor import Tensor
from ..tensornetwork.tensordata import TensorData

BLOB = 192  # size_of::<MessageBinaryBlob>() (mpi_types.rs:81-83)


def _varint(n: int) -> bytes:
    assert n >= 0
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    shift = val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _enc_tensor(t: Tensor, out: List[bytes]) -> None:
    out.append(_varint(len(t.tensors)))
    for c in t.tensors:
        _enc_tensor(c, out)
    out.append(_varint(len(t.legs)))
    out.extend(_varint(l) for l in t.legs)
    out.append(_varint(len(t.bond_dims)))
    out.extend(_varint(d) for d in t.bond_dims)
    td = t.tensordata
    if td.kind == "uncontracted":
        out.append(_varint(0))
    elif td.kind == "file":
        path, adj = td.file
        p = str(path).encode()
        out += [_varint(1), _varint(len(p)), p, b"\x01" if adj else b"\x00"]
    elif td.kind == "gate":
        name, angles, adj = td.gate
        nb = name.encode()
        out += [_varint(2), _varint(len(nb)), nb, _varint(len(angles))]
        out.extend(struct.pack("<d", float(a)) for a in angles)
        out.append(b"\x01" if adj else b"\x00")
    elif td.kind == "matrix":
        m = td.matrix
        arr = np.ascontiguousarray(m if isinstance(m, np.ndarray) else m.to_numpy(), dtype=np.complex128)
        out += [_varint(3), _varint(arr.ndim)]
        out.extend(_varint(s) for s in arr.shape)
        out += [_varint(arr.size), arr.tobytes()]          # (re, im) little-endian f64 pairs, row-major
    else:
        raise ValueError(f"unknown TensorData kind {td.kind!r}")


def serialize(t: Tensor) -> bytes:
    """postcard::to_stdvec(&tensor) (serialization.rs:4-9)."""
    parts: List[bytes] = []
    _enc_tensor(t, parts)
    return b"".join(parts)


def serialize_tensor(t: Tensor) -> bytes:
    """serialize_tensor (serialization.rs:43-67): the postcard bytes zero-padded to whole 192-byte blobs."""
    raw = serialize(t)
    return raw + b"\x00" * (-len(raw) % BLOB)


def _dec_tensor(buf: memoryview, pos: int) -> Tuple[Tensor, int]:
    n, pos = _read_varint(buf, pos)
    children = []
    for _ in range(n):
        c, pos = _dec_tensor(buf, pos)
        children.append(c)
    n, pos = _read_varint(buf, pos)
    legs = []
    for _ in range(n):
        v, pos = _read_varint(buf, pos)
        legs.append(v)
    n, pos = _read_varint(buf, pos)
    dims = []
    for _ in range(n):
        v, pos = _read_varint(buf, pos)
        dims.append(v)
    kind, pos = _read_varint(buf, pos)
    if kind == 0:
        td = TensorData.Uncontracted()
    elif kind == 1:
        n, pos = _read_varint(buf, pos)
        path = bytes(buf[pos:pos + n]).decode(); pos += n
        td = TensorData.File(path, bool(buf[pos])); pos += 1
    elif kind == 2:
        n, pos = _read_varint(buf, pos)
        name = bytes(buf[pos:pos + n]).decode(); pos += n
        n, pos = _read_varint(buf, pos)
        angles = list(struct.unpack_from(f"<{n}d", buf, pos)); pos += 8 * n
        td = TensorData.Gate(name, angles, bool(buf[pos])); pos += 1
    elif kind == 3:
        nd, pos = _read_varint(buf, pos)
        shape = []
        for _ in range(nd):
            v, pos = _read_varint(buf, pos)
            shape.append(v)
        cnt, pos = _read_varint(buf, pos)
        arr = np.frombuffer(buf, dtype=np.complex128, count=cnt, offset=pos).reshape(shape).copy(); pos += 16 * cnt
        td = TensorData.Matrix(arr)
    else:
        raise ValueError(f"bad TensorData variant {kind}")
    t = Tensor(legs, dims, tensors=children, tensordata=td)
    return t, pos


def deserialize(data: bytes) -> Tensor:
    """postcard::from_bytes (serialization.rs:33-39); trailing blob padding is ignored like postcard ignores the rest."""
    t, _ = _dec_tensor(memoryview(data), 0)
    return t


deserialize_tensor = deserialize
