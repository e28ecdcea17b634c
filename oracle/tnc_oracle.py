"""CPU ORACLE -- test infrastructure only, never a product path.

A numpy restatement of the reference's pairwise-contraction hot path
(qc-tum/TNC @ 5dd62b3).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
module; ``tnc_b200`` itself must never do so (it fails loudly when the CUDA
library is missing instead of falling back to this code).

Parity pinning: the arithmetic of the path lives in the un-vendored crate
``tetra@2c2a23af`` (HPTT + faer/MKL, Cargo.lock:3264-3266), so the oracle
restates the *published* algorithm (transpose-transpose-GEMM of row-major
complex128 tensors) and is pinned by the reference's own golden vectors:
``tnc/src/tensornetwork/contraction_test_data.json`` (tests/golden/
contraction_kat.json, 3 KATs at 1e-14), the outer-product KAT
(contraction.rs:227-264), the statevector / partial-amplitude KATs
(io/qasm/qasm_importer.rs:171-298), the closed-form amplitude / expectation
KATs (builders/circuit_builder.rs:372-427) and the adjoint identity
(gates.rs:657-680).  See tests/test_oracle_golden.py.

Every function cites the reference file:line it follows (paths relative to
/root/reference/).
"""
from __future__ import annotations

import cmath
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

FRAC_1_SQRT_2 = 0.70710678118654752440


# ----------------------------------------------------------------------------
# gates  (tnc/src/gates.rs:147-627)
# ----------------------------------------------------------------------------
def _exp_i(x: float) -> complex:
    # Complex64::new(0, x).exp() == (cos x, sin x)
    return complex(math.cos(x), math.sin(x))


def gate_matrix(name: str, angles: Sequence[float]) -> np.ndarray:
    """`Gate::compute` for the 18 registered gates (gates.rs:147-627).
    Row-major 2x2 or 2x2x2x2 complex128."""
    z, o, i = 0j, 1 + 0j, 1j
    a = list(angles)

    def need(n):
        if len(a) != n:
            raise ValueError(f"Expected {n} angles, but got {len(a)}.")  # gates.rs:103-107

    if name == "x":
        need(0); d = [z, o, o, z]
    elif name == "y":
        need(0); d = [z, -i, i, z]
    elif name == "z":
        need(0); d = [o, z, z, -o]
    elif name == "h":
        need(0); h = complex(FRAC_1_SQRT_2, 0); d = [h, h, h, -h]
    elif name == "t":
        need(0); d = [o, z, z, complex(FRAC_1_SQRT_2, FRAC_1_SQRT_2)]
    elif name == "u":
        need(3); th, ph, la = a
        s, c = math.sin(th / 2), math.cos(th / 2)
        d = [complex(c, 0), -_exp_i(la) * s, _exp_i(ph) * s, _exp_i(ph + la) * c]
    elif name == "sx":
        need(0); p, q = complex(.5, .5), complex(.5, -.5); d = [p, q, q, p]
    elif name == "sy":
        need(0); p, q = complex(.5, .5), complex(-.5, -.5); d = [p, q, p, p]
    elif name == "sz":
        need(0); d = [o, z, z, i]
    elif name == "rx":
        need(1); s, c = math.sin(a[0] / 2), math.cos(a[0] / 2)
        d = [o * c, -i * s, -i * s, o * c]
    elif name == "ry":
        need(1); s, c = math.sin(a[0] / 2), math.cos(a[0] / 2)
        d = [o * c, -o * s, o * s, o * c]
    elif name == "rz":
        need(1); d = [cmath.exp(-i * a[0] / 2), z, z, cmath.exp(i * a[0] / 2)]
    elif name == "cx":
        need(0); d = [o, z, z, z, z, o, z, z, z, z, z, o, z, z, o, z]
    elif name == "cz":
        need(0); d = [o, z, z, z, z, o, z, z, z, z, o, z, z, z, z, -o]
    elif name == "swap":
        need(0); d = [o, z, z, z, z, z, o, z, z, o, z, z, z, z, z, o]
    elif name == "cp":
        need(1); e = _exp_i(a[0]); d = [o, z, z, z, z, o, z, z, z, z, o, z, z, z, z, e]
    elif name == "iswap":
        need(0); d = [o, z, z, z, z, z, i, z, z, i, z, z, z, z, z, o]
    elif name == "fsim":
        need(2); th, ph = a
        p = complex(math.cos(th), 0); q = complex(0, -math.sin(th)); c = cmath.exp(complex(0, -ph))
        d = [o, z, z, z, z, p, q, z, z, q, p, z, z, z, z, c]
    else:
        raise KeyError(f"Gate '{name}' not found.")  # gates.rs:54
    arr = np.array(d, dtype=np.complex128)
    return arr.reshape((2, 2) if arr.size == 4 else (2, 2, 2, 2))


def matrix_adjoint(data: np.ndarray) -> np.ndarray:
    """gates.rs:82-99: swap the first half of the dims with the second, conj."""
    if data.ndim > 0:
        assert data.ndim & (data.ndim - 1) == 0
        half = data.ndim // 2
        perm = list(range(half, data.ndim)) + list(range(half))
        data = np.transpose(data, perm)
    return np.ascontiguousarray(np.conj(data))


def load_data_hdf5(path) -> np.ndarray:
    """load_data (io/hdf5.rs:37-43, 90-103): the first member (name order) of /tensors.  A reader of its own -- it shares no
    code with csrc/hdf5io.cpp -- for the "earliest"-format files libhdf5 writes by default (superblock 0, symbol-table
    groups, version-1 object headers, contiguous little-endian {re, im} doubles); enough for what the tests store."""
    b = open(path, "rb").read()

    def u(off, n=8):
        return int.from_bytes(b[off:off + n], "little")

    def messages(addr):
        size, p, out = u(addr + 8, 4), addr + 16, []
        while p < addr + 16 + size:
            t, s = u(p, 2), u(p + 2, 2)
            out.append((t, b[p + 8:p + 8 + s]))
            p += 8 + s
        return out

    def members(header):
        st = [d for t, d in messages(header) if t == 0x11][0]
        btree, heap = int.from_bytes(st[:8], "little"), int.from_bytes(st[8:16], "little")
        seg = u(heap + 24)
        out = []

        def walk(node):
            level, used = b[node + 5], u(node + 6, 2)
            for i in range(used):
                child = u(node + 24 + 16 * i + 8)
                if level:
                    walk(child)
                    continue
                for q in range(u(child + 6, 2)):
                    e = child + 8 + 40 * q
                    off = seg + u(e)
                    out.append((b[off:b.index(b"\0", off)], u(e + 8)))
        walk(btree)
        return sorted(out)

    assert b[:8] == b"\x89HDF\r\n\x1a\n" and b[8] == 0, "oracle reader: superblock 0 files only"
    tensors = dict(members(u(64)))[b"tensors"]
    _, header = members(tensors)[0]
    shape, addr = None, None
    for t, d in messages(header):
        if t == 0x01:
            shape = [int.from_bytes(d[8 + 8 * i:16 + 8 * i], "little") for i in range(d[1])]
        elif t == 0x08:
            assert d[0] == 3 and d[1] == 1, "oracle reader: contiguous layout only"
            addr = int.from_bytes(d[2:10], "little")
    n = int(np.prod(shape, dtype=np.int64)) if shape else 1
    if addr == 0xFFFFFFFFFFFFFFFF:                       # declared but never written (the "-1" output tensor): fill value 0
        return np.zeros(shape, dtype=np.complex128)
    return np.frombuffer(b, dtype=np.complex128, count=n, offset=addr).reshape(shape).copy()


def load_gate(name: str, angles: Sequence[float], adjoint: bool = False) -> np.ndarray:
    """gates.rs:50-66.  The reference's specialised adjoints equal the generic
    conj-transpose (pinned by gates.rs:657-680), so the generic rule is used."""
    m = gate_matrix(name, angles)
    return matrix_adjoint(m) if adjoint else m


# ----------------------------------------------------------------------------
# data model (tnc/src/tensornetwork/tensor.rs:21-37, tensordata.rs:15-26)
# ----------------------------------------------------------------------------
@dataclass
class OTensor:
    """Leaf: legs/dims + payload.  Composite: children only.
    payload: None (Uncontracted) | np.ndarray (Matrix) | ("gate", name, angles, adjoint) | ("file", path, adjoint)."""
    legs: List[int] = field(default_factory=list)
    dims: List[int] = field(default_factory=list)
    data: object = None
    children: List["OTensor"] = field(default_factory=list)

    @property
    def is_composite(self) -> bool:
        return len(self.children) > 0

    def materialise(self) -> np.ndarray:
        """TensorData::into_data (tensordata.rs:40-59)."""
        if self.data is None:
            raise RuntimeError("Cannot convert uncontracted tensor to data")  # tensordata.rs:42
        if isinstance(self.data, tuple) and self.data[0] == "file":    # TensorData::File (tensordata.rs:43-49)
            data = load_data_hdf5(self.data[1])
            return matrix_adjoint(data) if self.data[2] else data
        if isinstance(self.data, tuple):
            _, name, angles, adj = self.data
            return load_gate(name, angles, adj)
        return np.ascontiguousarray(np.asarray(self.data, dtype=np.complex128)).reshape(self.dims)


@dataclass
class OPath:
    """ContractionPath (contractionpath.rs:29-35), replace-left pairs."""
    toplevel: List[Tuple[int, int]] = field(default_factory=list)
    nested: Dict[int, "OPath"] = field(default_factory=dict)


def sym_diff_legs(x_legs, x_dims, y_legs, y_dims):
    """Tensor::symmetric_difference (tensor.rs:463-479): x's survivors then y's."""
    legs, dims = [], []
    for l, d in zip(x_legs, x_dims):
        if l not in y_legs:
            legs.append(l); dims.append(d)
    for l, d in zip(y_legs, y_dims):
        if l not in x_legs:
            legs.append(l); dims.append(d)
    return legs, dims


def external_legs(t: OTensor):
    """Tensor::external_tensor (tensor.rs:482-498)."""
    if not t.is_composite:
        return list(t.legs), list(t.dims)
    legs, dims = [], []
    for c in t.children:
        cl, cd = external_legs(c)
        legs, dims = sym_diff_legs(legs, dims, cl, cd)
    return legs, dims


# ----------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------
def contract_pair(a_legs, a: np.ndarray, b_legs, b: np.ndarray, backend: str = "numpy"):
    """tetra::contract as called at contraction.rs:78-84: C[(b\\a)++(a\\b)] =
    sum over shared legs, row-major.  TTGT: permute both operands so the shared
    legs are adjacent, reshape to matrices, one ZGEMM (book/src/theory.md:32-44)."""
    a_legs, b_legs = list(a_legs), list(b_legs)
    shared = [l for l in a_legs if l in b_legs]
    bf = [i for i, l in enumerate(b_legs) if l not in a_legs]
    af = [i for i, l in enumerate(a_legs) if l not in b_legs]
    bk = [b_legs.index(l) for l in shared]
    ak = [a_legs.index(l) for l in shared]
    out_legs = [b_legs[i] for i in bf] + [a_legs[i] for i in af]
    out_shape = [b.shape[i] for i in bf] + [a.shape[i] for i in af]
    N = int(np.prod([b.shape[i] for i in bf], dtype=np.int64)) if bf else 1
    M = int(np.prod([a.shape[i] for i in af], dtype=np.int64)) if af else 1
    K = int(np.prod([a.shape[i] for i in ak], dtype=np.int64)) if ak else 1
    if backend == "torch":
        import torch
        bt = b.permute(bf + bk).contiguous().reshape(N, K)
        at = a.permute(ak + af).contiguous().reshape(K, M)
        return out_legs, torch.matmul(bt, at).reshape(out_shape)
    bm = np.ascontiguousarray(np.transpose(b, bf + bk)).reshape(N, K)
    am = np.ascontiguousarray(np.transpose(a, ak + af)).reshape(K, M)
    return out_legs, (bm @ am).reshape(out_shape)


def contract_tensor_network(tn: OTensor, path: OPath, backend: str = "numpy",
                            stats: Optional[dict] = None) -> OTensor:
    """contract_tensor_network (contraction.rs:30-52) + contract_tensors (:60-88)."""
    tensors = list(tn.children)
    for idx in sorted(path.nested):  # FxHashMap order is unspecified; ascending here
        tensors[idx] = contract_tensor_network(tensors[idx], path.nested[idx], backend, stats)
    tensors = [_leaf_with_data(t, backend) if not t.is_composite else t for t in tensors]
    for (i, j) in path.toplevel:
        ta, tb = tensors[i], tensors[j]
        tensors[i], tensors[j] = None, None  # mem::take (contraction.rs:61-62)
        if ta is None or tb is None or ta.data is None or tb.data is None:
            raise RuntimeError("Cannot convert uncontracted tensor to data")
        legs, res = contract_pair(ta.legs, ta.data, tb.legs, tb.data, backend)
        if stats is not None:
            m = n = k = 1
            for l, d in zip(ta.legs, ta.dims):
                if l in tb.legs: k *= d
                else: m *= d
            for l, d in zip(tb.legs, tb.dims):
                if l not in ta.legs: n *= d
            stats["pairs"] = stats.get("pairs", 0) + 1
            stats["flops"] = stats.get("flops", 0) + 8 * m * n * k
            stats["bytes"] = stats.get("bytes", 0) + 16 * (m * k + k * n + m * n)
        tensors[i] = OTensor(legs=legs, dims=list(res.shape), data=res)
    # retain non-empty (contraction.rs:48-51)
    rest = [t for t in tensors if t is not None and (t.data is not None or t.is_composite)]
    assert len(rest) <= 1, "Not fully contracted"
    return rest[0] if rest else OTensor()


def _leaf_with_data(t: OTensor, backend: str) -> OTensor:
    if t.data is None:
        return t
    d = t.materialise()
    if backend == "torch":
        import torch
        d = torch.from_numpy(np.ascontiguousarray(d))
    return OTensor(legs=list(t.legs), dims=list(t.dims), data=d)


def permute_to(t: OTensor, target_legs: Sequence[int]) -> OTensor:
    """Permutor::apply (builders/circuit_builder.rs:86-114): transpose the data
    so that the legs appear in `target_legs` order; empty target = identity."""
    if len(target_legs) == 0:
        return t
    perm = [t.legs.index(l) for l in target_legs]
    data = np.ascontiguousarray(np.transpose(np.asarray(t.data), perm))
    return OTensor(legs=list(target_legs), dims=[t.dims[p] for p in perm], data=data)


# ----------------------------------------------------------------------------
# path helpers (tnc/src/contractionpath.rs:197-215)
# ----------------------------------------------------------------------------
def ssa_replace_ordering(path: OPath) -> OPath:
    nested = {i: ssa_replace_ordering(p) for i, p in path.nested.items()}
    hs: Dict[int, int] = {}
    top = []
    n = len(path.toplevel) + 1
    for (t0, t1) in path.toplevel:
        n0, n1 = hs.get(t0, t0), hs.get(t1, t1)
        hs[n] = n0
        n += 1
        top.append((n0, n1))
    return OPath(toplevel=top, nested=nested)


# ----------------------------------------------------------------------------
# circuit builder (tnc/src/builders/circuit_builder.rs:135-335)
# ----------------------------------------------------------------------------
class OCircuit:
    def __init__(self):
        self.open_edges: List[int] = []
        self.next_edge = 0
        self.tensors: List[OTensor] = []

    def allocate_register(self, size: int) -> List[int]:
        base = len(self.open_edges)
        for _ in range(size):  # :185-192
            e = self.next_edge; self.next_edge += 1
            self.open_edges.append(e)
            self.tensors.append(OTensor([e], [2], np.array([1, 0], dtype=np.complex128)))
        return list(range(base, base + size))

    def append_gate(self, name: str, angles: Sequence[float], qubits: Sequence[int], adjoint=False):
        if len(set(qubits)) != len(qubits):
            raise ValueError("Qubit arguments must be unique")  # :206-209
        old = [self.open_edges[q] for q in qubits]
        new = [self.next_edge + e for e in range(len(qubits))]
        self.next_edge += len(qubits)
        for q, e in zip(qubits, new):
            self.open_edges[q] = e
        legs = old + new  # :212-214
        self.tensors.append(OTensor(legs, [2] * len(legs), ("gate", name, tuple(angles), adjoint)))

    def into_amplitude_network(self, bitstring: str):
        assert len(bitstring) == len(self.open_edges)
        final_legs = []
        tensors = list(self.tensors)
        for c, e in zip(bitstring, self.open_edges):  # :248-261
            if c == "*":
                final_legs.append(e); continue
            if c not in "01":
                raise ValueError("Only 0, 1 and * are allowed in bitstring")
            v = np.array([1, 0] if c == "0" else [0, 1], dtype=np.complex128)
            tensors.append(OTensor([e], [2], v))
        return OTensor(children=tensors), final_legs

    def into_statevector_network(self):
        return self.into_amplitude_network("*" * len(self.open_edges))

    def into_expectation_value_network(self) -> OTensor:
        offset = self.next_edge  # :311
        tensors = list(self.tensors)
        for t in self.tensors:  # tensor_adjoint :287-308
            half = len(t.legs) // 2
            legs = [l + offset for l in (t.legs[half:] + t.legs[:half])]
            dims = t.dims[half:] + t.dims[:half]
            if isinstance(t.data, tuple):
                d = ("gate", t.data[1], t.data[2], not t.data[3])
            else:
                d = matrix_adjoint(np.asarray(t.data).reshape(t.dims))
            tensors.append(OTensor(legs, dims, d))
        for e in self.open_edges:  # :324-329
            tensors.append(OTensor([e, e + offset], [2, 2], ("gate", "z", (), False)))
        return OTensor(children=tensors)


# ----------------------------------------------------------------------------
# cost accounting (contraction_cost.rs:26-32,71-74)
# ----------------------------------------------------------------------------
def pair_mnk(a_legs, a_dims, b_legs, b_dims):
    m = n = k = 1
    for l, d in zip(a_legs, a_dims):
        if l in b_legs: k *= d
        else: m *= d
    for l, d in zip(b_legs, b_dims):
        if l not in a_legs: n *= d
    return m, n, k
