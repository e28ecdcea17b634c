"""Circuit -> tensor network (tnc/src/builders/circuit_builder.rs:135-335) and Permutor (:72-129)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .. import Context, DeviceTensor, default_context
from .._lib import check
from ..tensornetwork.tensor import Tensor
from ..tensornetwork.tensordata import TensorData


class Permutor:
    def __init__(self, target_legs: Sequence[int]):
        self.target_leg_order = list(target_legs)

    def is_identity(self) -> bool:
        return not self.target_leg_order

    @staticmethod
    def permutation_between(given: Sequence[int], target: Sequence[int]) -> List[int]:
        """circuit_builder.rs:125-129: the permutation p with [given[p[i]] for i] == target (`given` and `target` are equal
        up to order) -- the axis order handed to tncb_permute."""
        pos = {l: i for i, l in enumerate(given)}
        assert len(pos) == len(given) == len(target) and set(pos) == set(target), "given and target must be permutations of each other"
        return [pos[l] for l in target]

    def apply(self, tensor: Tensor, ctx: Optional[Context] = None) -> Tensor:
        """Transposes the (device) data so that the legs appear in the target order
        (circuit_builder.rs:86-114); one tncb_permute launch."""
        assert tensor.is_leaf()
        if self.is_identity():
            return tensor
        ctx = ctx or default_context()
        perm = self.permutation_between(tensor.legs, self.target_leg_order)
        m = tensor.tensordata.matrix
        dt = m if isinstance(m, DeviceTensor) else DeviceTensor.from_numpy(ctx, np.asarray(m).reshape(tensor.bond_dims))
        out = C.c_void_p()
        parr = (C.c_int * max(len(perm), 1))(*perm)
        check(ctx._l.tncb_permute(ctx.handle, dt.handle, parr, C.byref(out)))
        dt.release()
        res = Tensor(self.target_leg_order, [tensor.bond_dims[p] for p in perm])
        res.set_tensor_data(TensorData.Matrix(DeviceTensor.adopt(ctx, out)))
        return res


class Circuit:
    def __init__(self):
        self.open_edges: List[int] = []
        self.next_edge = 0
        self.tensors: List[Tensor] = []

    @staticmethod
    def _ket(bit: int) -> TensorData:
        return TensorData.new_from_data([2], [1, 0] if bit == 0 else [0, 1])

    def num_qubits(self) -> int:
        return len(self.open_edges)

    def allocate_register(self, size: int) -> List[int]:
        base = self.num_qubits()
        for _ in range(size):
            e = self.next_edge
            self.next_edge += 1
            self.open_edges.append(e)
            t = Tensor.new_from_const([e], 2)
            t.set_tensor_data(self._ket(0))
            self.tensors.append(t)
        return list(range(base, base + size))

    def append_gate(self, gate, angles: Sequence[float] = (), qubits: Sequence[int] = (), adjoint: bool = False) -> None:
        """append_gate(TensorData.Gate(...), qubits=[...]) or append_gate("h", [], [q])."""
        td = gate if isinstance(gate, TensorData) else TensorData.Gate(gate, angles, adjoint)
        if len(set(qubits)) != len(qubits):
            raise ValueError("Qubit arguments must be unique")
        old = [self.open_edges[q] for q in qubits]
        new = [self.next_edge + e for e in range(len(qubits))]
        self.next_edge += len(qubits)
        for q, e in zip(qubits, new):
            self.open_edges[q] = e
        t = Tensor.new_from_const(old + new, 2)
        t.set_tensor_data(td)
        self.tensors.append(t)

    def into_amplitude_network(self, bitstring: str) -> Tuple[Tensor, Permutor]:
        assert len(bitstring) == self.num_qubits()
        tensors = list(self.tensors)
        final_legs = []
        for c, e in zip(bitstring, self.open_edges):
            if c == "*":
                final_legs.append(e)
                continue
            if c not in "01":
                raise ValueError("Only 0, 1 and * are allowed in bitstring")
            t = Tensor.new_from_const([e], 2)
            t.set_tensor_data(self._ket(int(c)))
            tensors.append(t)
        return Tensor.new_composite(tensors), Permutor(final_legs)

    def into_statevector_network(self) -> Tuple[Tensor, Permutor]:
        return self.into_amplitude_network("*" * self.num_qubits())

    def into_expectation_value_network(self) -> Tensor:
        offset = self.next_edge
        tensors = list(self.tensors)
        for t in self.tensors:
            half = len(t.legs) // 2
            legs = [l + offset for l in (t.legs[half:] + t.legs[:half])]
            dims = t.bond_dims[half:] + t.bond_dims[:half]
            adj = Tensor(legs, dims)
            adj.set_tensor_data(t.tensordata.adjoint())
            tensors.append(adj)
        for e in self.open_edges:
            z = Tensor.new_from_const([e, e + offset], 2)
            z.set_tensor_data(TensorData.Gate("z"))
            tensors.append(z)
        return Tensor.new_composite(tensors)
