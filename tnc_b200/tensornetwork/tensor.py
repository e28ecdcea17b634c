"""Tensor (tnc/src/tensornetwork/tensor.rs:21-37) with its leg-set algebra (:383-498)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .tensordata import TensorData


class Tensor:
    def __init__(self, legs: Sequence[int] = (), bond_dims: Sequence[int] = (),
                 tensors: Optional[List["Tensor"]] = None, tensordata: Optional[TensorData] = None):
        assert len(legs) == len(bond_dims)
        self.tensors: List[Tensor] = list(tensors) if tensors else []
        self.legs: List[int] = [int(l) for l in legs]
        self.bond_dims: List[int] = [int(d) for d in bond_dims]
        self.tensordata: TensorData = tensordata if tensordata is not None else TensorData()

    # ---- constructors (tensor.rs:43-93) ----
    @classmethod
    def new(cls, legs, bond_dims) -> "Tensor":
        return cls(legs, bond_dims)

    @classmethod
    def new_from_map(cls, legs, bond_dims_map: Dict[int, int]) -> "Tensor":
        return cls(legs, [bond_dims_map[l] for l in legs])

    @classmethod
    def new_from_const(cls, legs, bond_dim: int) -> "Tensor":
        return cls(legs, [bond_dim] * len(legs))

    @classmethod
    def new_composite(cls, tensors: Iterable["Tensor"]) -> "Tensor":
        return cls(tensors=list(tensors))

    # ---- accessors ----
    def tensor(self, i: int) -> "Tensor":
        return self.tensors[i]

    def nested_tensor(self, idx: Sequence[int]) -> "Tensor":
        t = self
        for i in idx:
            t = t.tensors[i]
        return t

    def total_num_tensors(self) -> int:
        return sum(t.total_num_tensors() for t in self.tensors) if self.is_composite() else 1

    def shape(self) -> List[int]:
        return list(self.bond_dims)

    def dims(self) -> int:
        return len(self.legs)

    def size(self) -> float:
        s = 1.0
        for d in self.bond_dims:
            s *= float(d)
        return s

    def is_leaf(self) -> bool:
        return not self.tensors

    def is_composite(self) -> bool:
        return bool(self.tensors)

    def is_empty(self) -> bool:
        return not self.tensors and not self.legs and self.tensordata.kind == "uncontracted"

    def push_tensor(self, t: "Tensor") -> None:
        assert not self.legs and self.tensordata.kind == "uncontracted", "Cannot push tensors into a leaf tensor"
        self.tensors.append(t)

    def push_tensors(self, ts: Iterable["Tensor"]) -> None:
        assert not self.legs and self.tensordata.kind == "uncontracted", "Cannot push tensors into a leaf tensor"
        self.tensors.extend(ts)

    def tensor_data(self) -> TensorData:
        return self.tensordata

    def set_tensor_data(self, td: TensorData) -> None:
        assert self.is_leaf() or td.kind == "uncontracted", "Cannot add data to composite tensor"
        self.tensordata = td

    def edges(self):
        return zip(self.legs, self.bond_dims)

    # ---- leg-set algebra (tensor.rs:383-479) ----
    def difference(self, other: "Tensor") -> "Tensor":
        p = [(l, d) for l, d in self.edges() if l not in other.legs]
        return Tensor([l for l, _ in p], [d for _, d in p])

    def union(self, other: "Tensor") -> "Tensor":
        legs, dims = list(self.legs), list(self.bond_dims)
        for l, d in other.edges():
            if l not in self.legs:
                legs.append(l); dims.append(d)
        return Tensor(legs, dims)

    def intersection(self, other: "Tensor") -> "Tensor":
        p = [(l, d) for l, d in self.edges() if l in other.legs]
        return Tensor([l for l, _ in p], [d for _, d in p])

    def symmetric_difference(self, other: "Tensor") -> "Tensor":
        p = [(l, d) for l, d in self.edges() if l not in other.legs]
        p += [(l, d) for l, d in other.edges() if l not in self.legs]
        return Tensor([l for l, _ in p], [d for _, d in p])

    __sub__ = difference
    __or__ = union
    __and__ = intersection
    __xor__ = symmetric_difference

    def external_tensor(self) -> "Tensor":
        """tensor.rs:482-498."""
        if self.is_leaf():
            return Tensor(self.legs, self.bond_dims, tensordata=self.tensordata)
        ext = Tensor()
        for t in self.tensors:
            ext = ext ^ (t.external_tensor() if t.is_composite() else t)
        return ext

    def is_connected(self) -> bool:
        n = len(self.tensors)
        parent = list(range(n))

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        for i in range(n):
            for j in range(i + 1, n):
                if set(self.tensors[i].legs) & set(self.tensors[j].legs):
                    parent[find(i)] = find(j)
        return len({find(i) for i in range(n)}) == 1

    # ---- result access ----
    def to_numpy(self) -> np.ndarray:
        """Row-major elements of a leaf with Matrix data (tetra `elements()`)."""
        if self.tensordata.kind != "matrix":
            raise RuntimeError("Cannot convert uncontracted tensor to data")
        m = self.tensordata.matrix
        return m if isinstance(m, np.ndarray) else m.to_numpy()

    def __repr__(self):
        if self.is_composite():
            return f"Tensor(composite, {len(self.tensors)} children)"
        return f"Tensor(legs={self.legs}, bond_dims={self.bond_dims}, data={self.tensordata.kind})"
