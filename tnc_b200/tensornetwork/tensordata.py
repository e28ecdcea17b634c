"""TensorData (tnc/src/tensornetwork/tensordata.rs:15-26)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import numpy as np


@dataclass
class TensorData:
    """kind: 'uncontracted' | 'gate' | 'matrix' | 'file'.
    matrix payloads are either a host ndarray or a tnc_b200.DeviceTensor."""
    kind: str = "uncontracted"
    gate: Optional[Tuple[str, Tuple[float, ...], bool]] = None
    matrix: object = None
    file: Optional[Tuple[str, bool]] = None

    # constructors mirroring the enum variants
    @classmethod
    def Uncontracted(cls) -> "TensorData":
        return cls()

    @classmethod
    def Gate(cls, name: str, angles: Sequence[float] = (), adjoint: bool = False) -> "TensorData":
        return cls(kind="gate", gate=(name, tuple(float(a) for a in angles), bool(adjoint)))

    @classmethod
    def Matrix(cls, data) -> "TensorData":
        return cls(kind="matrix", matrix=data)

    @classmethod
    def File(cls, path: str, adjoint: bool = False) -> "TensorData":
        return cls(kind="file", file=(path, adjoint))

    @classmethod
    def new_from_data(cls, dimensions: Sequence[int], data, layout=None) -> "TensorData":
        """tensordata.rs:31-37; flat row-major data."""
        if layout is not None:
            raise NotImplementedError("only the default (row-major) layout is supported")
        arr = np.asarray(data, dtype=np.complex128).reshape(tuple(dimensions))
        return cls.Matrix(np.ascontiguousarray(arr))

    def adjoint(self) -> "TensorData":
        """tensordata.rs:62-72."""
        if self.kind == "uncontracted":
            return TensorData()
        if self.kind == "gate":
            n, a, adj = self.gate
            return TensorData.Gate(n, a, not adj)
        if self.kind == "file":
            return TensorData.File(self.file[0], not self.file[1])
        m = self.matrix
        if not isinstance(m, np.ndarray):
            m = m.to_numpy()
        if m.ndim > 0:
            half = m.ndim // 2
            m = np.transpose(m, list(range(half, m.ndim)) + list(range(half)))
        return TensorData.Matrix(np.ascontiguousarray(np.conj(m)))
