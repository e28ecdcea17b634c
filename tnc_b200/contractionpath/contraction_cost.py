"""Analytic cost model (tnc/src/contractionpath/contraction_cost.rs:26-32,71-74,146-193).
These formulas also define the FLOP / byte accounting of the benchmark."""
from __future__ import annotations

from typing import List, Tuple

from ..tensornetwork.tensor import Tensor
from . import ContractionPath


def contract_cost_tensors(t1: Tensor, t2: Tensor) -> float:
    k = (t1 & t2).size()
    return ((k - 1.0) * 2.0 + k * 6.0) * (t1 ^ t2).size()


def contract_op_cost_tensors(t1: Tensor, t2: Tensor) -> float:
    return (t1 | t2).size()


def contract_size_tensors(t1: Tensor, t2: Tensor) -> float:
    return (t1 ^ t2).size() + t1.size() + t2.size()


def contract_path_cost(inputs: List[Tensor], path: ContractionPath, only_count_ops: bool = False) -> Tuple[float, float]:
    cost_fn = contract_op_cost_tensors if only_count_ops else contract_cost_tensors
    op, mem = 0.0, 0.0
    inputs = list(inputs)
    for i in sorted(path.nested):
        o, m = contract_path_cost(inputs[i].tensors, path.nested[i], only_count_ops)
        op += o
        mem = max(mem, m)
        inputs[i] = inputs[i].external_tensor()
    for (i, j) in path.toplevel:
        op += cost_fn(inputs[i], inputs[j])
        mem = max(mem, contract_size_tensors(inputs[i], inputs[j]))
        inputs[i] = inputs[i] ^ inputs[j]
    return op, mem
