"""TreeReconfigure: greedy path refined by subtree reconfiguration (tnc/src/contractionpath/paths/tree_reconfiguration.rs).

The reference calls cotengra (`rustengra::cotengra_optimized_greedy(inputs, outputs, size_dict, subtree_size)`,
tree_reconfiguration.rs:54-58), which is not part of the reference tree.  Here the greedy start is the repo's Cotengrust
mirror and the refinement is the native subset-DP in csrc/reconf.cpp (tncb_path_reconfigure): planning on the host, no
GPU work.  For networks with at most `subtree_size` tensors the DP covers the whole tree, i.e. the result is THE optimal
order -- which is what the reference's two tests pin (best_flops 600 and 332685, tree_reconfiguration.rs:141-165).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np

from ..._lib import check, lib
from ...tensornetwork.tensor import Tensor
from ..contraction_cost import contract_path_cost
from .. import ContractionPath, ssa_replace_ordering
from .cotengrust import optimize_greedy


class _Legs:
    """Leg ids -> bit positions; leaf leg sets as rows of 64-bit words (the layout of tncb_path_reconfigure)."""

    def __init__(self, inputs: Sequence[Sequence[int]], size_dict: Dict[int, float]):
        order: Dict[int, int] = {}
        count: Dict[int, int] = {}
        for t in inputs:
            for l in t:
                order.setdefault(l, len(order))
                count[l] = count.get(l, 0) + 1
        if any(c > 2 for c in count.values()):
            raise ValueError("a leg joins more than two tensors: not the reference's tensor model")
        self.pos = order
        self.ids = list(order)
        self.words = max(1, (len(order) + 63) // 64)
        self.log2 = np.zeros(self.words * 64, dtype=np.float64)
        for l, p in order.items():
            self.log2[p] = np.log2(float(size_dict[l]))
        self.n = len(inputs)

    def rows(self, inputs, without=()) -> np.ndarray:
        rows = np.zeros((len(inputs), self.words), dtype=np.uint64)
        skip = set(without)
        for i, t in enumerate(inputs):
            for l in t:
                if l in skip:
                    continue
                p = self.pos[l]
                rows[i, p >> 6] |= np.uint64(1) << np.uint64(p & 63)
        return rows


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _time_model(objective: str):
    if objective == "flops":
        return None
    if objective != "time":
        raise ValueError("objective must be 'flops' or 'time'")
    from ..contraction_cost import GPU_RATES as R
    return np.array([R["crt_flops"], R["crt_k_half"], R["dmma_flops"], R["hbm_bytes"], R["launch_s"], R["dmma_k_half"], R["crt_k_max"],
                     R["crt_conv_bytes"]], dtype=np.float64)


def reconfigure_ssa_path(inputs: Sequence[Sequence[int]], size_dict: Dict[int, float], ssa_path: Sequence[Tuple[int, int]],
                         subtree_size: int = 8, max_sweeps: int = 32, size_weight: float = 0.0, seed: int = 0,
                         sliced: Sequence[int] = (), objective: str = "flops"):
    """Refine `ssa_path`; legs in `sliced` are treated as fixed (dimension 1).  objective "flops" minimises
    sum prod dims(legs(a) | legs(b)) + size_weight * (output elements), "time" the device-time model
    (contraction_cost.gpu_time_tensors).  Returns (ssa_path, flops, max_size, objective_value) of ONE slice."""
    n = len(inputs)
    if n < 2:
        return [], 0.0, 0.0, 0.0
    lg = _Legs(inputs, size_dict)
    rows = lg.rows(inputs, sliced)
    ssa = np.ascontiguousarray(np.asarray(ssa_path, dtype=np.int32).reshape(n - 1, 2))
    tm = _time_model(objective)
    flops, size, obj = C.c_double(), C.c_double(), C.c_double()
    check(lib().tncb_path_reconfigure(n, lg.words, _ptr(rows, C.c_uint64), _ptr(lg.log2, C.c_double), _ptr(ssa, C.c_int),
                                      int(subtree_size), int(max_sweeps), float(size_weight), _ptr(tm, C.c_double) if tm is not None else None,
                                      int(seed), C.byref(flops), C.byref(size), C.byref(obj)))
    return [(int(a), int(b)) for a, b in ssa], flops.value, size.value, obj.value


def leg_scores(inputs, size_dict, ssa_path, sliced=(), size_weight: float = 0.0, objective: str = "flops"):
    """Per-leg slicing scores of a tree: {leg: (cost_without, size_without)} of ONE slice with that leg fixed, plus
    (cost, max_size) of the tree as it is."""
    n = len(inputs)
    lg = _Legs(inputs, size_dict)
    rows = lg.rows(inputs, sliced)
    ssa = np.ascontiguousarray(np.asarray(ssa_path, dtype=np.int32).reshape(n - 1, 2))
    cw = np.zeros(lg.words * 64); sw = np.zeros(lg.words * 64)
    tm = _time_model(objective)
    cost, size = C.c_double(), C.c_double()
    check(lib().tncb_path_leg_scores(n, lg.words, _ptr(rows, C.c_uint64), _ptr(lg.log2, C.c_double), _ptr(ssa, C.c_int), float(size_weight),
                                     _ptr(tm, C.c_double) if tm is not None else None,
                                     _ptr(cw, C.c_double), _ptr(sw, C.c_double), C.byref(cost), C.byref(size)))
    skip = set(sliced)
    return {l: (float(cw[p]), float(sw[p])) for l, p in lg.pos.items() if l not in skip}, cost.value, size.value


def slice_and_reconfigure(inputs, size_dict, ssa_path, target_size: float, subtree_size: int = 10, max_sweeps: int = 8,
                          size_weight: float = 0.0, seed: int = 0, max_slices: int = 64, objective: str = "flops"):
    """Fix legs one at a time (among the legs that shrink the largest tensor, the one that leaves the cheapest total =
    objective of one slice x number of slices), re-running the subtree reconfiguration after each, until the largest tensor
    has at most `target_size` elements.  Returns (sliced_legs, ssa_path, flops_per_slice, max_size, objective_per_slice)."""
    sliced: List[int] = []
    path, flops, size, obj = reconfigure_ssa_path(inputs, size_dict, ssa_path, subtree_size, max_sweeps, size_weight, seed, sliced, objective)
    while size > target_size and len(sliced) < max_slices:
        scores, _, size = leg_scores(inputs, size_dict, path, sliced, size_weight, objective)
        best, best_key = None, None
        for l, (cw, sw) in scores.items():
            key = (0 if sw < size else 1, cw * float(size_dict[l]), sw)
            if best_key is None or key < best_key:
                best, best_key = l, key
        sliced.append(best)
        path, flops, size, obj = reconfigure_ssa_path(inputs, size_dict, path, subtree_size, max_sweeps, size_weight, seed + len(sliced), sliced, objective)
    return sliced, path, flops, size, obj


class TreeReconfigure:
    """Mirror of the reference's `TreeReconfigure::new(&tensor, subtree_size, CostType::Flops)` + FindPath."""

    def __init__(self, tensor: Tensor, subtree_size: int = 8, minimize: str = "flops", max_sweeps: int = 32, seed: int = 0):
        if minimize != "flops":
            raise ValueError("Currently, only Flops is supported")        # tree_reconfiguration.rs:25-29
        self.tensor = tensor
        self.subtree_size = subtree_size
        self.max_sweeps = max_sweeps
        self.seed = seed
        self.best_path = ContractionPath()
        self.best_flops = float("inf")
        self.best_size = float("inf")

    def find_path(self) -> None:
        nested = {}
        inputs = list(self.tensor.tensors)
        for idx, t in enumerate(inputs):
            if t.is_composite():
                sub = TreeReconfigure(t, self.subtree_size, "flops", self.max_sweeps, self.seed)
                sub.find_path()
                nested[idx] = sub.get_best_path()
                inputs[idx] = t.external_tensor()
        legs = [list(t.legs) for t in inputs]
        size_dict = {l: float(d) for t in inputs for l, d in t.edges()}
        ssa = optimize_greedy(legs, list(self.tensor.external_tensor().legs), size_dict) if len(legs) > 1 else []
        if len(legs) > 2:
            ssa = reconfigure_ssa_path(legs, size_dict, ssa, self.subtree_size, self.max_sweeps, 0.0, self.seed)[0]
        self.best_path = ContractionPath(nested, [tuple(p) for p in ssa])
        self.best_flops, self.best_size = contract_path_cost(self.tensor.tensors, self.get_best_replace_path(), True)

    def get_best_path(self) -> ContractionPath:
        return self.best_path

    def get_best_replace_path(self) -> ContractionPath:
        return ssa_replace_ordering(self.best_path)

    def get_best_flops(self) -> float:
        return self.best_flops

    def get_best_size(self) -> float:
        return self.best_size
