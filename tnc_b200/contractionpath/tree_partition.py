"""Partition a flat network by cutting a contraction tree (planning only, metadata).

The reference obtains its partition vector from KaHyPar + simulated annealing over minutes of 48 threads
(tnc/src/tensornetwork/partitioning.rs:33-175, contractionpath/repartitioning/simulated_annealing.rs:406-592); the
vector is an *input* of the partitioned contraction.  This module derives one deterministically from any
replace-left path of the flat network: the contraction tree is cut below its top `parts - 1` nodes, the subtrees become
the partitions (each keeps the tree's own pairs as its local path) and the cut-off top nodes become the fan-in path of
`intermediate_reduce_tensor_network` (mpi/communication.rs:199-249).  The critical-path cost of the result
(contraction_cost.rs:259-289) can only be <= the cost of the tree, so a good tree gives a good partitioning for every
rank count; which subtree to split next is chosen to shorten the critical path."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

from ..tensornetwork.partitioning import partition_tensor_network
from ..tensornetwork.tensor import Tensor
from . import ContractionPath
from .contraction_cost import contract_cost_tensors


class _Node:
    __slots__ = ("left", "right", "slot", "cost", "total", "leaves", "pair_index")

    def __init__(self, slot, left=None, right=None, cost=0.0, pair_index=-1):
        self.slot, self.left, self.right, self.cost, self.pair_index = slot, left, right, cost, pair_index
        self.total = cost + (left.total if left else 0.0) + (right.total if right else 0.0)
        self.leaves = (left.leaves + right.leaves) if left else 1


def _build_tree(tn: Tensor, path: ContractionPath):
    assert tn.is_composite() and not path.nested and all(t.is_leaf() for t in tn.tensors), "tree_cut works on a flat network and a simple path"
    cur: List[Optional[_Node]] = [_Node(i) for i in range(len(tn.tensors))]
    meta: List[Optional[Tensor]] = [Tensor(t.legs, t.bond_dims) for t in tn.tensors]
    root = None
    for q, (i, j) in enumerate(path.toplevel):
        cost = contract_cost_tensors(meta[i], meta[j])
        root = _Node(i, cur[i], cur[j], cost, q)
        cur[i], cur[j] = root, None
        meta[i], meta[j] = meta[j] ^ meta[i], None
    assert root is not None and sum(c is not None for c in cur) == 1, "path does not contract the network fully"
    return root


def _critical(node: _Node, cut: set) -> float:
    """latency of `node` when every subtree in `cut` runs on its own device and the nodes above are fan-in pairs"""
    if id(node) in cut or node.left is None:
        return node.total
    return node.cost + max(_critical(node.left, cut), _critical(node.right, cut))


def tree_cut(tn: Tensor, path: ContractionPath, parts: int, min_leaves: int = 2):
    """Returns (partitioning vector, partitioned network, nested replace-left path, critical-path flops, serial flops).
    Partition ids are numbered in order of first appearance, which is the order `partition_tensor_network` assigns
    (partitioning.rs:165-175), so nested[k] belongs to composite child k."""
    root = _build_tree(tn, path)
    cut: List[_Node] = [root]
    while len(cut) < parts:
        ids = {id(c) for c in cut}
        best, best_key = None, None
        for c in cut:                      # split the subtree that shortens the critical path most (ties: the biggest)
            if c.left is None or c.left.leaves < min_leaves or c.right.leaves < min_leaves:
                continue
            trial = (ids - {id(c)}) | {id(c.left), id(c.right)}
            key = (_critical(root, trial), -c.total)
            if best_key is None or key < best_key:
                best, best_key = c, key
        if best is None:
            break
        cut.remove(best)
        cut += [best.left, best.right]
    # leaves of every subtree
    part_of: Dict[int, int] = {}

    def mark(node: _Node, pid: int):
        stack = [node]
        while stack:
            n = stack.pop()
            if n.left is None:
                part_of[n.slot] = pid
            else:
                stack += [n.left, n.right]
    for pid, c in enumerate(cut):
        mark(c, pid)
    # renumber in order of first appearance
    order: List[int] = []
    for s in range(len(tn.tensors)):
        if part_of[s] not in order:
            order.append(part_of[s])
    renum = {old: new for new, old in enumerate(order)}
    partitioning = [renum[part_of[s]] for s in range(len(tn.tensors))]
    members: Dict[int, List[int]] = {}
    for s, p in enumerate(partitioning):
        members.setdefault(p, []).append(s)
    local_index = {s: k for p, ms in members.items() for k, s in enumerate(ms)}
    cut_ids = {id(c): renum[pid] for pid, c in enumerate(cut)}
    # pairs below a cut node -> that partition's local path; pairs above -> fan-in path (partition indices)
    owner_of_pair: Dict[int, int] = {}

    def assign(node: _Node, pid: int):
        stack = [node]
        while stack:
            n = stack.pop()
            if n.left is not None:
                owner_of_pair[n.pair_index] = pid
                stack += [n.left, n.right]
    for c in cut:
        assign(c, cut_ids[id(c)])
    nested: Dict[int, List[Tuple[int, int]]] = {p: [] for p in members}
    toplevel: List[Tuple[int, int]] = []
    for q, (i, j) in enumerate(path.toplevel):
        if q in owner_of_pair:
            nested[owner_of_pair[q]].append((local_index[i], local_index[j]))
        else:
            toplevel.append((partitioning[i], partitioning[j]))
    ptn = partition_tensor_network(tn, partitioning)
    npath = ContractionPath({p: ContractionPath.simple(v) for p, v in nested.items()}, toplevel)
    crit = _critical(root, {id(c) for c in cut})
    return partitioning, ptn, npath, crit, root.total


def flatten_nested(ptn: Tensor, path: ContractionPath, partitioning: Sequence[int]) -> ContractionPath:
    """The flat replace-left path equivalent to a partitioned (network, nested path): local paths first (ascending
    child), then the fan-in pairs, with composite-local slots mapped back to flat slots.  `partitioning` must be the
    vector `ptn` was built from (ids in order of first appearance)."""
    members: Dict[int, List[int]] = {}
    for s, p in enumerate(partitioning):
        members.setdefault(p, []).append(s)
    out: List[Tuple[int, int]] = []
    home: Dict[int, int] = {}
    for p in sorted(path.nested):
        ms = members[p]
        for (i, j) in path.nested[p].toplevel:
            out.append((ms[i], ms[j]))
        # the partition's result ends in the slot of its last pair's left operand (or its only tensor)
        home[p] = ms[path.nested[p].toplevel[-1][0]] if path.nested[p].toplevel else ms[0]
    for p in members:
        home.setdefault(p, members[p][0])
    for (x, y) in path.toplevel:
        out.append((home[x], home[y]))
    return ContractionPath.simple(out)
