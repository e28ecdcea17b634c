"""Greedy path finder: mirrors tnc::contractionpath::paths::cotengrust::Cotengrust
(tnc/src/contractionpath/paths/cotengrust.rs:25-160) for OptMethod::Greedy.

The reference delegates to the external crate `cotengrust` 0.2.0 @ Ectras/cotengrust#2998e988
(`optimize_greedy_rust(inputs, output, size_dict, None, None, None, Some(42), false, true)`,
cotengrust.rs:58-68), which is not vendored.  This is a restatement of its published
algorithm (cotengra's greedy with temperature 0, so the seed is unused): f32 log-sizes, score
= logsub(size(out), logadd(size(a), size(b)) + ln(costmod)), a max-heap on (-score, -insertion
counter), candidates per shared index over node pairs, new candidates against the ascending
neighbours of every new node, and leftover disconnected nodes merged smallest-first.  It is
pinned by the reference's five greedy KATs (cotengrust.rs:241-305, tests/test_paths.py).
Known gap: cotengrust enumerates the initial candidates in FxHashMap order; ascending index
order is used here, so score *ties* on large networks may break differently.
"""
from __future__ import annotations

import heapq
from enum import Enum
from typing import Dict, List, Tuple

import numpy as np

from ...tensornetwork.tensor import Tensor
from .. import ContractionPath, ssa_replace_ordering
from ..contraction_cost import contract_path_cost

f32 = np.float32


class OptMethod(Enum):
    Greedy = "greedy"
    RandomGreedy = "random_greedy"   # OptMethod::RandomGreedy(ntrials), cotengrust.rs:69-83


def _logadd(lx: f32, ly: f32) -> f32:
    mx = max(lx, ly)
    return f32(mx + np.log1p(np.exp(-abs(f32(lx - ly)), dtype=f32), dtype=f32))


def _logsub(lx: f32, ly: f32) -> f32:
    if lx < ly:
        return f32(-ly - np.log1p(-np.exp(f32(lx - ly), dtype=f32), dtype=f32))
    if lx > ly:
        return f32(lx + np.log1p(-np.exp(f32(ly - lx), dtype=f32), dtype=f32))
    return f32(0.0)


class _Processor:
    """cotengrust's ContractionProcessor (metadata only)."""

    def __init__(self, inputs: List[List[int]], output: List[int], size_dict: Dict[int, float]):
        self.indmap: Dict[int, int] = {}
        self.sizes: List[f32] = []
        self.appearances: List[int] = []
        self.nodes: Dict[int, List[Tuple[int, int]]] = {}
        self.edges: Dict[int, set] = {}
        for i, term in enumerate(inputs):
            legs: Dict[int, int] = {}
            for ind in term:
                ix = self.indmap.get(ind)
                if ix is None:
                    ix = len(self.indmap)
                    self.indmap[ind] = ix
                    self.edges[ix] = {i}
                    self.appearances.append(1)
                    self.sizes.append(f32(np.log(f32(size_dict[ind]), dtype=f32)))
                else:
                    self.appearances[ix] += 1
                    self.edges[ix].add(i)
                legs[ix] = legs.get(ix, 0) + 1
            self.nodes[i] = sorted(legs.items())
        for ind in output:
            self.appearances[self.indmap[ind]] += 1
        self.ssa = len(inputs)
        self.ssa_path: List[Tuple[int, int]] = []

    def size(self, legs) -> f32:
        s = f32(0.0)
        for ix, _ in legs:
            s = f32(s + self.sizes[ix])
        return s

    def compute_legs(self, la, lb):
        out, ia, ib = [], 0, 0
        while ia < len(la) and ib < len(lb):
            (xa, ca), (xb, cb) = la[ia], lb[ib]
            if xa < xb:
                out.append((xa, ca)); ia += 1
            elif xb < xa:
                out.append((xb, cb)); ib += 1
            else:
                c = ca + cb
                if c != self.appearances[xa]:
                    out.append((xa, c))
                ia += 1; ib += 1
        out.extend(la[ia:]); out.extend(lb[ib:])
        return out

    def neighbors(self, i: int):
        js = set()
        for ix, _ in self.nodes[i]:
            js |= self.edges[ix]
        js.discard(i)
        return sorted(js)

    def contract_given_legs(self, i: int, j: int, klegs) -> int:
        for n in (i, j):
            for ix, _ in self.nodes.pop(n):
                self.edges[ix].discard(n)
        k = self.ssa
        self.ssa += 1
        for ix, _ in klegs:
            self.edges[ix].add(k)
        for ix in [ix for ix, s in self.edges.items() if not s]:
            del self.edges[ix]
        self.nodes[k] = klegs
        self.ssa_path.append((i, j))
        return k

    def contract(self, i: int, j: int) -> int:
        return self.contract_given_legs(i, j, self.compute_legs(self.nodes[i], self.nodes[j]))

    def optimize_greedy(self, costmod: float = 1.0, temperature: float = 0.0, rng=None) -> None:
        log_a = f32(np.log(f32(costmod), dtype=f32))
        coeff_t = f32(temperature)
        node_size = {i: self.size(l) for i, l in self.nodes.items()}
        heap, contractions, c = [], {}, 0

        def score(sa, sb, sab):
            sc = _logsub(sab, f32(_logadd(sa, sb) + log_a))
            if coeff_t != 0.0:  # Gumbel noise, as in cotengra's random-greedy
                sc = f32(sc - coeff_t * f32(-np.log(-np.log(rng.random()))))
            return sc

        for ix in sorted(self.edges):
            ns = sorted(self.edges[ix])
            for p in range(len(ns)):
                for q in range(p + 1, len(ns)):
                    i, j = ns[p], ns[q]
                    klegs = self.compute_legs(self.nodes[i], self.nodes[j])
                    ksize = self.size(klegs)
                    # python's heapq is a min-heap: (score, -c) pops the smallest score, then the
                    # earliest insertion == Rust's max-heap on (-score, c) with c decreasing
                    heapq.heappush(heap, (float(score(node_size[i], node_size[j], ksize)), -c))
                    contractions[c] = (i, j, ksize, klegs)
                    c -= 1
        while heap:
            _, negc = heapq.heappop(heap)
            i, j, ksize, klegs = contractions.pop(-negc)
            if i not in self.nodes or j not in self.nodes:
                continue
            k = self.contract_given_legs(i, j, klegs)
            if len(self.nodes) == 1:
                return
            node_size[k] = ksize
            for l in self.neighbors(k):
                mlegs = self.compute_legs(klegs, self.nodes[l])
                msize = self.size(mlegs)
                heapq.heappush(heap, (float(score(ksize, node_size[l], msize)), -c))
                contractions[c] = (k, l, msize, mlegs)
                c -= 1

    def optimize_remaining_by_size(self) -> None:
        if len(self.nodes) <= 1:
            return
        # Rust max-heap on (-size, node): smallest size first, ties -> larger node id first
        heap = [(float(self.size(l)), -n) for n, l in self.nodes.items()]
        heapq.heapify(heap)
        _, ni = heapq.heappop(heap); _, nj = heapq.heappop(heap)
        k = self.contract(-ni, -nj)
        while len(self.nodes) > 1:
            heapq.heappush(heap, (float(self.size(self.nodes[k])), -k))
            _, ni = heapq.heappop(heap); _, nj = heapq.heappop(heap)
            k = self.contract(-ni, -nj)


def optimize_greedy(inputs: List[List[int]], output: List[int], size_dict: Dict[int, float]) -> List[Tuple[int, int]]:
    """SSA path, like `optimize_greedy_rust(..., use_ssa=true)`."""
    if not inputs:
        return []
    p = _Processor(inputs, output, size_dict)
    p.optimize_greedy()
    p.optimize_remaining_by_size()
    return p.ssa_path


def optimize_random_greedy(inputs, output, size_dict, ntrials: int, seed: int = 42, objective: str = "flops"):
    """Random-greedy (cotengra): repeated greedy runs with a random cost modifier in [0, 50) (first trial: plain
    greedy) and a log-uniform temperature in [0.001, 1]; keeps the path with the lowest flop count (or peak size).
    The reference calls cotengrust's implementation (cotengrust.rs:69-83); RNG streams are not reproduced."""
    rng = np.random.default_rng(seed)
    best, best_cost = None, None
    for trial in range(ntrials):
        p = _Processor(inputs, output, size_dict)
        if trial == 0:
            p.optimize_greedy()
        else:
            costmod = float(rng.uniform(0.0, 50.0)) or 1e-3
            temp = float(np.exp(rng.uniform(np.log(1e-3), np.log(1.0))))
            p.optimize_greedy(max(costmod, 1e-3), temp, rng)
        p.optimize_remaining_by_size()
        cost = _ssa_path_cost(inputs, output, size_dict, p.ssa_path, objective)
        if best_cost is None or cost < best_cost:
            best, best_cost = p.ssa_path, cost
    return best


def _ssa_path_cost(inputs, output, size_dict, ssa_path, objective="flops"):
    legs = {i: set(t) for i, t in enumerate(inputs)}
    count: Dict[int, int] = {}
    for t in inputs:
        for l in t:
            count[l] = count.get(l, 0) + 1
    for l in output:
        count[l] = count.get(l, 0) + 1
    nxt, flops, peak = len(inputs), 0.0, 0.0
    for (i, j) in ssa_path:
        a, b = legs.pop(i), legs.pop(j)
        allv = a | b
        out = set()
        for l in allv:
            c = (l in a) + (l in b)
            if c == count[l]:
                continue
            out.add(l)
        for l in a & b:
            count[l] -= 1
        fl = 1.0
        for l in allv:
            fl *= size_dict[l]
        sz = 1.0
        for l in out:
            sz *= size_dict[l]
        flops += fl
        peak = max(peak, sz)
        legs[nxt] = out
        nxt += 1
    return flops if objective == "flops" else peak


class Cotengrust:
    def __init__(self, tensor: Tensor, opt_method: OptMethod = OptMethod.Greedy, ntrials: int = 32, objective: str = "flops"):
        self.tensor = tensor
        self.opt_method = opt_method
        self.ntrials = ntrials
        self.objective = objective
        self.best_path = ContractionPath()
        self.best_flops = float("inf")
        self.best_size = float("inf")

    def _optimize_single(self, inputs: List[Tensor], output: Tensor):
        if not inputs:
            return []
        size_dict = {l: float(d) for t in inputs for l, d in t.edges()}
        if self.opt_method == OptMethod.RandomGreedy:
            return optimize_random_greedy([list(t.legs) for t in inputs], list(output.legs), size_dict, self.ntrials, 42, self.objective)
        return optimize_greedy([list(t.legs) for t in inputs], list(output.legs), size_dict)

    def find_path(self) -> None:
        """cotengrust.rs:125-151: nested children first, then the toplevel over external tensors."""
        nested = {}
        inputs = list(self.tensor.tensors)
        for idx, t in enumerate(inputs):
            if t.is_composite():
                ct = Cotengrust(t, self.opt_method, self.ntrials, self.objective)
                ct.find_path()
                nested[idx] = ct.get_best_path()
                inputs[idx] = t.external_tensor()
        outer = self._optimize_single(inputs, self.tensor.external_tensor())
        self.best_path = ContractionPath(nested, outer)
        self.best_flops, self.best_size = contract_path_cost(self.tensor.tensors, self.get_best_replace_path(), True)

    def get_best_path(self) -> ContractionPath:
        return self.best_path

    def get_best_replace_path(self) -> ContractionPath:
        return ssa_replace_ordering(self.best_path)

    def get_best_flops(self) -> float:
        return self.best_flops

    def get_best_size(self) -> float:
        return self.best_size
