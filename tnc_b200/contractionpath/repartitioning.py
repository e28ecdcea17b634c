"""Partition evaluation and refinement: mirrors tnc::contractionpath::repartitioning
(`compute_solution`, repartitioning.rs:25-76) and the simulated-annealing balancer
(repartitioning/simulated_annealing.rs: `evaluate_partitioning` :185-214, the intermediate-tensor move
model :251-352, acceptance rule :118-127, temperatures 2.0 -> 0.05 :583-590).

Planning only (metadata): the output -- a partition vector -- is an *input* of the partitioned
contraction.  Differences to the reference, on purpose: the annealing schedule is driven by a
*step budget* instead of wall-clock time and uses one seeded chain instead of 48 rayon chains, so a
seed reproduces a partitioning (the reference's result depends on machine speed, :107,153-160)."""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..tensornetwork.partitioning import partition_tensor_network
from ..tensornetwork.tensor import Tensor
from . import ContractionPath
from . import contraction_cost as _cc
from .contraction_cost import contract_path_cost
from .paths.cotengrust import Cotengrust


def communication_path_op_costs(inputs: Sequence[Tensor], path, tensor_cost: Sequence[float]):
    """contraction_cost.rs:196-208 with only_count_ops = true: ((parallel, serial), memory)."""
    return _cc.communication_path_op_costs(inputs, path, True, tensor_cost)


LocalCache = Dict[Tuple, Tuple[List[Tuple[int, int]], float, Tuple[Tuple[int, ...], Tuple[int, ...]]]]


def _local(tn: Tensor, ids: Tuple[int, ...], cache: Optional[LocalCache], cost_fn=None):
    """greedy local path (replace-left) + op cost + external legs of one partition.  `cache` belongs to ONE
    call of compute_solution / balance_partitions (i.e. one network); its key also carries the members'
    legs and dims so that a cache can never answer for another network's tensors."""
    key = (ids, tuple((tuple(tn.tensors[i].legs), tuple(tn.tensors[i].bond_dims)) for i in ids), cost_fn)
    hit = cache.get(key) if cache is not None else None
    if hit is None:
        comp = Tensor.new_composite([tn.tensors[i] for i in ids])
        opt = Cotengrust(comp)
        opt.find_path()
        p = opt.get_best_replace_path()
        if cost_fn is None:
            cost, _ = contract_path_cost(comp.tensors, p, True)
        else:
            cost, _ = _cc._path_custom_cost(comp.tensors, p, cost_fn, _cc.contract_size_tensors)
        ext = comp.external_tensor()
        hit = (p.toplevel, cost, (tuple(ext.legs), tuple(ext.bond_dims)))
        if cache is not None:
            cache[key] = hit
    return hit


def compute_solution(tn: Tensor, partitioning: Sequence[int], cache: Optional[LocalCache] = None, cost_fn=None, fanin_cost_fn=None):
    """repartitioning.rs:25-76 with CommunicationScheme::Greedy: returns
    (partitioned_tn, path, parallel_cost, sum_cost).  `cost_fn` / `fanin_cost_fn` (default: the reference's operation
    count) replace the per-pair cost of local and fan-in pairs, e.g. contraction_cost.gpu_time_tensors."""
    ptn = partition_tensor_network(tn, partitioning)
    ids_order: List[int] = []
    for p in partitioning:
        if p not in ids_order:
            ids_order.append(p)
    nested, costs, exts = {}, [], []
    for k, pid in enumerate(ids_order):
        ids = tuple(i for i, q in enumerate(partitioning) if q == pid)
        top, cost, (el, ed) = _local(tn, ids, cache, cost_fn)
        nested[k] = ContractionPath.simple(top)
        costs.append(cost)
        exts.append(Tensor(list(el), list(ed)))
    comm = Cotengrust(Tensor.new_composite(exts))
    comm.find_path()
    toplevel = comm.get_best_replace_path().toplevel
    if cost_fn is None and fanin_cost_fn is None:
        (par, ser), _ = communication_path_op_costs(exts, toplevel, costs)
    else:
        f = fanin_cost_fn or cost_fn
        par, _ = _cc._communication_custom_cost(exts, toplevel, f, True, costs)
        ser, _ = _cc._communication_custom_cost(exts, toplevel, f, False, costs)
    return ptn, ContractionPath(nested, toplevel), par, ser


def _trial_move(tn: Tensor, num_partitions: int, cur: List[int], rng, cache: Optional[LocalCache] = None) -> Optional[List[int]]:
    trial = list(cur)
    src = int(rng.integers(0, num_partitions))
    members = [i for i, q in enumerate(trial) if q == src]
    if len(members) < 3:
        return None
    dst = int(rng.integers(0, num_partitions - 1))
    dst += dst >= src
    if rng.random() < 0.25:
        trial[members[int(rng.integers(0, len(members)))]] = dst
        return trial
    top, _, _ = _local(tn, tuple(members), cache)
    if len(top) < 2:
        return None
    pi = int(rng.integers(0, len(top) - 1))
    leaves = {top[pi][0], top[pi][1]}
    for (i, j) in reversed(top[:pi]):
        if i in leaves:
            leaves.add(j)
    if len(leaves) >= len(members):
        return None
    for li in leaves:
        trial[members[li]] = dst
    return trial


def balance_partitions(tn: Tensor, num_partitions: int, initial: Sequence[int], steps: int = 400, seed: int = 42,
                       n_trials: int = 8, restart_iter: int = 50, t_start: float = 2.0, t_end: float = 0.05,
                       cost_fn=None, fanin_cost_fn=None) -> Tuple[List[int], float]:
    """Simulated annealing over partitionings (score = critical-path op cost of `compute_solution`).
    Structure of simulated_annealing.rs:80-160: every iteration runs `n_trials` independent trial moves
    from the current solution (each accepted with probability exp(-log2(score/current)/T)), continues
    from the best of them, restarts from the best-so-far after `restart_iter` iterations without
    improvement; T goes 2.0 -> 0.05 log-linearly over the *step budget* (`steps` evaluations).
    Move model: the sub-tree below a random pair of a partition's local path moves to another
    partition (intermediate-tensor model :251-352); with probability 1/4 a single tensor moves (:216-249)."""
    rng = np.random.default_rng(seed)
    cache: LocalCache = {}                     # per call: never shared between networks
    cur = list(initial)
    _, _, cur_score, _ = compute_solution(tn, cur, cache, cost_fn, fanin_cost_fn)
    best, best_score = list(cur), cur_score
    iters = max(1, steps // n_trials)
    last_improvement = 0
    for it in range(iters):
        temp = 2.0 ** (np.log2(t_start) + (np.log2(t_end) - np.log2(t_start)) * it / max(1, iters - 1))
        cand, cand_score = None, None
        for _ in range(n_trials):
            t_sol, t_score = cur, cur_score
            trial = _trial_move(tn, num_partitions, cur, rng, cache)
            if trial is not None:
                _, _, score, _ = compute_solution(tn, trial, cache, cost_fn, fanin_cost_fn)
                if np.exp(-np.log2(score / cur_score) / temp) >= rng.random():
                    t_sol, t_score = trial, score
            if cand_score is None or t_score < cand_score:
                cand, cand_score = t_sol, t_score
        cur, cur_score = list(cand), cand_score
        if cur_score < best_score:
            best, best_score = list(cur), cur_score
            last_improvement = 0
        last_improvement += 1
        if last_improvement == restart_iter:
            cur, cur_score = list(best), best_score
    return best, best_score
