"""Planning (outside every timer): partition vectors + nested paths of the bench network for 2 / 4 / 8 ranks.

    python tools/plan_partitions.py            # writes bench_inputs/c4_partitions.json

Pipeline (all seeded, metadata only): FM bisection -> 400-evaluation simulated annealing into 2 parts
(contractionpath/repartitioning.py, the step-budget restatement of the reference's SA balancer) -> the resulting
nested path flattened into one contraction tree -> tree_cut(N) (contractionpath/tree_partition.py): N subtrees +
the N-1 top nodes as the fan-in path.  The critical-path flops of every N are <= those of the 2-part solution."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

OUT = os.path.join(ROOT, "bench_inputs", "c4_partitions.json")
NETWORK = {"qubits": 36, "rounds": 10, "p1": 0.5, "p2": 0.5, "seed": 1}


def build_network():
    from tnc_b200.builders import random_circuit
    return random_circuit(NETWORK["qubits"], NETWORK["rounds"], NETWORK["p1"], NETWORK["p2"], np.random.default_rng(NETWORK["seed"]))


def network_hash(tn) -> str:
    h = hashlib.sha256()
    for t in tn.tensors:
        h.update(repr((list(t.legs), list(t.bond_dims), t.tensordata.kind, t.tensordata.gate if t.tensordata.kind == "gate" else None)).encode())
    return h.hexdigest()[:16]


def plan(tn, parts_list=(2, 4, 8), sa_steps=400, seed=1):
    from tnc_b200.contractionpath.contraction_cost import contract_path_cost
    from tnc_b200.contractionpath.repartitioning import balance_partitions, compute_solution
    from tnc_b200.contractionpath.tree_partition import flatten_nested, tree_cut
    from tnc_b200.tensornetwork.partitioning import find_partitioning
    t0 = time.time()
    init = find_partitioning(tn, 2, seed=seed)
    best, _ = balance_partitions(tn, 2, init, steps=sa_steps, seed=seed)
    ptn, ppath, _, _ = compute_solution(tn, best)
    order = sorted(set(best), key=best.index)
    flat = flatten_nested(ptn, ppath, [order.index(b) for b in best])
    out = {"network": NETWORK, "network_hash": network_hash(tn), "sa_steps": sa_steps, "seed": seed,
           "tree_flops": contract_path_cost(tn.tensors, flat, False)[0], "plans": {}}
    for n in parts_list:
        pv, ptn_n, path_n, crit, total = tree_cut(tn, flat, n)
        out["plans"][str(n)] = {"partitioning": pv, "nested": {str(k): [list(p) for p in v.toplevel] for k, v in path_n.nested.items()},
                                "toplevel": [list(p) for p in path_n.toplevel], "critical_path_flops": crit, "total_flops": total,
                                "partition_sizes": [len(c.tensors) for c in ptn_n.tensors]}
    out["planning_seconds"] = time.time() - t0
    return out


def load(tn, n):
    """(partitioned network, nested path, facts) for n ranks from the committed plan file; None if absent or stale."""
    from tnc_b200.contractionpath import ContractionPath
    from tnc_b200.tensornetwork.partitioning import partition_tensor_network
    try:
        with open(OUT) as f:
            d = json.load(f)
    except Exception:
        return None
    if d.get("network_hash") != network_hash(tn) or str(n) not in d["plans"]:
        return None
    p = d["plans"][str(n)]
    path = ContractionPath({int(k): ContractionPath.simple([tuple(x) for x in v]) for k, v in p["nested"].items()},
                           [tuple(x) for x in p["toplevel"]])
    return partition_tensor_network(tn, p["partitioning"]), path, {k: p[k] for k in ("critical_path_flops", "total_flops", "partition_sizes")}


if __name__ == "__main__":
    tn = build_network()
    d = plan(tn)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        json.dump(d, f)
    for n, p in d["plans"].items():
        print(n, "parts: critical path %.3e flop, total %.3e, sizes %s" % (p["critical_path_flops"], p["total_flops"], p["partition_sizes"]))
    print("planning took %.1f s ->" % d["planning_seconds"], OUT)
