"""Offline path search (planning, CPU only) for the big single-amplitude networks (BASELINE config 5): parallel trials of
    random-greedy start  ->  subtree reconfiguration (csrc/reconf.cpp)  ->  slice one leg + reconfigure, until the largest
    tensor fits `--width` (log2 elements),
ranked by the GPU time model (contraction_cost.gpu_time_tensors per pair x number of slices).  The reference gets such
paths from cotengra (hyperoptimization.rs:69-76, tree_reconfiguration.rs:54-58), which is not in this image.  Writes the
replace-left path and the sliced legs as JSON for tools/bench_sliced.py.
usage: python tools/search_path.py --depth 12 --width 31 --trials 4 --workers 8 --out bench_inputs/sycamore53_d12.json"""
import argparse
import json
import math
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(kind, qubits, depth, seed):
    from tnc_b200.builders import random_circuit, sycamore_circuit
    if kind == "sycamore":
        return sycamore_circuit(qubits, depth, np.random.default_rng(seed)).into_amplitude_network("0" * qubits)[0]
    return random_circuit(qubits, depth, 0.5, 0.5, np.random.default_rng(seed))


def worker(args):
    kind, qubits, depth, net_seed, trials, seed, width, subtree = args
    from tnc_b200.contractionpath import ContractionPath, ssa_replace_ordering
    from tnc_b200.contractionpath.paths.cotengrust import _Processor
    from tnc_b200.contractionpath.paths import slice_and_reconfigure
    from tnc_b200.contractionpath.slicing import path_time
    tn = build(kind, qubits, depth, net_seed)
    inputs = [list(t.legs) for t in tn.tensors]
    meta = [(t.legs, t.bond_dims) for t in tn.tensors]
    size_dict = {l: float(d) for t in tn.tensors for l, d in t.edges()}
    rng = np.random.default_rng(seed)
    best = None
    for trial in range(trials):
        p = _Processor(inputs, [], size_dict)
        if trial == 0 and seed % 8 == 0:
            p.optimize_greedy(1.0, 0.0, rng)
        else:
            p.optimize_greedy(float(rng.uniform(0.1, 4.0)), float(np.exp(rng.uniform(np.log(1e-3), np.log(0.3)))), rng)
        p.optimize_remaining_by_size()
        objective = "time" if trial % 2 == 0 else "flops"
        size_weight = 0.0 if objective == "time" else float(rng.choice([1.0, 4.0, 16.0, 64.0]))
        sliced, ssa, flops, size, _ = slice_and_reconfigure(inputs, size_dict, list(p.ssa_path), 2.0 ** width, subtree, 8, size_weight,
                                                            int(rng.integers(1 << 30)), 64, objective)
        if size > 2.0 ** width:
            continue
        path = ssa_replace_ordering(ContractionPath.simple([tuple(x) for x in ssa]))
        n_slices = 1
        for l in sliced:
            n_slices *= int(size_dict[l])
        t_model = path_time(meta, path, sliced) * n_slices
        if best is None or t_model < best["model_seconds"]:
            best = {"model_seconds": t_model, "sliced_legs": [int(l) for l in sliced], "n_slices": n_slices, "flops_mnk_per_slice": flops,
                    "peak_elements": size, "objective": objective, "size_weight": size_weight, "toplevel": [list(x) for x in path.toplevel]}
    return best


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--circuit", default="sycamore", choices=["sycamore", "random"])
    ap.add_argument("--qubits", type=int, default=53); ap.add_argument("--depth", type=int, default=12)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--width", type=float, default=31.0, help="log2 of the largest tensor allowed (elements)")
    ap.add_argument("--subtree", type=int, default=10)
    ap.add_argument("--trials", type=int, default=4); ap.add_argument("--workers", type=int, default=8)
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    t0 = time.time()
    with mp.get_context("spawn").Pool(a.workers) as pool:
        res = pool.map(worker, [(a.circuit, a.qubits, a.depth, a.seed, a.trials, 1000 + s, a.width, a.subtree) for s in range(a.workers)])
    res = [r for r in res if r]
    best = min(res, key=lambda r: r["model_seconds"])
    best.update({"network": f"{a.circuit} {a.qubits}q depth/rounds {a.depth} seed {a.seed}",
                 "finder": f"random-greedy + subtree reconfiguration (subtree {a.subtree}) + slicing to 2^{a.width:g}, best of {a.trials * a.workers} trials by the GPU time model (tools/search_path.py)"})
    print("best of %d: %d slices x %.3e MNK (total 8MNK %.3e), width 2^%.1f, model %.3f s on one GPU; all: %s; %.0f s" % (
        a.trials * a.workers, best["n_slices"], best["flops_mnk_per_slice"], 8 * best["flops_mnk_per_slice"] * best["n_slices"],
        math.log2(best["peak_elements"]), best["model_seconds"], sorted(round(r["model_seconds"], 3) for r in res), time.time() - t0))
    json.dump(best, open(a.out, "w"))
