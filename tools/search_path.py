"""Offline path search (planning, CPU only): parallel random-greedy trials for a Sycamore-53 amplitude network, best path
(by flops, ties by width) written as a replace-left path JSON that tools/bench_network.py --path-file consumes.
usage: python tools/search_path.py DEPTH TRIALS_PER_WORKER WORKERS OUT.json"""
import json
import math
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(depth):
    from tnc_b200.builders import sycamore_circuit
    return sycamore_circuit(53, depth, np.random.default_rng(1)).into_amplitude_network("0" * 53)[0]


def worker(args):
    depth, trials, seed = args
    from tnc_b200.contractionpath import ContractionPath, ssa_replace_ordering
    from tnc_b200.contractionpath.paths.cotengrust import _Processor, _ssa_path_cost
    tn = build(depth)
    inputs = [list(t.legs) for t in tn.tensors]
    size_dict = {l: float(d) for t in tn.tensors for l, d in t.edges()}
    rng = np.random.default_rng(seed)
    best = None
    for trial in range(trials):
        p = _Processor(inputs, [], size_dict)
        costmod = float(rng.uniform(0.0, 50.0)) or 1e-3
        temp = float(np.exp(rng.uniform(np.log(1e-3), np.log(1.0))))
        p.optimize_greedy(max(costmod, 1e-3), temp, rng)
        p.optimize_remaining_by_size()
        f = _ssa_path_cost(inputs, [], size_dict, p.ssa_path, "flops")
        w = _ssa_path_cost(inputs, [], size_dict, p.ssa_path, "size")
        key = (f, w)
        if best is None or key < best[0]:
            best = (key, list(p.ssa_path), costmod, temp)
    return best


if __name__ == "__main__":
    depth, trials, workers, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    t0 = time.time()
    with mp.get_context("spawn").Pool(workers) as pool:
        res = pool.map(worker, [(depth, trials, 1000 + s) for s in range(workers)])
    best = min(res, key=lambda r: r[0])
    from tnc_b200.contractionpath import ContractionPath, ssa_replace_ordering
    path = ssa_replace_ordering(ContractionPath.simple([tuple(x) for x in best[1]]))
    (f, w) = best[0]
    print("best of %d trials: 8MNK-ish flops %.3e, width 2^%.1f (costmod %.2f, temperature %.4f), %.0f s" % (trials * workers, 8 * f, math.log2(w), best[2], best[3], time.time() - t0))
    json.dump({"network": f"sycamore 53q depth/rounds {depth} seed 1", "finder": f"random-greedy {trials * workers} trials (tools/search_path.py)",
               "flops_mnk": f, "peak_elements": w, "toplevel": path.toplevel}, open(out, "w"))
