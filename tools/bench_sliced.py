"""BASELINE config 5: a big single-amplitude network (Sycamore-53) contracted as 2^s independent slices of ONE replace-left
path (tools/search_path.py), slices round-robin over the ranks, one ncclAllReduce of the scalar at the end
(contractionpath/slicing.py SlicedPlan -> tncb_plan_stage_slices / tncb_plan_run_slices).

  python tools/bench_sliced.py --path-file bench_inputs/sycamore53_d12.json                      # 1 GPU
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         tools/bench_sliced.py --path-file bench_inputs/sycamore53_d12.json                      # 8 GPUs

Timed region per step: every slice of this rank through the compiled plan (leaves resident) + the all-reduce + D2H of the
amplitude, wall clock between barriers, max over ranks.  --cpu-slices S times S slices of the same path on the CPU
oracle (torch MKL, all host threads; rank 0 only) and extrapolates -- the bounded CPU sample the brief asks for.
--check-file compares the amplitude with another path file's amplitude (two independent paths / slicings of the same network)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(desc):
    from tnc_b200.builders import random_circuit, sycamore_circuit
    w = desc.split()          # "<kind> <Q>q depth/rounds <D> seed <S>"
    kind, qubits, depth, seed = w[0], int(w[1][:-1]), int(w[3]), int(w[5])
    if kind == "sycamore":
        return sycamore_circuit(qubits, depth, np.random.default_rng(seed)).into_amplitude_network("0" * qubits)[0]
    return random_circuit(qubits, depth, 0.5, 0.5, np.random.default_rng(seed))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--path-file", required=True); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--cpu-slices", type=int, default=0); ap.add_argument("--check-file", default="")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import tnc_b200 as tb
    from tnc_b200.contractionpath import ContractionPath
    from tnc_b200.contractionpath.slicing import SlicedPlan, path_cost
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    meta_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        meta_group = dist.new_group(backend="gloo")
    ctx = tb.Context(local)
    if world > 1:
        from tnc_b200.dist import init_device_comm
        init_device_comm(ctx, meta_group)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def run_file(pf, steps, warmup):
        d = json.load(open(pf))
        tn = build(d["network"])
        path = ContractionPath.simple([tuple(x) for x in d["toplevel"]])
        legs = d["sliced_legs"]
        meta = [(t.legs, t.bond_dims) for t in tn.tensors]
        flops_slice, peak, _ = path_cost(meta, path, legs)
        t0 = time.perf_counter()
        sp = SlicedPlan(tn, path, legs, ctx=ctx)
        setup = time.perf_counter() - t0
        ts, amp = [], None
        ctx.reset_stats()
        for it in range(warmup + steps):
            if world > 1:
                dist.barrier()
            ctx.synchronize()
            t0 = time.perf_counter()
            amp = complex(sp.run(rank, world).to_numpy())
            dt = max_over_ranks(time.perf_counter() - t0)
            if it >= warmup:
                ts.append(dt)
        st = ctx.stats()
        sec = float(np.median(ts))
        pairs = len(path.toplevel) * sp.n_slices
        return d, tn, path, legs, {"network": d["network"], "finder": d.get("finder", ""), "n_gpus": world, "slices": sp.n_slices, "pairs_per_slice": len(path.toplevel),
                                   "flops_8mnk_total": flops_slice * sp.n_slices, "peak_tensor_GiB": peak * 16 / 2 ** 30,
                                   "seconds": sec, "seconds_all": [round(t, 4) for t in ts], "setup_seconds_untimed": setup,
                                   "pairs_per_s": pairs / sec, "tflops": flops_slice * sp.n_slices / sec * 1e-12,
                                   "amplitude": [amp.real, amp.imag], "arena_peak_GiB": st["arena_peak_bytes"] / 2 ** 30,
                                   "engine_counts": ctx.engine_counts(), "model_seconds_1gpu": d.get("model_seconds")}

    d, tn, path, legs, out = run_file(a.path_file, a.steps, a.warmup)
    if a.check_file:
        _, _, _, _, other = run_file(a.check_file, 1, 0)
        a0, a1 = complex(*out["amplitude"]), complex(*other["amplitude"])
        out["check"] = {"other_path": a.check_file, "other_slices": other["slices"], "other_seconds": other["seconds"], "other_amplitude": other["amplitude"],
                        "rel_diff": abs(a0 - a1) / abs(a0)}
    if a.cpu_slices and rank == 0:
        import torch as _t
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_gpu_networks import to_oracle, to_opath
        from oracle import tnc_oracle as orc
        from bench import effective_cpus
        from tnc_b200.contractionpath.slicing import SlicedNetwork
        _t.set_num_threads(effective_cpus())
        sn = SlicedNetwork(tn, legs)
        t0 = time.perf_counter(); acc = 0j
        for s in range(a.cpu_slices):
            acc += complex(orc.contract_tensor_network(to_oracle(sn.slice(sn.assignments[s])), to_opath(path), backend="torch").data)
        cpu = time.perf_counter() - t0
        out["cpu_baseline"] = {"kind": "port", "cores": effective_cpus(), "sample": f"{a.cpu_slices} of {out['slices']} slices of the same path (oracle, torch-CPU MKL)",
                               "seconds_sample": cpu, "seconds_extrapolated": cpu * out["slices"] / a.cpu_slices,
                               "speedup_extrapolated": cpu * out["slices"] / a.cpu_slices / out["seconds"]}
    if rank == 0:
        print(json.dumps(out))
        if a.out:
            with open(a.out, "a") as f:
                f.write(json.dumps(out) + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
