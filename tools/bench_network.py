"""Times contract_tensor_network on a random-circuit amplitude network (BASELINE configs 3/4)
through the public API, with the oracle (torch-CPU MKL) beside it.  Timed region mirrors
benchmark/src/main.rs:355-360: path finding excluded, leaf materialisation + H2D + final D2H included."""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=24); ap.add_argument("--rounds", type=int, default=12)
    ap.add_argument("--seed", type=int, default=1); ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--cpu", action="store_true"); ap.add_argument("--plan", action="store_true")
    ap.add_argument("--resident", action="store_true", help="plan.stage() once, then plan.run(): no host data movement per run")
    ap.add_argument("--circuit", default="random", choices=["random", "sycamore"])
    ap.add_argument("--trials", type=int, default=0, help="random-greedy trials (0 = plain greedy)")
    ap.add_argument("--path-file", default="", help="cache the replace-left path as JSON (path finding is not timed)")
    a = ap.parse_args()
    import tnc_b200 as tb
    from tnc_b200.builders import random_circuit, sycamore_circuit
    from tnc_b200.contractionpath import ContractionPath
    from tnc_b200.contractionpath.paths import Cotengrust, OptMethod
    from tnc_b200.tensornetwork import contract_tensor_network, NetworkPlan
    if a.circuit == "sycamore":
        tn = sycamore_circuit(a.qubits, a.rounds, np.random.default_rng(a.seed)).into_amplitude_network("0" * a.qubits)[0]
    else:
        tn = random_circuit(a.qubits, a.rounds, 0.5, 0.5, np.random.default_rng(a.seed))
    if a.path_file and os.path.exists(a.path_file):
        path = ContractionPath.simple([tuple(x) for x in json.load(open(a.path_file))["toplevel"]])
    else:
        opt = Cotengrust(tn, OptMethod.RandomGreedy, a.trials) if a.trials else Cotengrust(tn)
        opt.find_path(); path = opt.get_best_replace_path()
        if a.path_file:
            json.dump({"network": f"{a.circuit} {a.qubits}q depth/rounds {a.rounds} seed {a.seed}", "finder": f"random-greedy {a.trials} trials" if a.trials else "greedy",
                       "toplevel": path.toplevel}, open(a.path_file, "w"))
    ctx = tb.Context(0)
    plan = NetworkPlan(tn, path, ctx=ctx)
    info = plan.info()
    run = (lambda: plan.execute(tn)) if a.plan else (lambda: contract_tensor_network(tn, path, ctx=ctx))
    if a.resident:
        plan.stage(tn)
        run = lambda: plan.run()
    amp = complex(run().to_numpy())
    ctx.reset_stats()
    ts = []
    for _ in range(a.steps):
        t0 = time.perf_counter(); r = run(); v = complex(r.to_numpy()); ts.append(time.perf_counter() - t0)
    st = ctx.stats()
    sec = float(np.median(ts))
    out = {"network": f"{a.circuit} {a.qubits}q {a.rounds}r seed{a.seed}", "pairs": info["pairs"], "flops": info["flops"],
           "peak_GiB": info["peak_bytes"] / 2**30, "gpu_ms": sec * 1e3, "gpu_ms_all": [round(t * 1e3, 3) for t in ts],
           "pairs_per_s": info["pairs"] / sec, "tflops": info["flops"] / sec * 1e-12, "amp": [amp.real, amp.imag],
           "launches_per_run": st["kernel_launches"] / a.steps, "arena_peak_GiB": st["arena_peak_bytes"] / 2**30, "plan_reuse": a.plan, "resident": a.resident}
    if a.cpu:
        import torch
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
        from test_gpu_networks import to_oracle, to_opath
        from oracle import tnc_oracle as orc
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from bench import effective_cpus
        torch.set_num_threads(effective_cpus())
        otn, op = to_oracle(tn), to_opath(path)
        t0 = time.perf_counter(); ref = orc.contract_tensor_network(otn, op, backend="torch"); cpu = time.perf_counter() - t0
        t0 = time.perf_counter(); ref = orc.contract_tensor_network(otn, op, backend="torch"); cpu = min(cpu, time.perf_counter() - t0)
        rv = complex(ref.data)
        out.update({"cpu_ms": cpu * 1e3, "cpu_cores": effective_cpus(), "cpu_tflops": info["flops"] / cpu * 1e-12,
                    "speedup": cpu / sec, "rel_err": abs(rv - amp) / abs(rv)})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
