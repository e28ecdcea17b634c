"""Worker for the world_size>1 tests (spawned by tests/test_dist.py and by torchrun on the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build_case(qubits=10, rounds=6, seed=22, parts=2):
    from tnc_b200.builders import random_circuit
    from tnc_b200.contractionpath.paths import Cotengrust
    from tnc_b200.tensornetwork.partitioning import find_partitioning, partition_tensor_network
    tn = random_circuit(qubits, rounds, 0.5, 0.5, np.random.default_rng(seed))
    part = find_partitioning(tn, parts, seed=1)
    sa_steps = int(os.environ.get("TNCB_SA", "0"))
    if sa_steps:   # refine like the reference's SA balancer (step-budget, seeded)
        from tnc_b200.contractionpath.repartitioning import balance_partitions
        part, _ = balance_partitions(tn, parts, part, steps=sa_steps, seed=1)
    ptn = partition_tensor_network(tn, part)
    opt = Cotengrust(ptn); opt.find_path()
    flat = Cotengrust(tn); flat.find_path()
    return tn, flat.get_best_replace_path(), ptn, opt.get_best_replace_path()


def cpu_logic(rank, world, port, q):
    """gloo, no GPU: metadata paths only (broadcast, mapping, scatter, fan-in schedule)."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tnc_b200.contractionpath import path
        from tnc_b200.dist import broadcast_path, fanin_schedule, get_tensor_mapping, scatter_tensor_network
        # integration_tests.rs:85-116 test_broadcast_contraction_path
        ref = [(0, 1), (0, 2), (3, 4), (0, 3)]
        got = broadcast_path(ref if rank == 0 else None, 0)
        assert got == ref
        tn, fpath, ptn, ppath = build_case(parts=world) if rank == 0 else (None, None, None, None)
        local_tn, local_path, comm = scatter_tensor_network(ptn, ppath, rank, world)
        toplevel = broadcast_path(ppath.toplevel if rank == 0 else None, 0)
        # over a gloo group the fan-in path rides along with the scatter (one collective): same content
        assert [tuple(x) for x in comm.toplevel] == [tuple(x) for x in toplevel]
        from tnc_b200.dist.communication import _scatter_with_toplevel
        top2, tn2, path2, comm2 = _scatter_with_toplevel(ptn, ppath, rank, world, None)
        assert [tuple(x) for x in top2] == [tuple(x) for x in toplevel] and comm2 == comm and path2 == local_path
        assert [(t.legs, t.bond_dims, t.tensordata.kind, t.tensordata.gate) for t in tn2.tensors] == \
               [(t.legs, t.bond_dims, t.tensordata.kind, t.tensordata.gate) for t in local_tn.tensors]
        if rank == 0:   # the scattered partition is the partition rank 0 holds
            orig = ptn.tensor(comm.tensor(0))
            assert [(t.legs, t.bond_dims, t.tensordata.kind, t.tensordata.gate) for t in orig.tensors] == \
                   [(t.legs, t.bond_dims, t.tensordata.kind, t.tensordata.gate) for t in local_tn.tensors]
        mine = comm.tensor(rank)
        assert mine is not None and local_tn.is_composite()
        assert len(local_path.toplevel) == len(local_tn.tensors) - 1
        ev = fanin_schedule(comm, toplevel)
        assert len(ev) == world - 1 and ev[-1]["receiver"] == 0
        assert ev[-1]["out_legs"] == []          # amplitude network: scalar at the end
        # the broadcast leg orders are the true orders of the contracted partitions
        from tnc_b200.dist.communication import contracted_legs
        assert comm.external[mine] == contracted_legs(local_tn, local_path)
        assert sorted(comm.external[mine][0]) == sorted(local_tn.external_tensor().legs)
        summary = (sorted(comm.tensor_mapping.items()), [(e["sender"], e["receiver"], tuple(e["recv_dims"])) for e in ev],
                   len(local_tn.tensors))
        q.put((rank, summary))
    finally:
        dist.destroy_process_group()


def cpu_fanin(rank, world, port, q):
    """gloo, no GPU: the WHOLE partitioned path of tnc_b200.dist (scatter, local contraction, path-driven fan-in, final hop;
    contract_partitioned and PartitionedPlan) with the device engine swapped for the oracle and NCCL p2p for gloo
    send / recv of host buffers.  What is under test is the host logic the GPU run depends on: which rank holds what, who
    sends to whom in which order, the leg order a receiver assumes for a raw buffer, `[local, received]` as pair (0, 1)."""
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import tnc_oracle as orc
        import tnc_b200.dist.communication as comm_mod
        import tnc_b200.tensornetwork.contraction as con_mod
        from tnc_b200.tensornetwork import Tensor, TensorData

        def to_o(t):
            if t.is_composite():
                return orc.OTensor(children=[to_o(c) for c in t.tensors])
            td = t.tensordata
            d = ("gate", td.gate[0], td.gate[1], td.gate[2]) if td.kind == "gate" else (None if td.kind == "uncontracted" else np.asarray(td.matrix))
            return orc.OTensor(list(t.legs), list(t.bond_dims), d)

        def to_op(p):
            return orc.OPath(list(p.toplevel), {i: to_op(x) for i, x in p.nested.items()})

        def fake_contract(tn, path, ctx=None):
            r = orc.contract_tensor_network(to_o(tn), to_op(path))
            out = Tensor(r.legs, r.dims)
            out.set_tensor_data(TensorData.Matrix(np.ascontiguousarray(r.data).reshape(r.dims)))
            return out

        def fake_send(ctx, t, peer):
            buf = np.ascontiguousarray(np.asarray(t.tensordata.matrix, dtype=np.complex128)).reshape(-1)
            dist.send(torch.from_numpy(buf.view(np.float64).copy()), dst=peer)            # the raw buffer only, like ncclSend

        def fake_recv(ctx, legs, dims, peer):
            n = int(np.prod(dims, dtype=np.int64)) if len(dims) else 1
            buf = torch.empty(2 * n, dtype=torch.float64)
            dist.recv(buf, src=peer)
            t = Tensor(legs, dims)
            t.set_tensor_data(TensorData.Matrix(buf.numpy().view(np.complex128).reshape(dims)))
            return t

        class FakePlan:
            def __init__(self, tn, path, ctx=None):
                self.tn, self.path = tn, path

            def stage(self, tn):
                self.tn = tn

            def run(self):
                return fake_contract(self.tn, self.path)

        con_mod.contract_tensor_network = fake_contract
        con_mod.NetworkPlan = FakePlan
        comm_mod._send, comm_mod._recv = fake_send, fake_recv
        tn, fpath, ptn, ppath = build_case(12, 6, 7, world) if rank == 0 else (None, None, None, None)
        res = comm_mod.contract_partitioned(ptn, ppath, None)
        plan = comm_mod.PartitionedPlan(ptn, ppath, None)
        res2, res3 = plan.run(), plan.run()
        if rank == 0:
            flat = complex(orc.contract_tensor_network(to_o(tn), to_op(fpath)).data)
            one = complex(orc.contract_tensor_network(to_o(ptn), to_op(ppath)).data)      # the nested path on one process
            for r in (res, res2, res3):
                assert r.legs == []
                amp = complex(np.asarray(r.tensordata.matrix))
                assert abs(amp - flat) <= 1e-10 * abs(flat) + 1e-14 and abs(amp - one) <= 1e-12 * abs(one) + 1e-14, (amp, flat, one)
            q.put((rank, ("ok", len(plan.events))))
        else:
            q.put((rank, ("ok", len(plan.events))))
    finally:
        dist.destroy_process_group()


def gpu_main():
    """torchrun entry on the GPU box: partitioned contraction over NCCL == flat contraction."""
    import time
    import torch
    import torch.distributed as dist
    import tnc_b200 as tb
    from tnc_b200.dist import contract_partitioned, init_device_comm
    from tnc_b200.tensornetwork import contract_tensor_network
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = tb.Context(local)
    init_device_comm(ctx)
    q, r = int(os.environ.get("TNCB_Q", "20")), int(os.environ.get("TNCB_R", "8"))
    seed = int(os.environ.get("TNCB_SEED", "5"))
    tn, fpath, ptn, ppath = build_case(q, r, seed, world) if rank == 0 else (None, None, None, None)
    for it in range(3):
        dist.barrier(); ctx.synchronize()
        t0 = time.perf_counter()
        res = contract_partitioned(ptn, ppath, ctx)
        if rank == 0:
            amp = complex(res.to_numpy())
        ctx.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
    # slicing over ranks with one NCCL all-reduce
    from tnc_b200.contractionpath.slicing import contract_sliced, find_slices
    from tnc_b200.dist import broadcast_serializing
    spec = None
    if rank == 0:
        spec = (tn, fpath, find_slices(tn, fpath, min_slices=max(4, world)))
    stn, spath, legs = broadcast_serializing(spec, 0)
    dist.barrier(); ctx.synchronize(); t0 = time.perf_counter()
    sres = contract_sliced(stn, spath, legs, ctx=ctx, rank=rank, world=world)
    samp = complex(sres.to_numpy()); ts = time.perf_counter() - t0
    if rank == 0:
        t0 = time.perf_counter(); flat = complex(contract_tensor_network(tn, fpath, ctx=ctx).to_numpy()); tf = time.perf_counter() - t0
        print(f"SLICED_OK world={world} slices={2 ** len(legs)} sliced={samp} abs_err={abs(samp - flat):.3e} t_sliced={ts*1e3:.2f}ms", flush=True)
        assert abs(samp - flat) <= 1e-9 * abs(flat) + 1e-14
        err = abs(amp - flat)
        print(f"DIST_OK world={world} q={q} r={r} partitioned={amp} flat={flat} abs_err={err:.3e} t_part={dt*1e3:.2f}ms t_flat={tf*1e3:.2f}ms", flush=True)
        assert err <= 1e-9 * abs(flat) + 1e-14
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    try:
        gpu_main()
    except BaseException:  # fail fast: never leave the other ranks waiting in a collective
        import traceback
        traceback.print_exc()
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(1)
