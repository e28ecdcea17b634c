"""Partitioned contraction across GPUs: mirrors tnc::mpi::communication
(tnc/src/mpi/communication.rs) with one process per GPU.

  reference (MPI, host memory)                      here
  ------------------------------------------------  ---------------------------------------------
  broadcast_path / broadcast_serializing :32-69     torch.distributed object broadcast (metadata)
  get_tensor_mapping :89-115                        tncb_fanin_mapping (C ABI; FxHashMap walk order)
  scatter_tensor_network :125-195                   metadata scatter; leaves are uploaded by the
                                                    owning rank straight to its GPU
  send_tensor / receive_tensor :72-85 (postcard,    tncb_comm_send / tncb_comm_recv: raw complex128
    192-byte blobs, serialization.rs:43-79)         buffer over NCCL p2p (NVLink), no serialisation
  intermediate_reduce_tensor_network :199-249       same loop; the receiver contracts
                                                    [local, received] with path![(0,1)] on its GPU

torch.distributed is plumbing only (rendezvous, metadata, barriers); tensor payloads never
pass through it."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from .. import Context, DeviceTensor
from .._lib import check, lib, u64_array
from ..contractionpath import ContractionPath
from ..tensornetwork.tensor import Tensor
from ..tensornetwork.tensordata import TensorData


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_serializing(obj, root: int = 0, group=None):
    """communication.rs:52-69."""
    dist = _dist()
    box = [obj if dist.get_rank(group) == root else None]
    dist.broadcast_object_list(box, src=root, group=group)
    return box[0]


def broadcast_path(path, root: int = 0, group=None):
    """communication.rs:32-48: rank `root` sends its (simple) path, everyone returns it."""
    return broadcast_serializing(path, root, group)


def get_tensor_mapping(path: ContractionPath, size: int) -> Dict[int, int]:
    """partition index -> rank (communication.rs:89-115).  The partition on the left of the
    last top-level pair goes to rank 0, the others to 1, 2, ... in the reference's FxHashMap walk
    order (reproduced inside tncb_fanin_mapping; KAT communication.rs:257-279: 0->0, 2->1, 1->2)."""
    parts = sorted(path.nested)
    if not parts and not path.toplevel:
        return {0: 0}
    flat = [x for p in path.toplevel for x in p]
    ranks = (C.c_int * max(len(parts), 1))()
    check(lib().tncb_fanin_mapping(len(parts), u64_array(parts), len(path.toplevel), u64_array(flat), size, ranks))
    return {p: int(ranks[i]) for i, p in enumerate(parts)}


class RankTensorMapping:
    """mpi/mpi_types.rs:6-62: a bidirectional (1:1) mapping between ranks and composite tensors; every tensor maps to a rank,
    not every rank to a tensor."""

    def __init__(self):
        self._pairs: List[Tuple[int, int]] = []

    @classmethod
    def from_dict(cls, tensor_to_rank: Dict[int, int]) -> "RankTensorMapping":
        m = cls()
        for t, r in tensor_to_rank.items():
            m.insert(r, t)
        return m

    def insert(self, rank: int, tensor: int) -> None:
        assert self.tensor(rank) is None, f"Rank {rank} is already associated with a tensor"
        assert self._rank_opt(tensor) is None, f"Tensor {tensor} is already associated with a rank"
        self._pairs.append((int(rank), int(tensor)))

    def _rank_opt(self, tensor: int) -> Optional[int]:
        return next((r for r, t in self._pairs if t == tensor), None)

    def rank(self, tensor: int) -> int:
        r = self._rank_opt(tensor)
        assert r is not None, f"Tensor {tensor} has no rank"
        return r

    def tensor(self, rank: int) -> Optional[int]:
        return next((t for r, t in self._pairs if r == rank), None)

    def __len__(self) -> int:
        return len(self._pairs)

    def is_empty(self) -> bool:
        return not self._pairs

    def __iter__(self):
        return iter(self._pairs)


@dataclass
class Communication:
    """Opaque in the reference (communication.rs:118-120); also carries the external legs of
    every partition so that receivers know the shape of what arrives."""
    tensor_mapping: Dict[int, int] = field(default_factory=dict)          # partition -> rank
    external: Dict[int, Tuple[List[int], List[int]]] = field(default_factory=dict)  # partition -> (legs, dims)
    toplevel: Optional[list] = None     # the fan-in path, when it travelled with the scatter (one collective instead of three)

    def rank(self, partition: int) -> int:
        return self.tensor_mapping[partition]

    def tensor(self, rank: int) -> Optional[int]:
        for p, r in self.tensor_mapping.items():
            if r == rank:
                return p
        return None


def contracted_legs(tn: Tensor, path: ContractionPath) -> Tuple[List[int], List[int]]:
    """Leg order of `contract_tensor_network(tn, path)` from metadata alone: replay
    `b ^ a` (contraction.rs:64) along the path.  The reference does not need this because the
    serialised tensor carries its legs (serialization.rs:43-67); here only raw data travels."""
    if tn.is_leaf():
        return list(tn.legs), list(tn.bond_dims)
    # plain (legs, dims) lists instead of Tensor objects: this runs on rank 0 inside the timed scatter for every partition
    ts: List[Optional[Tuple[List[int], List[int]]]] = []
    for i, c in enumerate(tn.tensors):
        if c.is_composite() and i in path.nested:
            ts.append(contracted_legs(c, path.nested[i]))
        else:
            ts.append((c.legs, c.bond_dims) if c.is_leaf() else None)
    for (i, j) in path.toplevel:
        (al, ad), (bl, bd) = ts[i], ts[j]
        sa, sb = set(al), set(bl)
        keep_b = [q for q, l in enumerate(bl) if l not in sa]       # tensor.rs:463-479: (b \ a) ++ (a \ b), order-preserving
        keep_a = [q for q, l in enumerate(al) if l not in sb]
        ts[i] = ([bl[q] for q in keep_b] + [al[q] for q in keep_a], [bd[q] for q in keep_b] + [ad[q] for q in keep_a])
        ts[j] = None
    rest = [t for t in ts if t is not None]
    assert len(rest) == 1, "Not fully contracted"
    return list(rest[0][0]), list(rest[0][1])


def fanin_schedule(comm: Communication, toplevel) -> List[dict]:
    """The fan-in as a list of events every rank derives identically from metadata:
    {receiver, sender, recv_legs, recv_dims, out_legs, out_dims} per top-level pair."""
    ext = {p: (list(l), list(d)) for p, (l, d) in comm.external.items()}
    events = []
    for (x, y) in toplevel:
        (xl, xd), (yl, yd) = ext[x], ext[y]
        out_l = [l for l in yl if l not in xl] + [l for l in xl if l not in yl]   # (b \ a) ++ (a \ b), a = local
        dim = {**dict(zip(xl, xd)), **dict(zip(yl, yd))}
        events.append({"x": x, "y": y, "receiver": comm.rank(x), "sender": comm.rank(y),
                       "recv_legs": yl, "recv_dims": yd, "out_legs": out_l, "out_dims": [dim[l] for l in out_l]})
        ext[x] = (out_l, [dim[l] for l in out_l])
        ext.pop(y)
    return events


def init_device_comm(ctx: Context, group=None) -> None:
    """Creates the NCCL communicator inside libtncb200 (unique id travels as metadata)."""
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = None
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        check(ctx._l.tncb_comm_unique_id(buf))
        uid = bytes(buf)
    uid = broadcast_serializing(uid, 0, group)
    arr = (C.c_uint8 * 128).from_buffer_copy(uid)
    check(ctx._l.tncb_comm_init(ctx.handle, world, rank, arr))


def _pack_tensor(t: Tensor):
    """Tensor tree as nested tuples of plain lists: pickles in half the time and two thirds of the bytes of the object tree
    (489 Tensor + TensorData instances for the bench network)."""
    if t.tensors:
        return (0, [_pack_tensor(c) for c in t.tensors], t.legs, t.bond_dims)
    td = t.tensordata
    return (1, t.legs, t.bond_dims, td.kind, td.gate, td.matrix, td.file)


def _unpack_tensor(rec) -> Tensor:
    if rec[0] == 0:
        return Tensor(rec[2], rec[3], tensors=[_unpack_tensor(c) for c in rec[1]])
    _, legs, dims, kind, gate, matrix, file = rec
    return Tensor(legs, dims, tensordata=TensorData(kind=kind, gate=gate, matrix=matrix, file=file))


def _is_cpu_group(group) -> bool:
    try:
        return str(_dist().get_backend(group)).lower() == "gloo"
    except Exception:
        return False


def _scatter_blobs(blobs: Optional[List[bytes]], rank: int, size: int, group) -> bytes:
    """One byte string per rank from rank 0 in two plain collectives (an 8-byte broadcast of the longest length, one
    scatter of length-prefixed rows) -- torch's scatter_object_list needs two scatters plus its own pickling pass.
    CPU (gloo) groups only."""
    import torch
    dist = _dist()
    hdr = torch.zeros(1, dtype=torch.int64)
    if rank == 0:
        hdr[0] = max(len(b) for b in blobs)
    dist.broadcast(hdr, src=0, group=group)
    width = int(hdr[0]) + 8
    mine = torch.empty(width, dtype=torch.uint8)
    if rank == 0:
        full = np.zeros((size, width), dtype=np.uint8)
        for i, b in enumerate(blobs):
            full[i, :8] = np.frombuffer(len(b).to_bytes(8, "little"), dtype=np.uint8)
            full[i, 8:8 + len(b)] = np.frombuffer(b, dtype=np.uint8)
        rows = torch.from_numpy(full)
        dist.scatter(mine, [rows[i] for i in range(size)], src=0, group=group)
    else:
        dist.scatter(mine, None, src=0, group=group)
    raw = mine.numpy()
    n = int.from_bytes(raw[:8].tobytes(), "little")
    return raw[8:8 + n].tobytes()


def scatter_tensor_network(r_tn: Optional[Tensor], path: Optional[ContractionPath], rank: int, size: int, group=None):
    """communication.rs:125-195.  Rank 0 passes the partitioned network and its path; the others
    pass None.  Returns (local_tn, local_path, Communication); ranks without a partition get an
    empty Tensor and an empty path.  Over a CPU (gloo) group everything a rank needs -- its partition, its local path, the
    Communication and the fan-in path (`comm.toplevel`) -- travels in ONE scatter."""
    import pickle
    dist = _dist()
    if rank == 0:
        mapping = get_tensor_mapping(path, size)
        external = {}
        for p in mapping:
            # true leg order of the contracted partition (depends on its local path)
            external[p] = contracted_legs(r_tn.tensor(p), path.nested[p])
        comm = Communication(mapping, external)
        per_rank = [None] * size
        for p, r in mapping.items():
            per_rank[r] = (r_tn.tensor(p), path.nested[p])
    else:
        comm, per_rank = None, [None] * size
    if _is_cpu_group(group):
        blobs = None
        if rank == 0:
            comm.toplevel = [tuple(x) for x in path.toplevel]
            blobs = [pickle.dumps((comm, None if x is None else (_pack_tensor(x[0]), x[1])), protocol=pickle.HIGHEST_PROTOCOL) for x in per_rank]
        comm, part = pickle.loads(_scatter_blobs(blobs, rank, size, group))
        if part is None:
            return Tensor(), ContractionPath(), comm
        return _unpack_tensor(part[0]), part[1], comm
    comm = broadcast_serializing(comm, 0, group)
    out = [None]
    dist.scatter_object_list(out, per_rank if rank == 0 else None, src=0, group=group)
    if out[0] is None:
        return Tensor(), ContractionPath(), comm
    local_tn, local_path = out[0]
    return local_tn, local_path, comm


def _scatter_with_toplevel(r_tn, path, rank: int, size: int, group):
    """(toplevel, local_tn, local_path, comm): the broadcast_path + scatter_tensor_network recipe of
    tnc/examples/distributed_contraction.rs:43-60, as one collective when the group allows it."""
    if _is_cpu_group(group):
        local_tn, local_path, comm = scatter_tensor_network(r_tn, path, rank, size, group)
        return comm.toplevel, local_tn, local_path, comm
    toplevel = broadcast_path(path.toplevel if rank == 0 else None, 0, group)
    local_tn, local_path, comm = scatter_tensor_network(r_tn, path, rank, size, group)
    return toplevel, local_tn, local_path, comm


def _send(ctx: Context, t: Tensor, peer: int) -> None:
    dt = t.tensordata.matrix
    check(ctx._l.tncb_comm_send(ctx.handle, dt.handle, peer))


def _recv(ctx: Context, legs, dims, peer: int) -> Tensor:
    h = C.c_void_p()
    check(ctx._l.tncb_comm_recv(ctx.handle, len(dims), u64_array(dims), peer, C.byref(h)))
    t = Tensor(legs, dims)
    t.set_tensor_data(TensorData.Matrix(DeviceTensor.adopt(ctx, h)))
    return t


def intermediate_reduce_tensor_network(local_tn: Tensor, toplevel, rank: int, comm: Communication, ctx: Context, events=None) -> Tensor:
    """communication.rs:199-249: path-driven fan-in.  `local_tn` is this rank's contracted
    partition (a leaf with device data, or an empty Tensor).  Returns the local tensor after the
    fan-in; on rank 0 that is the final result.  `events`: a precomputed fanin_schedule(comm, toplevel)."""
    from ..tensornetwork.contraction import contract_tensor_network
    final_rank = 0
    if events is None:
        events = fanin_schedule(comm, toplevel)
    for ev in events:
        receiver, sender = ev["receiver"], ev["sender"]
        final_rank = receiver
        if receiver == rank:
            received = _recv(ctx, ev["recv_legs"], ev["recv_dims"], sender)
            tn = Tensor.new_composite([local_tn, received])
            local_tn = contract_tensor_network(tn, ContractionPath.single(0, 1), ctx=ctx)
            assert local_tn.legs == ev["out_legs"]
        if sender == rank:
            _send(ctx, local_tn, receiver)
    if final_rank != 0:
        legs, dims = None, None
        if rank == final_rank:
            _send(ctx, local_tn, 0)
        if rank == 0:
            # the final tensor's legs are the last event's output
            local_tn = _recv(ctx, events[-1]["out_legs"], events[-1]["out_dims"], final_rank)
    return local_tn


def contract_partitioned(r_tn: Optional[Tensor], path: Optional[ContractionPath], ctx: Context, group=None) -> Tensor:
    """The recipe of tnc/examples/distributed_contraction.rs:43-83 / benchmark/src/main.rs:369-399:
    broadcast the fan-in path, scatter, contract locally, fan in.  Result on rank 0."""
    from ..tensornetwork.contraction import contract_tensor_network
    dist = _dist()
    rank, size = dist.get_rank(group), dist.get_world_size(group)
    toplevel, local_tn, local_path, comm = _scatter_with_toplevel(r_tn, path, rank, size, group)
    if local_tn.is_composite():
        local_tn = contract_tensor_network(local_tn, local_path, ctx=ctx)
        mine = comm.tensor(rank)
        assert local_tn.legs == comm.external[mine][0], "fan-in metadata out of sync with the device result"
    return intermediate_reduce_tensor_network(local_tn, toplevel, rank, comm, ctx)


class PartitionedPlan:
    """Scatter once, run many: the partitioned contraction with this rank's partition compiled into a NetworkPlan
    whose leaves stay on the device (tncb_plan_stage).  `run()` = local contraction + NCCL fan-in with no host data
    movement -- the "inputs already resident in HBM" form of benchmark/src/main.rs:369-399; `contract_partitioned`
    is the end-to-end form (broadcast + scatter + leaf upload inside)."""

    def __init__(self, r_tn: Optional[Tensor], path: Optional[ContractionPath], ctx: Context, group=None):
        from ..tensornetwork.contraction import NetworkPlan
        dist = _dist()
        self.ctx = ctx
        self.rank, self.size = dist.get_rank(group), dist.get_world_size(group)
        self.toplevel, local_tn, local_path, self.comm = _scatter_with_toplevel(r_tn, path, self.rank, self.size, group)
        self.mine = self.comm.tensor(self.rank)
        self.plan = None
        if local_tn.is_composite():
            self.plan = NetworkPlan(local_tn, local_path, ctx=ctx)
            self.plan.stage(local_tn)
        self.pairs_local = len(local_path.toplevel) if local_tn.is_composite() else 0
        self.events = fanin_schedule(self.comm, self.toplevel)      # metadata only: derived once, identical on every rank

    def run(self) -> Tensor:
        local = self.plan.run() if self.plan is not None else Tensor()
        if self.plan is not None:
            assert local.legs == self.comm.external[self.mine][0], "fan-in metadata out of sync with the device result"
        return intermediate_reduce_tensor_network(local, self.toplevel, self.rank, self.comm, self.ctx, self.events)
