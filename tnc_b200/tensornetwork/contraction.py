"""contract_tensor_network (tnc/src/tensornetwork/contraction.rs:30-52) over the C ABI.

The Python side only marshals the `Tensor` tree and the `ContractionPath` into the plain C
structs of include/tncb.h; schedule construction, leaf materialisation, the single host->device
upload and every pair kernel run inside libtncb200."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from .. import Context, DeviceTensor, default_context
from .._lib import TncbError, TncbPath, TncbTn, check, u64_array
from ..contractionpath import ContractionPath
from .tensor import Tensor
from .tensordata import TensorData

_KIND = {"uncontracted": 0, "matrix": 1, "gate": 2, "device": 3}


class _Marshal:
    """Keeps every ctypes buffer alive for the duration of the call."""

    def __init__(self):
        self.keep: List[object] = []
        self.device_inputs: List[DeviceTensor] = []

    def tn(self, t: Tensor) -> TncbTn:
        node = TncbTn()
        if t.is_composite():
            arr = (TncbTn * len(t.tensors))(*[self.tn(c) for c in t.tensors])
            self.keep.append(arr)
            node.n_children = len(t.tensors)
            node.children = arr
            node.kind = 0
            return node
        legs, dims = u64_array(t.legs), u64_array(t.bond_dims)
        self.keep += [legs, dims]
        node.n_children = 0
        node.rank = len(t.legs)
        node.legs, node.dims = legs, dims
        td = t.tensordata
        if td.kind == "gate":
            name, angles, adj = td.gate
            ang = (C.c_double * max(len(angles), 1))(*angles)
            nm = C.c_char_p(name.encode())
            self.keep += [ang, nm]
            node.kind = 2
            node.gate_name = nm
            node.gate_angles = ang
            node.n_gate_angles = len(angles)
            node.gate_adjoint = int(adj)
        elif td.kind == "matrix":
            m = td.matrix
            if isinstance(m, DeviceTensor):
                if m.handle is None:   # consumed by an earlier call (the Rust move left TensorData::Uncontracted behind)
                    raise TncbError(-3, "Cannot convert uncontracted tensor to data (device tensor already consumed)")
                node.kind = 3
                node.device = m.handle
                self.device_inputs.append(m)
            else:
                a = np.asarray(m, dtype=np.complex128, order="C")
                if list(a.shape) != list(t.bond_dims):
                    a = a.reshape(t.bond_dims)
                self.keep.append(a)
                node.kind = 1
                node.host_re_im = a.ctypes.data_as(C.POINTER(C.c_double))
        elif td.kind == "file":
            node.kind = 99  # TensorData::File needs HDF5 -> TNCB_ERR_UNSUPPORTED
        else:
            node.kind = 0
        return node

    def path(self, p: ContractionPath) -> TncbPath:
        out = TncbPath()
        flat = [x for pair in p.toplevel for x in pair]
        pairs = u64_array(flat)
        self.keep.append(pairs)
        out.n_pairs = len(p.toplevel)
        out.pairs = pairs
        idx = sorted(p.nested)
        if idx:
            ni = u64_array(idx)
            arr = (TncbPath * len(idx))(*[self.path(p.nested[i]) for i in idx])
            self.keep += [ni, arr]
            out.n_nested = len(idx)
            out.nested_index = ni
            out.nested = arr
        return out


def contract_tensor_network(tn: Tensor, contract_path: ContractionPath, ctx: Optional[Context] = None) -> Tensor:
    """Fully contracts `tn` with the replace-left `contract_path`; returns the resulting
    leaf `Tensor` whose data stays on the device (`.to_numpy()` downloads it)."""
    ctx = ctx or default_context()
    m = _Marshal()
    c_tn = m.tn(tn)
    c_path = m.path(contract_path)
    out = C.c_void_p()
    n_out = C.c_int()
    legs = u64_array([0] * 64)
    rc = ctx._l.tncb_contract_tensor_network(ctx.handle, C.byref(c_tn), C.byref(c_path), C.byref(out), C.byref(n_out), legs)
    check(rc)
    for d in m.device_inputs:  # consumed by the call
        d.release()
    if not out.value:
        return Tensor()  # nothing left (empty network)
    dt = DeviceTensor.adopt(ctx, out)
    res = Tensor([legs[i] for i in range(n_out.value)], dt.shape)
    res.set_tensor_data(TensorData.Matrix(dt))
    return res


class NetworkPlan:
    """Compile once / execute many (tncb_plan_*): same structure, new payloads."""

    def __init__(self, tn: Tensor, contract_path: ContractionPath, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        m = _Marshal()
        c_tn, c_path = m.tn(tn), m.path(contract_path)
        h = C.c_void_p()
        check(self.ctx._l.tncb_plan_create(self.ctx.handle, C.byref(c_tn), C.byref(c_path), C.byref(h)))
        self.handle = h

    def info(self) -> dict:
        n, k = C.c_uint64(), C.c_uint64()
        pk = C.c_uint64()
        fl, by = C.c_double(), C.c_double()
        check(self.ctx._l.tncb_plan_info(self.handle, C.byref(n), C.byref(fl), C.byref(by), C.byref(pk), C.byref(k)))
        return {"pairs": n.value, "flops": fl.value, "bytes": by.value, "peak_bytes": pk.value, "kernels": k.value}

    def stage(self, tn: Tensor) -> None:
        """Materialise + upload the leaves once (tncb_plan_stage); `run()` then needs no host data."""
        m = _Marshal()
        c_tn = m.tn(tn)
        check(self.ctx._l.tncb_plan_stage(self.ctx.handle, self.handle, C.byref(c_tn)))

    def run(self) -> Tensor:
        out, n_out, legs = C.c_void_p(), C.c_int(), u64_array([0] * 64)
        check(self.ctx._l.tncb_plan_run(self.ctx.handle, self.handle, C.byref(out), C.byref(n_out), legs))
        if not out.value:
            return Tensor()
        dt = DeviceTensor.adopt(self.ctx, out)
        res = Tensor([legs[i] for i in range(n_out.value)], dt.shape)
        res.set_tensor_data(TensorData.Matrix(dt))
        return res

    def stage_slices(self, slice_tns) -> None:
        """Materialise + upload the leaf blocks of every slice network once (tncb_plan_stage_slices)."""
        m = _Marshal()
        nodes = [m.tn(t) for t in slice_tns]
        ptrs = (C.POINTER(TncbTn) * len(nodes))(*[C.pointer(n) for n in nodes])
        check(self.ctx._l.tncb_plan_stage_slices(self.ctx.handle, self.handle, len(nodes), ptrs))

    def run_slices(self, first: int = 0, stride: int = 1) -> Tensor:
        """Sum of the slices first, first + stride, ... on the device, no host work per slice."""
        out, n_out, legs = C.c_void_p(), C.c_int(), u64_array([0] * 64)
        check(self.ctx._l.tncb_plan_run_slices(self.ctx.handle, self.handle, int(first), int(stride), C.byref(out), C.byref(n_out), legs))
        dt = DeviceTensor.adopt(self.ctx, out)
        res = Tensor([legs[i] for i in range(n_out.value)], dt.shape)
        res.set_tensor_data(TensorData.Matrix(dt))
        return res

    def execute(self, tn: Tensor) -> Tensor:
        m = _Marshal()
        c_tn = m.tn(tn)
        out, n_out, legs = C.c_void_p(), C.c_int(), u64_array([0] * 64)
        check(self.ctx._l.tncb_plan_execute(self.ctx.handle, self.handle, C.byref(c_tn), C.byref(out), C.byref(n_out), legs))
        for d in m.device_inputs:
            d.release()
        if not out.value:
            return Tensor()
        dt = DeviceTensor.adopt(self.ctx, out)
        res = Tensor([legs[i] for i in range(n_out.value)], dt.shape)
        res.set_tensor_data(TensorData.Matrix(dt))
        return res

    def __del__(self):
        try:
            if self.handle:
                self.ctx._l.tncb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass
