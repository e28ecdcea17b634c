"""Import and export of tensors / tensor networks as HDF5 files (tnc/src/io/hdf5.rs).

File structure (hdf5.rs:1-15): one group ``tensors``; every member is an n-dimensional complex dataset whose shape is the
tensor's bond dimensions, with an integer attribute ``bids`` (its bond ids); the member called ``-1`` is the output
tensor: it carries the open bonds of the network in ``bids`` and no data.

The reference binds libhdf5 through the crate hdf5-metno; here the format is read and written by libtncb200 itself
(csrc/hdf5io.cpp behind ``tncb_hdf5_*``), host only.  ``TensorData.File`` leaves are loaded by the library while it
stages the leaves of a network (``into_data``, tensordata.rs:43-49)."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .._lib import check, lib, u64_array
from ..tensornetwork.tensor import Tensor
from ..tensornetwork.tensordata import TensorData

__all__ = ["load_tensor", "load_data", "store_data", "store_tensor", "Hdf5File"]


class Hdf5File:
    """An opened file: the members of one group (default ``/tensors``) in ascending name order, the order
    ``Group::member_names`` returns (hdf5.rs:56,92)."""

    def __init__(self, filename, group: Optional[str] = None):
        self._l = lib()
        h = C.c_void_p()
        check(self._l.tncb_hdf5_open(os.fsencode(filename), None if group is None else group.encode(), C.byref(h)))
        self._h = h

    def close(self) -> None:
        if self._h is not None:
            self._l.tncb_hdf5_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def member_names(self) -> List[str]:
        n = self._l.tncb_hdf5_count(self._h)
        return [self._l.tncb_hdf5_name(self._h, i).decode("utf-8", "surrogateescape") for i in range(n)]

    def shape(self, i: int) -> List[int]:
        rank = C.c_int()
        dims = u64_array([0] * 32)
        check(self._l.tncb_hdf5_shape(self._h, i, C.byref(rank), dims, None))
        return [int(dims[q]) for q in range(rank.value)]

    def attr(self, i: int, name: str) -> List[int]:
        n = C.c_size_t()
        check(self._l.tncb_hdf5_attr(self._h, i, name.encode(), 0, None, C.byref(n)))
        out = (C.c_int64 * max(n.value, 1))()
        check(self._l.tncb_hdf5_attr(self._h, i, name.encode(), n.value, out, C.byref(n)))
        return [int(out[q]) for q in range(n.value)]

    def read(self, i: int) -> np.ndarray:
        shape = self.shape(i)
        elems = C.c_uint64()
        check(self._l.tncb_hdf5_shape(self._h, i, None, None, C.byref(elems)))
        if elems.value == 0 and not shape:             # null dataspace ("-1" created without data): nothing to read
            return np.zeros((0,), dtype=np.complex128)
        out = np.empty(shape, dtype=np.complex128)
        check(self._l.tncb_hdf5_read(self._h, i, out.ctypes.data))
        return out


def _bond_ids(f: Hdf5File, i: int) -> List[int]:
    bids = f.attr(i, "bids")
    if any(b < 0 for b in bids):                       # read_1d::<usize> fails on negative ids
        raise ValueError("negative bond id in '%s'" % f.member_names()[i])
    return bids


def load_tensor(filename) -> Tensor:
    """Loads a tensor network (hdf5.rs:28-34, read_tensor :54-88): a composite of one Matrix leaf per member other than
    ``-1``, in member order; the composite's legs are the ``bids`` of ``-1``."""
    with Hdf5File(filename) as f:
        names = f.member_names()
        if "-1" not in names:
            raise KeyError("'%s' has no output tensor '-1' in /tensors" % filename)
        out_bond_ids = _bond_ids(f, names.index("-1"))
        tn = Tensor()
        for i, name in enumerate(names):
            if name == "-1":
                continue
            bond_ids = _bond_ids(f, i)
            data = f.read(i)
            t = Tensor.new(bond_ids, list(data.shape))
            t.set_tensor_data(TensorData.Matrix(data))
            tn.push_tensor(t)
        tn.legs = [int(b) for b in out_bond_ids]       # set_legs: bond dims of the composite are not touched (tensor.rs)
        return tn


def load_data(filename) -> np.ndarray:
    """Loads a single tensor (hdf5.rs:37-43, read_data :90-103): the first member of /tensors."""
    with Hdf5File(filename) as f:
        if not f.member_names():
            raise IndexError("'%s' has no member in /tensors" % filename)
        return f.read(0)


def store_data(filename, tensor) -> None:
    """Stores a single tensor as /tensors/-1 (hdf5.rs:46-52, write_data :105-113)."""
    arr = np.asarray(tensor, dtype=np.complex128, order="C")       # (ascontiguousarray would turn a scalar into shape (1,))
    dims = u64_array(arr.shape)
    check(lib().tncb_hdf5_store_data(os.fsencode(filename), arr.ndim, dims, arr.ctypes.data))


def store_tensor(filename, tensors: Sequence[Tuple[str, Sequence[int], Optional[np.ndarray]]], out_bond_ids: Sequence[int]) -> None:
    """Writes a network file in the layout load_tensor reads: (name, bond ids, data) per tensor plus the output tensor
    ``-1``.  The reference creates such files only in its tests (hdf5.rs:141-170)."""
    names = [n for n, _, _ in tensors] + ["-1"]
    arrays = [None if a is None else np.asarray(a, dtype=np.complex128, order="C") for _, _, a in tensors] + [None]
    bids = [list(b) for _, b, _ in tensors] + [list(out_bond_ids)]
    n = len(names)
    c_names = (C.c_char_p * n)(*[s.encode() for s in names])
    shapes = [u64_array(a.shape if a is not None else []) for a in arrays]
    c_ranks = (C.c_int * n)(*[a.ndim if a is not None else 0 for a in arrays])
    c_dims = (C.POINTER(C.c_uint64) * n)(*[C.cast(s, C.POINTER(C.c_uint64)) for s in shapes])
    c_data = (C.c_void_p * n)(*[a.ctypes.data if a is not None else None for a in arrays])
    bid_arrays = [u64_array(b) for b in bids]
    c_nb = (C.c_int64 * n)(*[len(b) for b in bids])
    c_bids = (C.POINTER(C.c_uint64) * n)(*[C.cast(b, C.POINTER(C.c_uint64)) for b in bid_arrays])
    check(lib().tncb_hdf5_store(os.fsencode(filename), n, c_names, c_ranks, c_dims, c_data, c_nb, c_bids))
