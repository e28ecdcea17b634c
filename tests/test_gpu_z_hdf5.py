"""(Named to sort after the other GPU files: the driver runs `pytest -x`.)
TensorData::File leaves through contract_tensor_network on the device (tensordata.rs:43-49 inside contraction.rs:66-76):
the library loads the HDF5 payloads while it stages the leaves, so a network whose gates come from files must give the same
amplitude as the same network with in-memory payloads -- bit for bit, because the schedule and the staged bytes are equal."""
import numpy as np
import pytest

from oracle import tnc_oracle as orc

pytestmark = pytest.mark.gpu


def _greedy(tn):
    from tnc_b200.contractionpath.paths import Cotengrust
    opt = Cotengrust(tn)
    opt.find_path()
    return opt.get_best_replace_path()


def _variants(tmp_path):
    """the same 12-qubit random-circuit amplitude network three times: gate leaves / Matrix leaves / File leaves (every
    third file stored adjointed and flagged adjoint, which must undo itself)"""
    from tnc_b200.builders import random_circuit
    from tnc_b200.gates import load_gate, load_gate_adjoint
    from tnc_b200.io import hdf5
    from tnc_b200.tensornetwork import Tensor, TensorData
    tn = random_circuit(12, 8, 0.5, 0.5, np.random.default_rng(77))
    mats, files = [], []
    for i, t in enumerate(tn.tensors):
        td = t.tensordata
        if td.kind == "gate":
            name, ang, adj = td.gate
            m = (load_gate_adjoint if adj else load_gate)(name, ang)
        else:
            m = np.asarray(td.matrix, dtype=np.complex128)
        m = m.reshape(t.bond_dims)
        mt = Tensor.new(t.legs, t.bond_dims); mt.set_tensor_data(TensorData.Matrix(m)); mats.append(mt)
        ft = Tensor.new(t.legs, t.bond_dims)
        p = str(tmp_path / ("leaf%d.h5" % i))
        r = m.ndim
        if i % 3 == 0 and r in (2, 4):           # store the adjoint, flag the leaf adjoint: into_data adjoints it back
            stored = np.conj(np.transpose(m, list(range(r // 2, r)) + list(range(r // 2))))
            hdf5.store_data(p, stored)
            ft.set_tensor_data(TensorData.File(p, True))
        else:
            hdf5.store_data(p, m)
            ft.set_tensor_data(TensorData.File(p, False))
        files.append(ft)
    return tn, Tensor.new_composite(mats), Tensor.new_composite(files)


def test_file_leaves_equal_matrix_leaves(ctx, tmp_path):
    from tnc_b200.tensornetwork import contract_tensor_network
    tn, mat_tn, file_tn = _variants(tmp_path)
    path = _greedy(tn)
    a_gate = complex(contract_tensor_network(tn, path, ctx=ctx).to_numpy())
    a_mat = complex(contract_tensor_network(mat_tn, path, ctx=ctx).to_numpy())
    a_file = complex(contract_tensor_network(file_tn, path, ctx=ctx).to_numpy())
    a_file2 = complex(contract_tensor_network(file_tn, path, ctx=ctx).to_numpy())      # second call: cached plan, files re-read
    assert a_mat == a_gate and a_file == a_mat and a_file2 == a_mat
    # against the oracle, which reads the same files with a reader of its own (oracle.load_data_hdf5) and adjoints them itself
    def to_o(t):
        if t.is_composite():
            return orc.OTensor(children=[to_o(c) for c in t.tensors])
        td = t.tensordata
        return orc.OTensor(list(t.legs), list(t.bond_dims), ("file", td.file[0], td.file[1]) if td.kind == "file" else np.asarray(td.matrix))
    ref = complex(orc.contract_tensor_network(to_o(file_tn), orc.OPath(list(path.toplevel), {})).data)
    ref_mat = complex(orc.contract_tensor_network(to_o(mat_tn), orc.OPath(list(path.toplevel), {})).data)
    assert ref == ref_mat
    assert abs(a_file - ref) <= 1e-9 * abs(ref) + 1e-18


def test_file_leaf_errors_leave_the_call_clean(ctx, tmp_path):
    from tnc_b200 import TncbError
    from tnc_b200.io import hdf5
    from tnc_b200.tensornetwork import Tensor, TensorData, contract_tensor_network
    from tnc_b200.contractionpath import ContractionPath
    rng = np.random.default_rng(1)
    a = rng.uniform(-1, 1, (2, 4)) + 1j * rng.uniform(-1, 1, (2, 4))
    b = rng.uniform(-1, 1, (4, 3)) + 1j * rng.uniform(-1, 1, (4, 3))
    hdf5.store_data(tmp_path / "a.h5", a)
    ta = Tensor.new([0, 1], [2, 4]); ta.set_tensor_data(TensorData.File(str(tmp_path / "a.h5"), False))
    tb = Tensor.new([1, 2], [4, 3]); tb.set_tensor_data(TensorData.Matrix(b))
    path = ContractionPath.simple([(0, 1)])
    got = contract_tensor_network(Tensor.new_composite([ta, tb]), path, ctx=ctx)
    assert got.legs == [2, 0]                                    # (b \ a) ++ (a \ b), tensor.rs:463-479
    np.testing.assert_allclose(got.to_numpy(), np.einsum("ik,kj->ji", a, b), rtol=0, atol=1e-14)
    # a missing file: `load_data(filename).unwrap()` panics in the reference, here TNCB_ERR_IO
    tm = Tensor.new([0, 1], [2, 4]); tm.set_tensor_data(TensorData.File(str(tmp_path / "missing.h5"), False))
    with pytest.raises(TncbError) as e:
        contract_tensor_network(Tensor.new_composite([tm, tb]), path, ctx=ctx)
    assert e.value.status == -10
    # a file whose shape is not the leaf's bond dimensions
    hdf5.store_data(tmp_path / "wrong.h5", a.reshape(4, 2))
    tw = Tensor.new([0, 1], [2, 4]); tw.set_tensor_data(TensorData.File(str(tmp_path / "wrong.h5"), False))
    with pytest.raises(TncbError) as e:
        contract_tensor_network(Tensor.new_composite([tw, tb]), path, ctx=ctx)
    assert e.value.status == -2
    # ... which the adjoint flag turns into the right shape: conj(a.reshape(4, 2)).T has dims [2, 4]
    tw.set_tensor_data(TensorData.File(str(tmp_path / "wrong.h5"), True))
    got = contract_tensor_network(Tensor.new_composite([tw, tb]), path, ctx=ctx)
    np.testing.assert_allclose(got.to_numpy(), np.einsum("ik,kj->ji", np.conj(a.reshape(4, 2)).T, b), rtol=0, atol=1e-14)
    # the context is still usable after the failures
    got = contract_tensor_network(Tensor.new_composite([ta, tb]), path, ctx=ctx)
    np.testing.assert_allclose(got.to_numpy(), np.einsum("ik,kj->ji", a, b), rtol=0, atol=1e-14)


def test_cpp_file_leaf(built_lib, tmp_path):
    """the same through the C++ mirror (tnc::TensorData::file, include/tnc.hpp)"""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([os.path.join(root, "build", "test_host_api"), "--file-leaf", str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "HOST_FILE_LEAF_OK" in r.stdout, r.stdout


def test_back_to_back_cached_calls_keep_their_own_payloads(built_lib):
    """Three contract_tensor_network calls with DIFFERENT payloads of one structure, issued without any synchronisation
    while the stream is still busy: each call's leaf upload must read its own payloads (the cached plan re-uses one
    pinned staging buffer; re-staging waits for the previous upload)."""
    import tnc_b200 as tb
    from tnc_b200.contractionpath import ContractionPath
    from tnc_b200.tensornetwork import Tensor, TensorData, contract_tensor_network
    c = tb.Context(0)
    try:
        rng = np.random.default_rng(4)

        def rnd(*s):
            return rng.uniform(-1, 1, s) + 1j * rng.uniform(-1, 1, s)

        def net(a, b, v):
            ta = Tensor.new([0, 1], [64, 64]); ta.set_tensor_data(TensorData.Matrix(a))
            tb_ = Tensor.new([1, 2], [64, 64]); tb_.set_tensor_data(TensorData.Matrix(b))
            tv = Tensor.new([2], [64]); tv.set_tensor_data(TensorData.Matrix(v))
            return Tensor.new_composite([ta, tb_, tv])

        path = ContractionPath.simple([(0, 1), (0, 2)])       # (a b) v: a 64^3 pair (K1, so no CUDA graph) and a matrix-vector pair
        data = [(rnd(64, 64), rnd(64, 64), rnd(64)) for _ in range(4)]
        for _ in range(2):                                     # first sighting runs pair by pair, the second compiles the plan
            contract_tensor_network(net(*data[0]), path, ctx=c).to_numpy()
        big = [tb.DeviceTensor.from_numpy(c, rnd(2048, 2048)) for _ in range(3)]
        for _ in range(4):                                     # a few ms of queued work: the uploads below wait behind it
            tb.contract_pair_into(c, [0, 1], big[0], [1, 2], big[1], big[2])
        nets = [net(*d) for d in data[1:]]
        res = [contract_tensor_network(n, path, ctx=c) for n in nets]
        for r, (a, b, v) in zip(res, data[1:]):
            assert r.legs == [0]
            exp = a @ b @ v
            assert np.abs(r.to_numpy() - exp).max() <= 1e-12 * np.abs(exp).max()
    finally:
        c.close()


@pytest.mark.parametrize("qubits,rounds,locations,seed", [(8, 4, [3], 1), (10, 5, [2, 7], 2), (12, 4, [0, 5, 11], 3)])
def test_observable_circuit_vs_oracle(ctx, qubits, rounds, locations, seed):
    """random_circuit_with_set_observable (random_circuit.rs:120-276): the light-cone expectation-value network contracts to
    the oracle's value (possibly several disconnected components -> outer products of scalars)."""
    from tnc_b200.builders import random_circuit_with_set_observable
    from tnc_b200.tensornetwork import contract_tensor_network
    tn = random_circuit_with_set_observable(qubits, rounds, 0.7, 0.7, locations, np.random.default_rng(seed), "line", qubits)
    path = _greedy(tn)

    def to_o(t):
        if t.is_composite():
            return orc.OTensor(children=[to_o(c) for c in t.tensors])
        td = t.tensordata
        d = ("gate", td.gate[0], td.gate[1], td.gate[2]) if td.kind == "gate" else np.asarray(td.matrix)
        return orc.OTensor(list(t.legs), list(t.bond_dims), d)

    res = contract_tensor_network(tn, path, ctx=ctx)
    ref = orc.contract_tensor_network(to_o(tn), orc.OPath(list(path.toplevel), {}))
    assert res.legs == ref.legs == []
    got, exp = complex(res.to_numpy()), complex(ref.data)
    assert abs(got - exp) <= 1e-9 * abs(exp) + 1e-15, (got, exp)
