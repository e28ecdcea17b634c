"""Pins the CPU oracle against every golden vector / known-answer test the reference holds
for the hot path.  CPU only."""
import math

import numpy as np
import pytest

from oracle import tnc_oracle as orc

F64_EPS = np.finfo(np.float64).eps


def approx_default(a, b):
    """float_cmp F64Margin{epsilon: f64::EPSILON, ulps: 4} per component."""
    a, b = np.asarray(a), np.asarray(b)
    for x, y in ((a.real, b.real), (a.imag, b.imag)):
        d = np.abs(x - y)
        ulp = np.spacing(np.maximum(np.abs(x), np.abs(y)))
        assert np.all((d <= F64_EPS) | (d <= 4 * ulp)), f"max diff {d.max()}"


def leaf(t):
    return orc.OTensor(list(t["legs"]), list(t["shape"]), t["data"])


# contraction.rs:121-178 test_tensor_contraction
def test_tensor_contraction_kat(kat):
    legs, res = orc.contract_pair(kat["A"]["legs"], kat["A"]["data"], kat["B"]["legs"], kat["B"]["data"])
    assert legs == kat["AxB"]["legs"] and list(res.shape) == kat["AxB"]["shape"]
    assert np.abs(res - kat["AxB"]["data"]).max() <= 1e-14
    legs, res = orc.contract_pair(kat["B"]["legs"], kat["B"]["data"], kat["C"]["legs"], kat["C"]["data"])
    assert legs == kat["BxC"]["legs"] and list(res.shape) == kat["BxC"]["shape"]
    assert np.abs(res - kat["BxC"]["data"]).max() <= 1e-14


# contraction.rs:181-224 test_tn_contraction
def test_tn_contraction_kat(kat):
    tn = orc.OTensor(children=[leaf(kat["A"]), leaf(kat["B"]), leaf(kat["C"])])
    res = orc.contract_tensor_network(tn, orc.OPath([(0, 1), (0, 2)]))
    assert res.legs == kat["ABxC"]["legs"] and res.dims == kat["ABxC"]["shape"]
    assert np.abs(res.data - kat["ABxC"]["data"]).max() <= 1e-14


def test_torch_backend_matches(kat):
    tn = orc.OTensor(children=[leaf(kat["A"]), leaf(kat["B"]), leaf(kat["C"])])
    res = orc.contract_tensor_network(tn, orc.OPath([(0, 1), (0, 2)]), backend="torch")
    assert np.abs(res.data.numpy() - kat["ABxC"]["data"]).max() <= 1e-13


# contraction.rs:227-264 test_outer_product_contraction
def test_outer_product_kat():
    t1 = orc.OTensor([0], [3], np.array([1, 2 + 5j, 3 - 1j]))
    t2 = orc.OTensor([1], [2], np.array([-4 + 2j, -1j]))
    res = orc.contract_tensor_network(orc.OTensor(children=[t1, t2]), orc.OPath([(0, 1)]))
    assert res.legs == [1, 0] and res.dims == [2, 3]
    exp = np.array([-4 + 2j, -18 - 16j, -10 + 10j, -1j, 5 - 2j, -1 - 3j]).reshape(2, 3)
    approx_default(res.data, exp)


def chain_path(n):
    return orc.OPath([(0, i) for i in range(1, n)])  # qasm_importer.rs:163-169 contract_tn


def odd_test_circuit():
    c = orc.OCircuit()
    q = c.allocate_register(3)
    c.append_gate("rx", [0.5], [q[0]]); c.append_gate("rx", [0.2], [q[1]]); c.append_gate("rx", [0.3], [q[2]])
    c.append_gate("cx", [], [q[0], q[1]]); c.append_gate("cx", [], [q[1], q[2]])
    return c


SV8 = np.array([0.953246407214305, -0.14406910361762032j, -0.014455126269118733, -0.09564366568448116j,
                -0.024421837348497916, 0.0036909997130494475j, -0.03678688170631573, -0.24340376901515096j])


# io/qasm/qasm_importer.rs:171-194 bell_contract
def test_bell_kat():
    c = orc.OCircuit(); q = c.allocate_register(2)
    c.append_gate("h", [], [q[0]]); c.append_gate("cx", [], [q[0], q[1]])
    tn, final = c.into_statevector_network()
    res = orc.permute_to(orc.contract_tensor_network(tn, chain_path(len(tn.children))), final)
    approx_default(res.data.reshape(-1), np.array([orc.FRAC_1_SQRT_2, 0, 0, orc.FRAC_1_SQRT_2]))


# qasm_importer.rs:196-224 custom_swap (myswap inlined to three cx)
def test_custom_swap_kat():
    c = orc.OCircuit(); q = c.allocate_register(2)
    c.append_gate("x", [], [q[0]])
    a, b = q[1], q[0]
    c.append_gate("cx", [], [a, b]); c.append_gate("cx", [], [b, a]); c.append_gate("cx", [], [a, b])
    tn, final = c.into_statevector_network()
    res = orc.permute_to(orc.contract_tensor_network(tn, chain_path(len(tn.children))), final)
    approx_default(res.data.reshape(-1), np.array([0, 1, 0, 0]))


# qasm_importer.rs:239-298
def test_statevector_order_kats():
    tn, final = odd_test_circuit().into_statevector_network()
    res = orc.permute_to(orc.contract_tensor_network(tn, chain_path(len(tn.children))), final)
    approx_default(res.data.reshape(-1), SV8)
    tn, final = odd_test_circuit().into_amplitude_network("1*0")
    res = orc.permute_to(orc.contract_tensor_network(tn, chain_path(len(tn.children))), final)
    approx_default(res.data.reshape(-1), SV8[[4, 6]])
    tn, final = odd_test_circuit().into_amplitude_network("*1*")
    res = orc.permute_to(orc.contract_tensor_network(tn, chain_path(len(tn.children))), final)
    approx_default(res.data.reshape(-1), SV8[[2, 3, 6, 7]])


# builders/circuit_builder.rs:372-396 hadamards_amplitude (any full path gives the same scalar)
def test_hadamards_amplitude_kat():
    c = orc.OCircuit(); q = c.allocate_register(5)
    for x in q:
        c.append_gate("h", [], [x])
    tn, final = c.into_amplitude_network("00000")
    assert final == []
    res = orc.contract_tensor_network(tn, chain_path(len(tn.children)))
    assert res.legs == []
    approx_default(res.data.reshape(-1), np.array([orc.FRAC_1_SQRT_2 ** 5]))


# circuit_builder.rs:399-427 rx_expectation_value
def test_rx_expectation_kat():
    c = orc.OCircuit(); q = c.allocate_register(2)
    c.append_gate("rx", [math.pi / 4], [q[0]]); c.append_gate("rx", [math.pi / 3], [q[1]])
    tn = c.into_expectation_value_network()
    # connected components first, then the outer product (a valid replace-left path)
    res = orc.contract_tensor_network(tn, chain_path(len(tn.children)))
    approx_default(res.data.reshape(-1), np.array([orc.FRAC_1_SQRT_2 * 0.5]))


# gates.rs:657-680: specialised adjoints == generic conj-transpose
def test_adjoint_identity():
    rng = np.random.default_rng(42)
    nparams = {"u": 3, "rx": 1, "ry": 1, "rz": 1, "cp": 1, "fsim": 2}
    special = {
        "rx": lambda a: orc.gate_matrix("rx", [-a[0]]), "ry": lambda a: orc.gate_matrix("ry", [-a[0]]),
        "rz": lambda a: orc.gate_matrix("rz", [-a[0]]), "cp": lambda a: orc.gate_matrix("cp", [-a[0]]),
        "fsim": lambda a: orc.gate_matrix("fsim", [-a[0], -a[1]]),
        "t": lambda a: np.conj(orc.gate_matrix("t", [])), "sx": lambda a: np.conj(orc.gate_matrix("sx", [])),
        "sz": lambda a: np.conj(orc.gate_matrix("sz", [])), "iswap": lambda a: np.conj(orc.gate_matrix("iswap", [])),
    }
    for g in ["x", "y", "z", "h", "t", "u", "sx", "sy", "sz", "rx", "ry", "rz", "cx", "cz", "swap", "cp", "iswap", "fsim"]:
        a = list(rng.uniform(-math.pi, math.pi, nparams.get(g, 0)))
        gen = orc.load_gate(g, a, True)
        m = orc.gate_matrix(g, a)
        d = 2 if m.size == 4 else 4
        assert np.array_equal(gen.reshape(d, d), np.conj(m.reshape(d, d).T))
        if g in special:
            approx_default(gen, special[g](a))
        if g in ("x", "y", "z", "h", "cx", "cz", "swap"):  # self-adjoint
            approx_default(gen, m)


def test_gate_errors():
    with pytest.raises(KeyError, match="Gate 'foo' not found."):
        orc.gate_matrix("foo", [])
    with pytest.raises(ValueError, match="Expected 0 angles, but got 2."):
        orc.gate_matrix("x", [2.0, 4.0])


# contractionpath.rs:267-327 ssa_replace_ordering KATs
def test_ssa_replace_ordering():
    p = orc.ssa_replace_ordering(orc.OPath([(0, 3), (1, 2), (6, 4), (5, 7), (9, 8), (11, 10)]))
    assert p.toplevel == [(0, 3), (1, 2), (6, 4), (5, 0), (6, 1), (6, 5)]
    n = orc.ssa_replace_ordering(orc.OPath([(0, 3), (1, 2), (6, 4), (5, 7), (9, 8), (11, 10)],
                                           {1: orc.OPath([(2, 1), (0, 3)]), 6: orc.OPath([(0, 2), (1, 3), (4, 5)])}))
    assert n.nested[1].toplevel == [(2, 1), (0, 2)] and n.nested[6].toplevel == [(0, 2), (1, 3), (0, 1)]
    # cotengrust.rs:241-305 pairs of (ssa path, replace path)
    for ssa, rep in [([(0, 1), (3, 2)], [(0, 1), (0, 2)]), ([(0, 1), (2, 3), (4, 5)], [(0, 1), (2, 3), (0, 2)]),
                     ([(2, 1), (0, 3)], [(2, 1), (0, 2)]), ([(0, 1), (2, 3), (5, 4)], [(0, 1), (2, 3), (2, 0)]),
                     ([(1, 5), (3, 4), (6, 0), (7, 2), (9, 8)], [(1, 5), (3, 4), (1, 0), (3, 2), (3, 1)])]:
        assert orc.ssa_replace_ordering(orc.OPath(ssa)).toplevel == rep


def test_nested_equals_flat():
    """integration_tests.rs:22-83 property: partitioned == flat."""
    rng = np.random.default_rng(3)
    def rt(legs, dims):
        return orc.OTensor(legs, dims, rng.standard_normal(dims) + 1j * rng.standard_normal(dims))
    t = [rt([0, 1], [2, 3]), rt([1, 2], [3, 4]), rt([2, 3], [4, 2]), rt([3, 0], [2, 2])]
    flat = orc.contract_tensor_network(orc.OTensor(children=list(t)), orc.OPath([(0, 1), (0, 2), (0, 3)]))
    nested_tn = orc.OTensor(children=[orc.OTensor(children=t[:2]), orc.OTensor(children=t[2:])])
    nested = orc.contract_tensor_network(nested_tn, orc.OPath([(0, 1)], {0: orc.OPath([(0, 1)]), 1: orc.OPath([(0, 1)])}))
    assert abs(complex(flat.data) - complex(nested.data)) < 1e-13
