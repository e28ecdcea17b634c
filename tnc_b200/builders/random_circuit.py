"""random_circuit (tnc/src/builders/random_circuit.rs:29-80): `rounds-1` rounds of random
single-qubit gates (sx/sy/sz, probability p1 per qubit) and fsim(0.3, 0.2) on the coupling map
(probability p2 per edge), closed with <0| bras (amplitude network).

The reference draws from Rust's StdRng (ChaCha12); that stream is not reproduced -- numpy's
PCG64 is used, so a seed names a different (equally distributed) circuit.  The structure --
gate set, leg numbering, traversal order -- is the reference's."""
from __future__ import annotations

from typing import Optional

import numpy as np

from ..tensornetwork.tensor import Tensor
from .circuit_builder import Circuit
from .connectivity import connectivity


def random_circuit_builder(qubits: int, rounds: int, single_qubit_probability: float, two_qubit_probability: float,
                           rng: np.random.Generator, layout: str = "sycamore", layout_n: int = 0) -> Circuit:
    single = ["sx", "sy", "sz"]
    edges = [(u, v) for (u, v) in connectivity(layout, layout_n or qubits) if u < qubits and v < qubits]
    c = Circuit()
    q = c.allocate_register(qubits)
    for _ in range(1, rounds):
        for i in range(qubits):
            if rng.random() < single_qubit_probability:
                c.append_gate(single[int(rng.integers(0, 3))], [], [q[i]])
        for (i, j) in edges:
            if rng.random() < two_qubit_probability:
                c.append_gate("fsim", [0.3, 0.2], [q[i], q[j]])
    return c


def random_circuit(qubits: int, rounds: int, single_qubit_probability: float, two_qubit_probability: float,
                   rng: np.random.Generator, layout: str = "sycamore", layout_n: int = 0) -> Tensor:
    c = random_circuit_builder(qubits, rounds, single_qubit_probability, two_qubit_probability, rng, layout, layout_n)
    return c.into_amplitude_network("0" * qubits)[0]


def random_sparse_tensor_data_with_rng(dims, sparsity: Optional[float], rng: np.random.Generator):
    """builders/tensorgeneration.rs:19-53: a zero tensor into which random positions receive U(0,1) + i U(0,1) values until
    the number of draws / size reaches `sparsity` (default 0.5; positions may repeat, as in the reference)."""
    from ..tensornetwork.tensordata import TensorData
    sparsity = 0.5 if sparsity is None else float(np.float32(sparsity))
    assert 0.0 <= sparsity <= 1.0
    dims = [int(d) for d in dims]
    size = int(np.prod(dims, dtype=np.int64)) if dims else 1
    data = np.zeros(dims, dtype=np.complex128)
    nnz = 0
    while np.float32(nnz) / np.float32(size) < np.float32(sparsity):
        loc = tuple(int(rng.integers(0, d)) for d in dims)
        data[loc] = complex(rng.random(), rng.random())
        nnz += 1
    return TensorData.Matrix(data)


def random_circuit_with_observable(qubits: int, rounds: int, single_qubit_probability: float, two_qubit_probability: float,
                                   observable_probability: float, rng: np.random.Generator, layout: str = "sycamore",
                                   layout_n: int = 0) -> Tensor:
    """random_circuit.rs:88-113: observables on a random subset of the qubits, then random_circuit_with_set_observable."""
    locations = [i for i in range(qubits) if rng.random() < observable_probability]
    return random_circuit_with_set_observable(qubits, rounds, single_qubit_probability, two_qubit_probability, locations, rng,
                                              layout, layout_n)


def random_circuit_with_set_observable(qubits: int, rounds: int, single_qubit_probability: float, two_qubit_probability: float,
                                       observable_location, rng: np.random.Generator, layout: str = "sycamore",
                                       layout_n: int = 0) -> Tensor:
    """random_circuit.rs:120-276: an expectation-value network <psi| U^dagger O U |psi> grown outwards from a layer of random
    Pauli observables: per round, fsim(0.3, 0.2) pairs and single-qubit gates are placed (with their adjoint mirror image on the
    other side) only where they touch the light cone of an observable; a random product state closes both sides.  Leg numbering
    and tensor order are the reference's (its two structural tests are replayed in tests/test_paths.py); the mirror partner of
    sy and sz is sx-adjoint exactly as in the reference (:134-147)."""
    from ..tensornetwork.tensordata import TensorData
    single = [(TensorData.Gate("sx"), TensorData.Gate("sx", (), True)),
              (TensorData.Gate("sy"), TensorData.Gate("sx", (), True)),
              (TensorData.Gate("sz"), TensorData.Gate("sx", (), True))]
    observables = ["x", "y", "z"]
    observable_location = [int(i) for i in observable_location]
    tn = Tensor()
    open_edges = {}
    next_edge = 0
    final_state = []
    for i in range(qubits):
        if i in observable_location:
            open_edges[i] = (next_edge, next_edge + 1)
            next_edge += 2
            t = Tensor.new_from_const([open_edges[i][0], open_edges[i][1]], 2)
            t.set_tensor_data(TensorData.Gate(observables[int(rng.integers(0, 3))]))
            final_state.append(t)
        else:
            open_edges[i] = (0, 0)
    tn.push_tensors(final_state)
    edges = [(u, v) for (u, v) in connectivity(layout, layout_n or qubits) if u < qubits and v < qubits]
    gates = []
    for _ in range(1, rounds):
        for (i, j) in edges:
            if rng.random() < two_qubit_probability and (open_edges[i][0] != open_edges[i][1] or open_edges[j][0] != open_edges[j][1]):
                if open_edges[i][0] != open_edges[i][1]:
                    left_i, right_i = open_edges[i]
                else:
                    next_edge += 1
                    left_i = right_i = next_edge - 1
                if open_edges[j][0] != open_edges[j][1]:
                    left_j, right_j = open_edges[j]
                else:
                    next_edge += 1
                    left_j = right_j = next_edge - 1
                left = Tensor.new_from_const([next_edge, next_edge + 1, left_i, left_j], 2)
                left.set_tensor_data(TensorData.Gate("fsim", (0.3, 0.2), False))
                gates.append(left)
                right = Tensor.new_from_const([right_i, right_j, next_edge + 2, next_edge + 3], 2)
                right.set_tensor_data(TensorData.Gate("fsim", (0.3, 0.2), True))
                gates.append(right)
                open_edges[i] = (next_edge, next_edge + 2)
                open_edges[j] = (next_edge + 1, next_edge + 3)
                next_edge += 4
        for i in range(qubits):
            left_index, right_index = open_edges[i]
            if rng.random() < single_qubit_probability and left_index != right_index:
                lg, rg = single[int(rng.integers(0, 3))]
                left = Tensor.new_from_const([next_edge, left_index], 2)
                left.set_tensor_data(lg)
                gates.append(left)
                right = Tensor.new_from_const([right_index, next_edge + 1], 2)
                right.set_tensor_data(rg)
                gates.append(right)
                open_edges[i] = (next_edge, next_edge + 1)
                next_edge += 2
    tn.push_tensors(gates)
    initial = []
    for i in range(qubits):
        left_index, right_index = open_edges[i]
        if left_index != right_index:
            state = random_sparse_tensor_data_with_rng([2], 1.0, rng)
            lt = Tensor.new_from_const([left_index], 2)
            lt.set_tensor_data(state)
            initial.append(lt)
            rt = Tensor.new_from_const([right_index], 2)
            rt.set_tensor_data(TensorData.Matrix(np.array(state.matrix)))    # `.clone()`: the same values on both sides
            initial.append(rt)
    tn.push_tensors(initial)
    return tn
