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
