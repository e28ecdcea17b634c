"""sycamore_circuit (tnc/src/builders/sycamore_circuit.rs:22-72): `depth` cycles of a random
single-qubit layer (sx/sy/sz on every qubit) followed by fsim(pi/2, pi/6) on the layer pattern
A B C D C D A B (1-based qubit labels), plus a final single-qubit layer.

The reference asserts qubits <= 49 (:27-30) although its tables reach qubit 53; the assert
is lifted to 53 here (BASELINE config 5), everything else is unchanged.  RNG: see
random_circuit.py."""
from __future__ import annotations

import math

import numpy as np

from .circuit_builder import Circuit
from .connectivity import SYCAMORE_A, SYCAMORE_B, SYCAMORE_C, SYCAMORE_D


def sycamore_circuit(qubits: int, depth: int, rng: np.random.Generator) -> Circuit:
    assert qubits <= 53, "the Sycamore tables cover 53 qubits"
    pattern = [SYCAMORE_A, SYCAMORE_B, SYCAMORE_C, SYCAMORE_D, SYCAMORE_C, SYCAMORE_D, SYCAMORE_A, SYCAMORE_B]
    single = ["sx", "sy", "sz"]
    c = Circuit()
    q = c.allocate_register(qubits)
    nxt = 0
    for rnd in range(depth + 1):
        for i in range(qubits):
            c.append_gate(single[int(rng.integers(0, 3))], [], [q[i]])
        if rnd < depth:
            layer = pattern[nxt % len(pattern)]
            nxt += 1
            for (i, j) in layer:
                if i > qubits or j > qubits:
                    continue
                c.append_gate("fsim", [math.pi / 2, math.pi / 6], [q[i - 1], q[j - 1]])
    return c
