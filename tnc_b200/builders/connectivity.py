"""Device coupling maps used by the circuit generators.

Data restated from the reference's topology tables (tnc/src/builders/connectivity.rs:59-254:
`sycamore_connect` and the four two-qubit layer patterns A-D of the Sycamore experiment);
`line` and `all` follow connectivity.rs:47-57,496-498.  Generated once with a script from
those tables; qubit labels are kept exactly as in the reference (the layer patterns are
1-based, `sycamore_connect` is used 0-based by `random_circuit`, random_circuit.rs:45-50).
The IBM heavy-hex layouts (Eagle/Osprey/Condor) are input generators outside the hot-path
scope and are not restated.
"""
from __future__ import annotations

from typing import List, Tuple

SYCAMORE_CONNECT: List[Tuple[int, int]] = [
    (52, 32), (32, 31), (31, 24), (24, 29), (29, 26), (26, 40), (40, 44), (44, 53),
    (37, 32), (32, 21), (21, 24), (24, 18), (18, 26), (26, 25), (25, 44), (44, 48),
    (37, 22), (22, 21), (21, 7), (7, 18), (18, 15), (15, 25), (25, 42), (42, 48),
    (35, 22), (22, 8), (8, 7), (7, 5), (5, 15), (15, 16), (16, 42), (42, 46),
    (35, 11), (11, 8), (8, 1), (1, 5), (5, 6), (6, 16), (16, 51), (51, 46),
    (11, 4), (4, 1), (1, 2), (2, 6), (6, 12), (12, 51), (51, 47), (14, 4),
    (4, 3), (3, 2), (2, 10), (10, 12), (12, 41), (41, 47), (36, 14), (14, 13),
    (13, 3), (3, 9), (9, 10), (10, 20), (20, 41), (41, 50), (36, 27), (27, 13),
    (13, 17), (17, 9), (9, 19), (19, 20), (20, 43), (43, 50), (38, 27), (27, 28),
    (28, 17), (17, 23), (23, 19), (19, 34), (34, 43), (43, 49), (38, 39), (39, 28),
    (28, 30), (30, 23), (23, 33), (33, 34), (34, 45), (45, 49),
]

SYCAMORE_A: List[Tuple[int, int]] = [
    (31, 32), (29, 24), (40, 26), (53, 44), (21, 22), (18, 7), (25, 15), (48, 42),
    (8, 11), (5, 1), (16, 6), (46, 51), (14, 4), (2, 3), (12, 10), (47, 41),
    (13, 27), (9, 17), (20, 19), (50, 43), (28, 39), (23, 30), (34, 33), (49, 45),
]

SYCAMORE_B: List[Tuple[int, int]] = [
    (32, 37), (24, 21), (26, 18), (44, 25), (22, 35), (7, 8), (15, 5), (42, 16),
    (1, 4), (6, 2), (51, 12), (14, 36), (3, 13), (10, 9), (41, 20), (27, 38),
    (17, 28), (19, 23), (43, 34),
]

SYCAMORE_C: List[Tuple[int, int]] = [
    (52, 32), (31, 24), (29, 26), (40, 44), (37, 22), (21, 7), (18, 15), (25, 42),
    (35, 11), (8, 1), (5, 6), (16, 51), (4, 3), (2, 10), (12, 41), (36, 27),
    (13, 17), (9, 19), (20, 43), (38, 39), (28, 30), (23, 33), (34, 45),
]

SYCAMORE_D: List[Tuple[int, int]] = [
    (32, 21), (24, 18), (26, 25), (44, 48), (22, 8), (7, 5), (15, 16), (42, 46),
    (11, 4), (1, 2), (6, 12), (51, 47), (14, 13), (3, 9), (10, 20), (41, 50),
    (27, 28), (17, 23), (19, 34), (43, 49),
]



def line_connect(n: int) -> List[Tuple[int, int]]:
    return [(i, i + 1) for i in range(n - 1)]


def all_connect(n: int) -> List[Tuple[int, int]]:
    assert n > 0
    return [(i, j) for i in range(n - 1) for j in range(i + 1, n)]


def connectivity(layout: str, n: int = 0) -> List[Tuple[int, int]]:
    """Connectivity::new (connectivity.rs:34-44) for the layouts restated here."""
    if layout == "sycamore":
        return list(SYCAMORE_CONNECT)
    if layout == "line":
        return line_connect(n)
    if layout == "all":
        return all_connect(n)
    raise NotImplementedError(f"connectivity layout {layout!r} is not restated (input generator, out of scope)")
