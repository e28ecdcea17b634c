"""ContractionPath (tnc/src/contractionpath.rs:29-35) and the SSA -> replace-left conversion
(:197-215).  The path is the *input* of the hot path."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

SimplePath = List[Tuple[int, int]]


@dataclass
class ContractionPath:
    nested: Dict[int, "ContractionPath"] = field(default_factory=dict)
    toplevel: SimplePath = field(default_factory=list)

    @classmethod
    def simple(cls, path: Sequence[Tuple[int, int]]) -> "ContractionPath":
        return cls({}, [(int(a), int(b)) for a, b in path])

    @classmethod
    def nested_path(cls, nested: Sequence[Tuple[int, "ContractionPath"]], toplevel) -> "ContractionPath":
        return cls({int(i): p for i, p in nested}, [(int(a), int(b)) for a, b in toplevel])

    @classmethod
    def single(cls, a: int, b: int) -> "ContractionPath":
        return cls.simple([(a, b)])

    def __len__(self) -> int:
        return len(self.toplevel)

    def is_empty(self) -> bool:
        return not self.toplevel

    def is_simple(self) -> bool:
        return not self.nested

    def into_simple(self) -> SimplePath:
        assert self.is_simple()
        return self.toplevel


def path(*pairs, nested=None) -> ContractionPath:
    """The `path!` macro (contractionpath.rs:154-167):
    path((0, 1), (0, 2), nested={2: [(0, 2), (0, 1)]})."""
    n = {}
    for k, v in (nested or {}).items():
        n[int(k)] = v if isinstance(v, ContractionPath) else path(*v)
    return ContractionPath(n, [(int(a), int(b)) for a, b in pairs])


def ssa_ordering(p: Sequence[Tuple[int, int, int]], n: int) -> ContractionPath:
    """contractionpath.rs:180-192."""
    hs: Dict[int, int] = {}
    out = []
    path_len = n
    for (u1, u2, u3) in p:
        t1 = hs[u1] if u1 >= path_len else u1
        t2 = hs[u2] if u2 >= path_len else u2
        hs.setdefault(u3, n)
        n += 1
        out.append((t1, t2))
    return ContractionPath.simple(out)


def ssa_replace_ordering(p: ContractionPath) -> ContractionPath:
    """contractionpath.rs:197-215: every SSA id maps to the slot of its left parent."""
    nested = {i: ssa_replace_ordering(q) for i, q in p.nested.items()}
    hs: Dict[int, int] = {}
    top = []
    n = len(p.toplevel) + 1
    for (t0, t1) in p.toplevel:
        n0, n1 = hs.get(t0, t0), hs.get(t1, t1)
        assert n not in hs
        hs[n] = n0
        n += 1
        top.append((n0, n1))
    return ContractionPath(nested, top)


def validate_path(p: ContractionPath) -> bool:
    """contractionpath/paths.rs:44-57: a slot used on the right never reappears on the left."""
    gone = set()
    for q in p.nested.values():
        if not validate_path(q):
            return False
    for (a, b) in p.toplevel:
        if a in gone:
            return False
        gone.add(b)
    return True
