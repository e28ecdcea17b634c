"""Host-side logic of K1' (the modular / CRT int8 engine, csrc/crt.cu), no GPU: the moduli are pairwise coprime, the
split CRT weights reconstruct exact integers, the modulus count / operand bits obey log2(P) >= a + b + log2 K + 3,
and a numpy emulation of the whole pipeline (truncate -> residues -> int32 GEMM mod m -> CRT with the LIBRARY's
tables) meets the guaranteed bound against exact big-integer dot products."""
import math
from fractions import Fraction

import numpy as np
import pytest


def test_moduli_pairwise_coprime_and_weights(built_lib):
    import tnc_b200 as tb
    for n in (2, 7, 16, 20):
        t = tb.tcgen05_tables(n)
        ms = t["moduli"]
        assert all(2 <= m <= 256 for m in ms) and ms[0] == 256 and 255 not in ms
        assert all(math.gcd(a, b) == 1 for i, a in enumerate(ms) for b in ms[i + 1:])
        P = math.prod(ms)
        assert abs(t["log2_product"] - math.log2(P)) < 1e-9
        for m, r1, r2 in zip(ms, t["rho1"], t["rho2"]):
            inv = pow((P // m) % m, -1, m)
            exact = Fraction(inv, m)
            assert Fraction(r1) * 2 ** 34 == int(Fraction(r1) * 2 ** 34)          # rho1 sits on the 2^-34 grid
            assert 0 <= r2 < 2.0 ** -34
            assert abs(Fraction(r1) + Fraction(r2) - exact) < Fraction(1, 2 ** 85)  # rho2 carries the rest to double precision


def test_crt_reconstructs_exact_integers(built_lib):
    import tnc_b200 as tb
    rng = np.random.default_rng(3)
    t = tb.tcgen05_tables(16)
    ms = t["moduli"]; P = math.prod(ms)
    r1, r2 = np.array(t["rho1"]), np.array(t["rho2"])
    for _ in range(200):
        x = int(rng.integers(-2 ** 62, 2 ** 62)) * int(rng.integers(0, 2 ** 60)) // int(rng.integers(1, 2 ** 30))   # |x| < P/4
        assert abs(x) < P // 4
        y = np.array([((x % m) + m // 2) % m - m // 2 for m in ms], dtype=np.float64)   # any representative in [-128, 127]
        s1, s2 = float(np.sum(y * r1)), float(np.dot(y, r2))       # exact / double
        assert s1 == float(sum(Fraction(int(v)) * Fraction(a) for v, a in zip(y, r1)))
        frac = (s1 - np.rint(s1 + s2)) + s2
        got = frac * float(P)
        assert abs(got - x) <= abs(x) * 2.0 ** -50 + float(P) * 2.0 ** -70, (x, got)


@pytest.mark.parametrize("k,rel", [(4096, 0.0), (4096, 1e-10), (1 << 20, 0.0), (256, 1e-6), (1 << 15, 1e-3)])
def test_bound_and_modulus_count(built_lib, k, rel):
    import tnc_b200 as tb
    b = tb.tcgen05_bound(k, rel)
    lp = tb.tcgen05_tables(b["n_moduli"])["log2_product"]
    assert lp >= b["bits_a"] + b["bits_b"] + math.log2(k) + 3                    # |C'| < P/4
    assert b["bits_a"] <= 53 and b["bits_b"] <= 53
    want = 53 if rel == 0 else min(53, math.ceil(math.log2(16 * k / rel)))
    assert min(b["bits_a"], b["bits_b"]) >= want                                 # the tolerance is honoured ...
    if rel:
        assert b["bound"] <= rel * (1 + 1e-12)
    if b["n_moduli"] > 2:                                                        # ... with the fewest moduli
        assert tb.tcgen05_tables(b["n_moduli"] - 1)["log2_product"] < 2 * want + math.log2(k) + 3
    assert tb.tcgen05_bound(4096)["n_moduli"] == 16 and tb.tcgen05_bound(4096, 1e-10)["n_moduli"] == 15


def emulate(A, B, n_moduli):
    """numpy restatement of crt.cu for C[n,m] = sum_k B[n,k] A[m,k] with the library's tables."""
    import tnc_b200 as tb
    K = A.shape[1]
    bd = tb.tcgen05_bound(K, 0.0, n_moduli)
    t = tb.tcgen05_tables(bd["n_moduli"])
    ms, r1, r2 = t["moduli"], np.array(t["rho1"]), np.array(t["rho2"])

    def prep(X, bits):
        mx = np.maximum(np.abs(X.real), np.abs(X.imag)).max(axis=1)
        e = np.where(mx > 0, np.frexp(mx)[1], 0)            # max * 2^-e in [0.5, 1)
        sc = np.ldexp(1.0, -e)[:, None]
        return np.trunc(X.real * sc * 2.0 ** bits), np.trunc(X.imag * sc * 2.0 ** bits), e
    ar, ai, ea = prep(A, bd["bits_a"]); br, bi, eb = prep(B, bd["bits_b"])
    s1r = s2r = s1i = s2i = 0.0
    RM = 6755399441055744.0
    for i, m in enumerate(ms):
        def res(X):
            q = (X * (1.0 / m) + RM) - RM
            r = (X - q * m).astype(np.int64)                 # (exact here: |X| < 2^53, q*m rounds but r is small; checked below)
            r = (r + 128) % 256 - 128 if m == 256 else r
            assert np.abs(r).max() <= 128
            return r
        a_r, a_i, b_r, b_i = res(ar), res(ai), res(br), res(bi)
        accr, acci = b_r @ a_r.T - b_i @ a_i.T, b_r @ a_i.T + b_i @ a_r.T
        assert max(np.abs(accr).max(), np.abs(acci).max()) < 2 ** 31
        magic = int(round(2 ** 32 / m))

        def zmod(acc):       # the epilogue of crt_gemm_kernel: __mulhi quotient, offset byte, one wrap
            z = acc - ((acc * magic) >> 32) * m + 128
            z = z - (z >> 8) * m
            assert z.min() >= 0 and z.max() <= 255 and np.all((z - 128 - acc) % m == 0)
            return (z - 128).astype(np.float64)
        zr, zi = zmod(accr), zmod(acci)
        s1r, s2r, s1i, s2i = s1r + zr * r1[i], s2r + zr * r2[i], s1i + zi * r1[i], s2i + zi * r2[i]
    P = float(math.prod(ms))
    sc = np.ldexp(P, -(bd["bits_a"] + bd["bits_b"])) * np.ldexp(1.0, eb[:, None] + ea[None, :])
    fin = lambda s1, s2: (s1 - np.rint(s1 + s2)) + s2
    return (fin(s1r, s2r) + 1j * fin(s1i, s2i)) * sc, bd


@pytest.mark.parametrize("n_moduli", [16, 13, 10])
def test_numpy_emulation_meets_the_bound(built_lib, n_moduli):
    rng = np.random.default_rng(11)
    M, N, K = 24, 20, 1024
    A = (rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K))) * np.exp(rng.uniform(-20, 20, (M, 1)))
    B = (rng.standard_normal((N, K)) + 1j * rng.standard_normal((N, K))) * np.exp(rng.uniform(-20, 20, (N, 1)))
    A[3] = 0.0                                              # an all-zero row
    got, bd = emulate(A, B, n_moduli)
    ref = B.astype(np.clongdouble) @ A.astype(np.clongdouble).T
    mxa = np.maximum(np.abs(A.real), np.abs(A.imag)).max(axis=1); mxb = np.maximum(np.abs(B.real), np.abs(B.imag)).max(axis=1)
    scale = mxb[:, None] * np.where(mxa > 0, mxa, 1.0)[None, :]
    err = np.abs(got - ref) / scale
    assert err.max() <= bd["bound"], (err.max(), bd)
    assert np.all(got[:, 3] == 0)
    if n_moduli == 16:
        assert np.abs(got - ref).max() / np.abs(ref).max() < 5e-15
