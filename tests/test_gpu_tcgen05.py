"""K1' (tcgen05 int8 digit slicing) parity: against the oracle and against the DMMA path.
Tolerance: ||gpu-cpu||_inf <= 1e-12 * max(1, ||cpu||_inf) with 8 slices (full mantissa)."""
import numpy as np
import pytest

from oracle import tnc_oracle as orc

pytestmark = pytest.mark.gpu


def rand_c(rng, shape, scale=1.0):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)) * scale


@pytest.fixture()
def tc_ctx(built_lib):
    import tnc_b200 as tb
    c = tb.Context(0)
    c.set_tcgen05_slices(8)
    c.set_tcgen05_threshold(1, 256)     # route every pair with M, N, K >= 256 to the tcgen05 engine
    yield c
    c.close()


def check(ctx, rng, a_legs, a_dims, b_legs, b_dims, tol=1e-12, scale_rows=False):
    import tnc_b200 as tb
    a, b = rand_c(rng, a_dims), rand_c(rng, b_dims)
    if scale_rows:  # wildly different magnitudes per slice of the leading free legs -> per-row exponents matter
        a = a * np.exp(rng.uniform(-40, 40, size=[a_dims[0]] + [1] * (len(a_dims) - 1)))
        b = b * np.exp(rng.uniform(-40, 40, size=[b_dims[0]] + [1] * (len(b_dims) - 1)))
    legs, got = tb.contract_pair(ctx, a_legs, a, b_legs, b)
    ref_legs, ref = orc.contract_pair(a_legs, a, b_legs, b)
    assert legs == ref_legs and got.shape == ref.shape
    err = np.abs(got - ref).max()
    assert err <= tol * max(1.0, np.abs(ref).max()), err
    return got, ref


def test_engine_is_really_tcgen05(tc_ctx):
    """Guard against silently testing the DMMA path: the tcgen05 step launches 5 kernels
    (2 exponent + 2 slicing + 1 GEMM) after the table build, the DMMA step 1 or 2."""
    import tnc_b200 as tb
    rng = np.random.default_rng(0)
    a = tb.DeviceTensor.from_numpy(tc_ctx, rand_c(rng, (256, 256)))
    b = tb.DeviceTensor.from_numpy(tc_ctx, rand_c(rng, (256, 256)))
    c = tb.DeviceTensor.empty(tc_ctx, (256, 256))
    tb.contract_pair_into(tc_ctx, [0, 1], a, [1, 2], b, c)     # builds the offset tables
    tc_ctx.reset_stats()
    tb.contract_pair_into(tc_ctx, [0, 1], a, [1, 2], b, c)
    assert tc_ctx.stats()["kernel_launches"] == 5
    tc_ctx.set_tcgen05_slices(0)
    tc_ctx.reset_stats()
    tb.contract_pair_into(tc_ctx, [0, 1], a, [1, 2], b, c)
    assert tc_ctx.stats()["kernel_launches"] <= 2      # k1_kernel (+ split-K reduce)
    tc_ctx.set_tcgen05_slices(8)


def test_tcgen05_square(tc_ctx):
    rng = np.random.default_rng(1)
    check(tc_ctx, rng, [0, 1], [256, 256], [1, 2], [256, 256])
    check(tc_ctx, rng, [0, 1], [512, 384], [1, 2], [384, 640])


def test_tcgen05_ragged(tc_ctx):
    rng = np.random.default_rng(2)
    check(tc_ctx, rng, [0, 1], [300, 333], [1, 2], [333, 260])      # M, N, K not multiples of 128
    check(tc_ctx, rng, [0, 1, 2], [7, 41, 300], [2, 3, 1], [300, 257, 41])  # permuted K legs, K = 287*... ragged


def test_tcgen05_permuted_circuit_like(tc_ctx):
    rng = np.random.default_rng(3)
    sh = list(range(100, 109)); af = list(range(9)); bf = list(range(50, 59))
    a_legs = [x for p in zip(af, sh) for x in p]
    b_legs = [x for p in zip(reversed(sh), bf) for x in p]
    check(tc_ctx, rng, a_legs, [2] * 18, b_legs, [2] * 18)           # M = N = K = 512, all dims 2, interleaved


def test_tcgen05_row_scaling(tc_ctx):
    rng = np.random.default_rng(4)
    # relative tolerance per output row group: compare in scaled units
    import tnc_b200 as tb
    a = rand_c(rng, [256, 256]); b = rand_c(rng, [256, 256])
    ra = np.exp(rng.uniform(-30, 30, size=(256, 1))); rb = np.exp(rng.uniform(-30, 30, size=(1, 256)))
    a2, b2 = a * ra, b * rb                                            # a rows (M) and b columns (N) scaled
    legs, got = tb.contract_pair(tc_ctx, [0, 1], a2, [1, 2], b2)
    _, ref = orc.contract_pair([0, 1], a2, [1, 2], b2)
    rel = np.abs(got - ref) / (rb.T * ra.T * 16.0 * np.ones_like(np.abs(ref)))
    assert rel.max() <= 1e-12, rel.max()


def test_tcgen05_long_k_chunks(tc_ctx):
    rng = np.random.default_rng(5)
    check(tc_ctx, rng, [0, 1], [256, 20000], [1, 2], [20000, 256])   # K > 8192: several int32-safe chunks


def test_tcgen05_matches_dmma_c2(tc_ctx, ctx):
    """Full-size C2 on both engines; also fewer slices degrade gracefully (7 bits per slice)."""
    import tnc_b200 as tb
    rng = np.random.default_rng(20240612)
    a_legs = list(range(12))
    b_legs = [x for p in zip([11, 9, 7, 5, 3, 1], range(12, 18)) for x in p]
    a = (rng.random([4] * 12) * 2 - 1) + 1j * (rng.random([4] * 12) * 2 - 1)
    b = (rng.random([4] * 12) * 2 - 1) + 1j * (rng.random([4] * 12) * 2 - 1)
    _, ref = tb.contract_pair(ctx, a_legs, a, b_legs, b)              # DMMA engine
    _, got = tb.contract_pair(tc_ctx, a_legs, a, b_legs, b)
    scale = np.abs(ref).max()
    e8 = np.abs(got - ref).max() / scale
    assert e8 <= 1e-13, e8
    tc_ctx.set_tcgen05_slices(6)
    _, got6 = tb.contract_pair(tc_ctx, a_legs, a, b_legs, b)
    e6 = np.abs(got6 - ref).max() / scale
    assert e8 < e6 <= 1e-9, (e8, e6)
    print(f"tcgen05 vs DMMA: 8 slices {e8:.2e}, 6 slices {e6:.2e}")
