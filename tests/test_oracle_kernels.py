"""Pins the oracle's SIMD-kernel restatements against the reference's own property tests
(reference src/math/util.rs:752-968: 32-ULP equality with NaN/inf tolerance; regression inputs from
proptest-regressions/math.txt and proptest-regressions/math/util.txt)."""
import math

import numpy as np
import pytest

from helpers import any_f64, assert_approx_eq, exact_fma

NCASES = 300


def _sizes(rng, maxsize):
    return int(rng.integers(0, maxsize))


def test_check_logaddexp(orc):
    # util.rs:880-890
    rng = np.random.default_rng(1)
    for _ in range(NCASES):
        x, y = rng.uniform(-10, 10, 2)
        a = math.log(math.exp(x) + math.exp(y))
        b = orc.logaddexp(x, y)
        assert abs(a - b) < 1e-10
        assert b == orc.logaddexp(y, x)
        assert x == orc.logaddexp(x, -math.inf)
        assert orc.logaddexp(-math.inf, -math.inf) == -math.inf
        assert math.isnan(orc.logaddexp(math.nan, x))
    # proptest-regressions/math.txt:7
    x, y = 4.8329699435311735, 9.38911339170414
    assert abs(math.log(math.exp(x) + math.exp(y)) - orc.logaddexp(x, y)) < 1e-10
    assert orc.logaddexp(x, y) == orc.logaddexp(y, x)


def test_check_neginf(orc):
    # util.rs:964-968
    assert orc.logaddexp(-math.inf, 2.0) == 2.0
    assert orc.logaddexp(2.0, -math.inf) == 2.0
    assert orc.logaddexp(1.5, 1.5) == 1.5 + math.log(2.0)


def _check_axpy(orc, x, y, a):
    out = orc.axpy(x, y, a)
    for xi, yi, oi in zip(x, y, out):
        assert_approx_eq(oi, exact_fma(a, xi, yi))


def test_axpy(orc):
    # util.rs:893-901 : out == a.mul_add(x, y)
    rng = np.random.default_rng(2)
    for _ in range(NCASES):
        n = _sizes(rng, 10)
        _check_axpy(orc, any_f64(rng, n), any_f64(rng, n), float(any_f64(rng, 1)[0]))
    # proptest-regressions/math.txt:8,10
    _check_axpy(orc, [2.9394791070664547e110, 0.0], [math.inf, 0.0], -2.4153502104628106e222)
    _check_axpy(orc, [0.0, 0.0, 0.0, 1.2271235629394547e205, 0.0, 0.0, -0.0, 0.0],
                [0.0, 0.0, 0.0, 7.121658452243713e81, 0.0, 0.0, 0.0, 0.0], -6.261465657118442e-124)
    # sizes that exercise SIMD body + SIMD tail + scalar tail (benches/sample.rs:126: 4,16,17,100,4567)
    for n in (4, 16, 17, 100, 4567):
        x, y = rng.normal(size=n), rng.normal(size=n)
        np.testing.assert_array_equal(orc.axpy(x, y, 0.37), [exact_fma(0.37, a, b) for a, b in zip(x, y)])


def _check_axpy_out(orc, a, x, y):
    out = orc.axpy_out(x, y, a)
    with np.errstate(all="ignore"):
        expect = np.asarray(y, dtype=np.float64) + np.float64(a) * np.asarray(x, dtype=np.float64)  # ndarray scaled_add
    for o, e in zip(out, expect):
        assert_approx_eq(o, e)


def test_axpy_out(orc):
    # util.rs:929-939
    rng = np.random.default_rng(3)
    for _ in range(NCASES):
        n = _sizes(rng, 10)
        _check_axpy_out(orc, float(any_f64(rng, 1)[0]), any_f64(rng, n), any_f64(rng, n))
    # proptest-regressions/math.txt:11
    _check_axpy_out(orc, 1.033664102276113e155, [-1.847508293460042e-54, 0.0, 0.0], [1.8293708670672727e101, 0.0, 0.0])


def test_multiply(orc):
    # util.rs:942-951
    rng = np.random.default_rng(4)
    for _ in range(NCASES):
        n = _sizes(rng, 10)
        x, y = any_f64(rng, n), any_f64(rng, n)
        with np.errstate(all="ignore"):
            expect = x * y
        for o, e in zip(orc.multiply(x, y), expect):
            assert_approx_eq(o, e)


def _seq_dot(x, y):
    s = 0.0
    with np.errstate(all="ignore"):
        for a, b in zip(x, y):
            s = float(np.float64(s) + np.float64(a) * np.float64(b))
    return s


def test_vector_dot(orc):
    # util.rs:954-961
    rng = np.random.default_rng(5)
    for _ in range(NCASES):
        n = _sizes(rng, 10)
        x, y = any_f64(rng, n), any_f64(rng, n)
        assert_approx_eq(orc.vector_dot(x, y), _seq_dot(x, y))
    # proptest-regressions/math.txt:12
    x, y = [0.0, 0.0, 0.0, -0.0], [-0.0, 0.0, 0.0, math.inf]
    assert_approx_eq(orc.vector_dot(x, y), _seq_dot(x, y))
    # well-conditioned larger sizes: relative agreement with numpy's pairwise dot
    for n in (4, 16, 17, 100, 4567):
        x, y = rng.normal(size=n), rng.normal(size=n)
        assert abs(orc.vector_dot(x, y) - float(np.dot(x, y))) <= 1e-12 * float(np.dot(np.abs(x), np.abs(y)))


def test_scalar_prods2(orc):
    # util.rs:904-913
    rng = np.random.default_rng(6)
    for _ in range(NCASES):
        n = _sizes(rng, 10)
        x1, x2, y1, y2 = (any_f64(rng, n) for _ in range(4))
        p1, p2 = orc.scalar_prods2(x1, x2, y1, y2)
        with np.errstate(all="ignore"):
            s = x1 + x2
        assert_approx_eq(p1, _seq_dot(s, y1))
        assert_approx_eq(p2, _seq_dot(s, y2))


def test_scalar_prods3(orc):
    # util.rs:915-926
    rng = np.random.default_rng(7)
    for _ in range(NCASES):
        n = _sizes(rng, 10)
        x1, x2, x3, y1, y2 = (any_f64(rng, n) for _ in range(5))
        p1, p2 = orc.scalar_prods3(x1, x2, x3, y1, y2)
        with np.errstate(all="ignore"):
            s = x1 - x2 + x3
        assert_approx_eq(p1, _seq_dot(s, y1))
        assert_approx_eq(p2, _seq_dot(s, y2))
    # proptest-regressions/math.txt:9
    p1, p2 = orc.scalar_prods3([0.0], [0.0], [-4.0946726283401733e139], [0.0], [1.3157422010991668e73])
    assert_approx_eq(p1, -4.0946726283401733e139 * 0.0)
    assert_approx_eq(p2, -4.0946726283401733e139 * 1.3157422010991668e73)


def test_array_update_variance_uses_old_mean(orc):
    # reference src/math/cpu_math.rs:605-631 — NOT textbook Welford: both terms use the old mean
    mean, var = orc.array_update_variance([1.0, 2.0], [0.5, 0.25], [3.0, -2.0], 0.25)
    np.testing.assert_array_equal(mean, [1.0 + 2.0 * 0.25, 2.0 + (-4.0) * 0.25])
    np.testing.assert_array_equal(var, [0.5 + 4.0, 0.25 + 16.0])


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 15, 16, 17, 31, 32, 33])
def test_dot_orders_cover_all_remainders(orc, n):
    # integer-valued inputs: every summation order is exact, so the 4-accumulator path must equal the plain sum
    rng = np.random.default_rng(n)
    x = rng.integers(-50, 50, n).astype(np.float64)
    y = rng.integers(-50, 50, n).astype(np.float64)
    z = rng.integers(-50, 50, n).astype(np.float64)
    assert orc.vector_dot(x, y) == float(np.sum(x * y))
    p1, p2 = orc.scalar_prods3(x, y, z, y, x)
    assert p1 == float(np.sum((x - y + z) * y)) and p2 == float(np.sum((x - y + z) * x))
