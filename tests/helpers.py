"""Shared helpers for the parity tests."""
import struct
from fractions import Fraction

import numpy as np


def f64_bits(x):
    return struct.unpack("<q", struct.pack("<d", float(x)))[0]


def ulps_eq(a, b, max_ulps=32):
    """approx::ulps_eq semantics (the reference's assert_ulps_eq!, src/math/util.rs:759)."""
    a = float(a)
    b = float(b)
    if a == b:
        return True
    if np.isnan(a) or np.isnan(b):
        return False
    if (a < 0) != (b < 0):
        return False
    return abs(f64_bits(a) - f64_bits(b)) <= max_ulps


def assert_approx_eq(a, b, max_ulps=32):
    """reference src/math/util.rs:752-760: NaN on one side matches NaN or inf on the other."""
    a = float(a)
    b = float(b)
    if np.isnan(a) and (np.isnan(b) or np.isinf(b)):
        return
    if np.isnan(b) and (np.isnan(a) or np.isinf(a)):
        return
    assert ulps_eq(a, b, max_ulps), f"{a!r} vs {b!r}"


def exact_fma(a, x, y):
    """f64::mul_add(a, x, y) computed exactly (single rounding)."""
    a, x, y = float(a), float(x), float(y)
    if not (np.isfinite(a) and np.isfinite(x) and np.isfinite(y)):
        with np.errstate(all="ignore"):
            return float(np.float64(a) * np.float64(x) + np.float64(y))
    r = Fraction(a) * Fraction(x) + Fraction(y)
    try:
        return float(r)
    except OverflowError:
        return float("inf") if r > 0 else float("-inf")


def any_f64(rng, n):
    """proptest's prop::num::f64::ANY look-alike: random bit patterns with some special values mixed in."""
    bits = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    vals = bits.view(np.float64).copy()
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 5e-324, 1.7976931348623157e308])
    pick = rng.random(n) < 0.2
    vals[pick] = rng.choice(special, size=int(pick.sum()))
    return vals


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-300)
    return float(np.max(np.abs(a - b) / denom)) if a.size else 0.0
