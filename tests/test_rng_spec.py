"""The project's random-stream specification (oracle/rng_spec.hpp): Philox4x32-10 known answers
(Random123 kat_vectors), accuracy of the deterministic log / sincos, and distributional checks."""
import math

import numpy as np
from scipy import stats


def test_philox_known_answers(orc):
    assert orc.philox(0, 0, 0) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    m = 0xFFFFFFFFFFFFFFFF
    assert orc.philox(m, m, m) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    # counter = 243f6a88 85a308d3 13198a2e 03707344, key = a4093822 299f31d0
    assert orc.philox(0x299F31D0A4093822, 0x0370734413198A2E, 0x85A308D3243F6A88) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_det_log_accuracy(orc):
    rng = np.random.default_rng(0)
    us = np.concatenate([rng.random(2000), 2.0 ** -rng.integers(1, 53, 200), [1.0, 2.0**-53, 0.5, 0.7071067811865476]])
    for u in us:
        got = orc.lib().orc_det_log(float(u))
        want = math.log(u)
        assert abs(got - want) <= 4e-16 * max(1.0, abs(want)), (u, got, want)


def test_det_sincos_accuracy(orc):
    rng = np.random.default_rng(1)
    for u in np.concatenate([rng.random(2000), [0.0, 0.125, 0.25, 0.375, 0.5, 0.625, 0.75, 0.875, 1 - 2.0**-53]]):
        s, c = orc.det_sincos2pi(float(u))
        assert abs(s - math.sin(2 * math.pi * u)) < 2e-15  # the libm reference itself carries the rounding of 2*pi*u
        assert abs(c - math.cos(2 * math.pi * u)) < 2e-15  # the libm reference itself carries the rounding of 2*pi*u


def test_normal_stream_distribution(orc):
    x, ctr = orc.fill_normal(42, 1, 0, 200001)
    assert ctr == 100001  # ceil(d/2) counters consumed
    assert abs(x.mean()) < 0.01 and abs(x.std() - 1) < 0.01
    assert stats.kstest(x, "norm").pvalue > 1e-3
    # streams are separated and reproducible
    y, _ = orc.fill_normal(42, 2, 0, 1000)
    z, _ = orc.fill_normal(42, 1, 0, 1000)
    np.testing.assert_array_equal(z, x[:1000])
    assert abs(np.corrcoef(y, z)[0, 1]) < 0.15
    # element i depends only on counter + i//2: a fill starting one pair later is the same stream shifted
    w, _ = orc.fill_normal(42, 1, 1, 10)
    np.testing.assert_array_equal(w, x[2:12])


def test_uniform_and_bool(orc):
    L = orc.lib()
    u = np.array([L.orc_next_f64(7, 3, c) for c in range(20000)])
    assert 0.0 <= u.min() and u.max() < 1.0
    assert stats.kstest(u, "uniform").pvalue > 1e-3
    b = np.array([L.orc_next_bool(7, 3, c) for c in range(20000)])
    assert abs(b.mean() - 0.5) < 0.02
    j = np.array([L.orc_uniform(7, 3, c, 0.9, 1.1) for c in range(2000)])
    assert 0.9 <= j.min() and j.max() < 1.1
