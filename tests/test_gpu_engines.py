"""GPU: the alternative engines of the same tiling must agree BIT FOR BIT on the same seeds.

64x16 (one CTA of 64 threads per chain, 16 elements per thread) exists in three builds:
  64,16,4    register-resident engine, bounds-checked rows                          (chain_engine.cuh)
  64,16,54   the same with SM_EXACT (rows zero-padded to 1024, no bounds checks)    - the default for 953 <= dim <= 1024
  64,16,107  decoupled engine: leader warp = scalar tree logic, teams = vector work (chain_engine_v2.cuh)
All three keep the per-thread partial sums, the warp reduce-scatter and the order of the per-warp partials identical, and the
decoupled engine consumes the random stream in the reference's order (src/nuts.rs:189-203, 334), so draws, statistics and the
leapfrog count are identical - which also pins the hand-over protocol of the decoupled engine (run-ahead, aborts, end slots)."""
import os

import numpy as np
import pytest

from nuts_rs_b200 import _abi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    from nuts_rs_b200 import lib

    assert lib.device_available(), lib.load().nuts_last_error()
    return lib


def _run(L, engine, N, d, num_tune, n_draws, maxdepth, seed=7, **kw):
    old = os.environ.get("NUTS_B200_ENGINE")
    os.environ["NUTS_B200_ENGINE"] = engine
    try:
        m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=np.exp(np.linspace(-1, 1, d)))
        s = L.Sampler(m, L.DiagNutsSettings(num_tune=num_tune, maxdepth=maxdepth, **kw), seed=seed)
        status = s.set_position(np.random.default_rng(1).normal(size=(N, d)))
        draws, stats = s.draw(n_draws)
        lf, _ = s.counters()
        s.close()
        m.close()
        return status, draws, stats, lf
    finally:
        if old is None:
            os.environ.pop("NUTS_B200_ENGINE", None)
        else:
            os.environ["NUTS_B200_ENGINE"] = old


CASES = [
    # N, d, num_tune, n_draws, maxdepth, extra settings
    (8, 1000, 0, 3, 3, {}),
    (64, 1000, 30, 40, 8, {}),
    (300, 700, 60, 30, 10, {}),          # adaptation incl. mass-matrix switches, deep trees
    (1100, 1000, 20, 10, 6, {}),         # more chains than resident teams: draws migrate between teams / CTAs
    (40, 900, 10, 12, 5, {"extra_doublings": 2}),   # doublings after the U-turn continue from the unchanged tree ends
    (40, 520, 5, 12, 6, {"mindepth": 3}),           # no turn checks below mindepth
]


@pytest.mark.parametrize("engine", ["64,16,54", "64,16,107"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: f"N{c[0]}_d{c[1]}_tune{c[2]}_depth{c[4]}" + ("_" + "_".join(c[5]) if c[5] else ""))
def test_engine_variants_bit_identical(L, engine, case):
    N, d, tune, n, md, kw = case
    ref = _run(L, "64,16,4", N, d, tune, n, md, **kw)
    got = _run(L, engine, N, d, tune, n, md, **kw)
    assert np.array_equal(ref[0], got[0])
    assert ref[3] == got[3], "leapfrog counters differ"
    assert np.array_equal(ref[1], got[1], equal_nan=True), "draws differ"
    for k in ref[2]:
        assert np.array_equal(ref[2][k], got[2][k], equal_nan=True), k
    assert ref[2]["depth"].max() >= min(md, 3)


def _run_aligned(L, align, kind, N, d, num_tune, n_draws, maxdepth, **mk):
    old = os.environ.get("NUTS_B200_ALIGN")
    os.environ["NUTS_B200_ALIGN"] = align
    try:
        m = L.CudaMath(N, d, kind, **mk)
        s = L.Sampler(m, L.DiagNutsSettings(num_tune=num_tune, maxdepth=maxdepth), seed=11)
        x0 = np.random.default_rng(2).normal(size=(N, d))
        if kind == _abi.NUTS_LOGP_FUNNEL:
            x0[:, 0] = 0.1
        status = s.set_position(x0)
        draws, stats = s.draw(n_draws)
        lf, _ = s.counters()
        s.close()
        m.close()
        return status, draws, stats, lf
    finally:
        if old is None:
            os.environ.pop("NUTS_B200_ALIGN", None)
        else:
            os.environ["NUTS_B200_ALIGN"] = old


@pytest.mark.parametrize("kind,d,N", [
    (_abi.NUTS_LOGP_GAUSS_DIAG, 20, 5000), (_abi.NUTS_LOGP_GAUSS_DIAG, 50, 3000), (_abi.NUTS_LOGP_GAUSS_RANK1, 100, 3000),
    (_abi.NUTS_LOGP_GAUSS_DIAG, 200, 2000), (_abi.NUTS_LOGP_GAUSS_DIAG, 500, 1400), (_abi.NUTS_LOGP_FUNNEL, 10, 5000),
    (_abi.NUTS_LOGP_GAUSS_DIAG, 100, 7),
])
def test_aligned_warp_tilings_bit_identical(L, kind, d, N):
    """The warp tilings exist in two builds: one warp team per CTA (tag 16 / 12 / 8), and all teams of an SM in one CTA that starts
    its work units together (SM_ALIGN, tag 21: the default, NUTS_B200_ALIGN=0 selects the other).  Which team runs which draw
    when does not enter the arithmetic: draws, statistics and leapfrog counts agree bit for bit - with more chains than resident
    teams (hand-over through the ready queue, teams that run out of work while their CTA keeps going) and with fewer."""
    mk = {_abi.NUTS_LOGP_GAUSS_DIAG: dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, d))),
          _abi.NUTS_LOGP_GAUSS_RANK1: dict(mu=0.0, rank1_scale=0.5), _abi.NUTS_LOGP_FUNNEL: dict(funnel_scale=3.0)}[kind]
    ref = _run_aligned(L, "0", kind, N, d, 12, 20, 6, **mk)
    got = _run_aligned(L, "1", kind, N, d, 12, 20, 6, **mk)
    assert np.array_equal(ref[0], got[0])
    assert ref[3] == got[3], "leapfrog counters differ"
    assert np.array_equal(ref[1], got[1], equal_nan=True), "draws differ"
    for k in ref[2]:
        assert np.array_equal(ref[2][k], got[2][k], equal_nan=True), k


def test_warm_up_build_is_bit_identical(L):
    """dim ~ 1000 with more chains than resident teams: launches inside the warm-up run the aligned build of the exact 64x16 tiling
    (four teams per CTA, named barriers, two draws per unit), everything else the plain one; NUTS_B200_TUNE_ENGINE=0 keeps the plain
    build throughout.  Same arithmetic: draws, statistics and leapfrog counts agree bit for bit, across the hand-over between the
    builds in the middle of the run."""
    N, d = 1300, 1000

    def run(flag):
        old = os.environ.get("NUTS_B200_TUNE_ENGINE")
        os.environ["NUTS_B200_TUNE_ENGINE"] = flag
        try:
            m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=np.exp(np.linspace(-1, 1, d)))
            s = L.Sampler(m, L.DiagNutsSettings(num_tune=24, maxdepth=6), seed=7)
            status = s.set_position(np.random.default_rng(1).normal(size=(N, d)))
            parts = [s.draw(k) for k in (10, 14, 3, 9)]  # two warm-up launches (aligned build), then sampling launches
            lf, _ = s.counters()
            s.close()
            m.close()
            draws = np.concatenate([p[0] for p in parts])
            stats = {k: np.concatenate([p[1][k] for p in parts]) for k in parts[0][1]}
            return status, draws, stats, lf
        finally:
            if old is None:
                os.environ.pop("NUTS_B200_TUNE_ENGINE", None)
            else:
                os.environ["NUTS_B200_TUNE_ENGINE"] = old

    ref, got = run("0"), run("1")
    assert np.array_equal(ref[0], got[0]) and ref[3] == got[3]
    assert np.array_equal(ref[1], got[1], equal_nan=True), "draws differ"
    for k in ref[2]:
        assert np.array_equal(ref[2][k], got[2][k], equal_nan=True), k
    assert ref[2]["tuning"][:24].all() and not ref[2]["tuning"][24:].any()
