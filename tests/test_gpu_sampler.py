"""GPU parity, Tier 3: whole draws (Chain::set_position + n x Chain::draw incl. adaptation) of the register-resident
engine against the CPU oracle on the same seeds, through nuts_set_position / nuts_draw.

Tolerance: 1e-9 relative on every draw and float statistic (BASELINE.json north_star), with tree depth, number of
leapfrogs, divergence flags and index_in_trajectory identical.  Each dimension below selects a different
(threads-per-chain, elements-per-thread) instantiation of the engine."""
import numpy as np
import pytest

from nuts_rs_b200 import _abi

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def L():
    from nuts_rs_b200 import lib

    assert lib.device_available(), lib.load().nuts_last_error()
    return lib


def _settings(L, **kw):
    step = kw.pop("step", None)
    s = L.DiagNutsSettings(**kw)
    if step is not None:
        ss = s.adapt_options.step_size_settings
        for k, v in step.items():
            obj = ss
            parts = k.split(".")
            for p in parts[:-1]:
                obj = getattr(obj, p)
            setattr(obj, parts[-1], v)
    return s


def _model_kwargs(kind, d):
    if kind == _abi.NUTS_LOGP_GAUSS_ISO:
        return dict(mu=3.0)
    if kind == _abi.NUTS_LOGP_GAUSS_DIAG:
        return dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, d)))
    if kind == _abi.NUTS_LOGP_GAUSS_RANK1:
        return dict(mu=0.0, rank1_scale=0.5)
    return dict(funnel_scale=3.0)


def compare_run(L, orc, kind, N, d, settings, n_draws, seed=42, x0=None, chain_offset=0, rtol=RTOL, strict=None, min_common=None,
                loose=None):
    """Run the GPU sampler and the oracle on the same seed and compare everything a draw exports, chain by chain.

    Default (`strict=None`): every draw and statistic of the whole run within `rtol` (1e-9) and all tree shapes identical.
    The map draw -> next draw can be chaotic: with adaptation the step size / mass matrix feed back on the 1e-16
    summation-order noise (about a decade of growth per 15 draws in the early windows), the funnel and badly conditioned
    targets amplify it through the dynamics, and once a step size is so small that all leaves weigh the same to rounding
    the reference's own `other.log_size >= self_log_size` shortcut (src/nuts.rs:199) flips on that noise and shifts the RNG
    stream.  For those runs the caller states how long exact agreement is required: the first `strict` draws of every
    chain within `rtol`, tree shapes / divergences / indices identical for at least `min_common` draws, and (optionally)
    floats within `loose` on the rest of the common prefix."""
    kw = _model_kwargs(kind, d)
    if x0 is None:
        x0 = np.random.default_rng(seed).normal(size=(N, d))
    math = L.CudaMath(N, d, kind, **kw)
    s = L.Sampler(math, settings, seed=seed, chain_id_offset=chain_offset)
    st = s.set_position(x0)
    om = orc.Model(kind, d, **kw)
    osamp = orc.Sampler(om, settings, seed=seed, nchains=N, chain_id_offset=chain_offset, nthreads=8)
    ost = osamp.set_position(x0)
    np.testing.assert_array_equal(st, ost)
    g0, o0 = s.state(), osamp.state()
    np.testing.assert_allclose(g0["step_size"], o0["step_size"], rtol=rtol)
    np.testing.assert_allclose(g0["stds"], o0["stds"], rtol=rtol)
    np.testing.assert_allclose(g0["mean"], o0["mean"], rtol=rtol, atol=1e-300)
    np.testing.assert_array_equal(g0["rng_counter"], o0["rng_counter"])
    draws, stats = s.draw(n_draws)
    odraws, ostats = osamp.draw(n_draws)
    if strict is None:
        strict, min_common = n_draws, n_draws
    discrete = ("depth", "n_steps", "diverging", "maxdepth_reached", "index_in_trajectory", "tuning")
    floats = ("logp", "energy", "energy_error", "step_size", "step_size_bar", "mean_tree_accept", "mean_tree_accept_sym",
              "max_energy_error", "fisher_distance")

    def check(c, sl, tol):
        if sl.stop <= sl.start:
            return
        scale = np.maximum(1.0, np.abs(odraws[sl, c]))
        err = np.max(np.abs(draws[sl, c] - odraws[sl, c]) / scale)
        assert err < tol, f"chain {c}: draws differ by {err} (tol {tol})"
        for name in floats:
            a, b = stats[name][sl, c], ostats[name][sl, c]
            fin = np.isfinite(b)
            np.testing.assert_array_equal(np.isfinite(a), fin, err_msg=name)
            t = tol * np.maximum(1.0, np.abs(b[fin]))
            if name in ("energy_error", "max_energy_error", "fisher_distance"):
                # differences / squares of O(d) quantities: scale by the magnitude of the energies involved
                t = t * np.maximum(1.0, np.abs(ostats["energy"][sl, c][fin])) * 10
            assert (np.abs(a[fin] - b[fin]) <= t).all(), (c, name, float(np.max(np.abs(a[fin] - b[fin]))))

    all_common = True
    for c in np.nonzero(st == 0)[0]:
        same = np.ones(n_draws, dtype=bool)
        for name in discrete:
            same &= stats[name][:, c] == ostats[name][:, c]
        n_common = n_draws if same.all() else int(np.argmin(same))
        all_common &= n_common == n_draws
        assert n_common >= min_common, f"chain {c}: tree shapes differ from draw {n_common} on (need {min_common})"
        k = min(strict, n_common)
        check(c, slice(0, k), rtol)
        if loose is not None:
            check(c, slice(k, n_common), loose)
    if all_common:
        g1, o1 = s.state(), osamp.state()
        np.testing.assert_array_equal(g1["rng_counter"][st == 0], o1["rng_counter"][st == 0])
    total, done = s.counters()
    assert done == n_draws
    s.close()
    math.close()
    return stats, ostats


def _no_adapt(L, **kw):
    """num_tune = 0: step size (from the initial search) and mass matrix (from the first gradient) stay fixed for the whole run —
    the north-star condition "identical RNG seed and step size"."""
    return _settings(L, num_tune=0, **kw)


def test_c1_reference_bench_shape(L, orc):
    """BASELINE config 1: 10-dim isotropic Gaussian, 4 chains, maxdepth=3, num_tune=1000, start 3.5 (benches/sample.rs:76-98)."""
    s = _settings(L, num_tune=1000, maxdepth=3)
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_ISO, 4, 10, s, 1000, x0=np.full((4, 10), 3.5), strict=15, min_common=60, loose=1e-3)


def test_c1_fixed_adaptation_whole_run(L, orc):
    """Config 1 without adaptation: 1000 draws x 4 chains, every draw within 1e-9 of the oracle."""
    s = _no_adapt(L, maxdepth=3)
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_ISO, 4, 10, s, 1000, x0=np.full((4, 10), 3.5))
    s = _no_adapt(L)
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_ISO, 4, 10, s, 300, x0=np.full((4, 10), 0.5 + 1.0))


def test_c1_default_depth_with_adaptation(L, orc):
    s = _settings(L, num_tune=300)
    stats, _ = compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_ISO, 6, 10, s, 500, strict=15, min_common=50, loose=1e-3)
    assert not stats["tuning"][300:].any() and stats["tuning"][:300].all()


ENGINE_DIMS = [(1, 4, 60), (2, 4, 60), (33, 5, 60), (64, 3, 60), (100, 6, 80), (129, 3, 60), (256, 3, 60), (400, 3, 50), (1000, 4, 50),
               (1025, 2, 40), (2048, 2, 30), (3000, 2, 30), (5000, 2, 24), (7000, 3, 16), (9000, 3, 16), (10000, 2, 20), (12000, 2, 16)]


@pytest.mark.parametrize("d,N,draws", ENGINE_DIMS)
def test_every_engine_configuration_trajectories(L, orc, d, N, draws):
    """One run per register tiling of the engine (32x1 ... 1024x16): diagonal Gaussian, fixed step size and mass matrix,
    every draw of the run within 1e-9 of the oracle with identical tree shapes."""
    s = _no_adapt(L, maxdepth=6)
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, s, draws, seed=d)


@pytest.mark.parametrize("d,N,draws", ENGINE_DIMS)
def test_every_engine_configuration_with_adaptation(L, orc, d, N, draws):
    """Same tilings with warmup inside the run (estimator updates, mass-matrix switches, dual averaging, step-size re-search)."""
    s = _settings(L, num_tune=draws // 2, maxdepth=6)
    # The first mass-matrix update (draw 1) divides 3-sample variances: a coordinate whose three samples nearly coincide loses
    # digits by cancellation, and the worst coordinate out of d gets worse as d grows.  So: draw 0 (before any update) within
    # 1e-9, tree shapes identical and floats within 1e-6 for the first 5 draws; test_short_schedule_full_adaptation_cycle and
    # the C1 tests follow the adaptation for 45 - 60 draws at small d.
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, s, draws, seed=d, strict=1, min_common=5, loose=1e-6)


@pytest.mark.parametrize("kind", [_abi.NUTS_LOGP_GAUSS_RANK1, _abi.NUTS_LOGP_GAUSS_ISO, _abi.NUTS_LOGP_GAUSS_DIAG])
def test_models_100d_fixed_adaptation(L, orc, kind):
    s = _no_adapt(L, maxdepth=8)
    if kind == _abi.NUTS_LOGP_GAUSS_RANK1:
        # condition number 51 with the un-adapted initial mass matrix: chains whose initial step size sits near the stability
        # limit amplify rounding noise exponentially, so exact agreement is required on a prefix only
        compare_run(L, orc, kind, 8, 100, s, 150, seed=3, strict=25, min_common=60)
    else:
        compare_run(L, orc, kind, 8, 100, s, 150, seed=3)


@pytest.mark.parametrize("kind", [_abi.NUTS_LOGP_GAUSS_RANK1, _abi.NUTS_LOGP_GAUSS_ISO])
def test_models_100d_with_adaptation(L, orc, kind):
    s = _settings(L, num_tune=100, maxdepth=8)
    compare_run(L, orc, kind, 8, 100, s, 160, seed=3, strict=12, min_common=30)


def test_funnel_divergences_match(L, orc):
    """Neal's funnel (BASELINE config 3 target): divergent leapfrogs, depth spread; chaotic, so a shorter strict prefix."""
    s = _settings(L, num_tune=100, maxdepth=8)
    stats, ostats = compare_run(L, orc, _abi.NUTS_LOGP_FUNNEL, 16, 10, s, 200, seed=5, strict=8, min_common=20, rtol=1e-8)
    assert stats["diverging"].sum() > 0
    s = _no_adapt(L, maxdepth=8)
    compare_run(L, orc, _abi.NUTS_LOGP_FUNNEL, 16, 10, s, 60, seed=6, strict=10, min_common=25, rtol=1e-8)


def test_fixed_step_no_turn_checks_exact_leapfrog_count(L, orc):
    """mindepth = maxdepth disables the U-turn checks: exactly 2^maxdepth - 1 leapfrogs per draw, fixed step, no jitter."""
    s = _settings(L, num_tune=0, maxdepth=5, mindepth=5,
                  step={"adapt_options.method": _abi.NUTS_STEPSIZE_FIXED, "adapt_options.fixed_step": 0.05, "has_jitter": 0})
    stats, _ = compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, 8, 50, s, 40, seed=8)
    assert (stats["n_steps"] == 31).all() and (stats["depth"] == 5).all() and stats["maxdepth_reached"].all()
    assert (stats["step_size"] == 0.05).all()


def test_deep_trees_maxdepth_10(L, orc):
    """A small fixed step forces deep trees (up to 2^10 - 1 leapfrogs): exercises every level of the checkpoint pool."""
    s = _settings(L, num_tune=0, maxdepth=10,
                  step={"adapt_options.method": _abi.NUTS_STEPSIZE_FIXED, "adapt_options.fixed_step": 0.004, "has_jitter": 1})
    stats, _ = compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, 4, 40, s, 12, seed=12)
    assert stats["depth"].max() >= 9


def test_extra_doublings_and_target_time(L, orc):
    s = _no_adapt(L, maxdepth=6, extra_doublings=1)
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_ISO, 4, 20, s, 80, seed=9)
    s = _no_adapt(L, maxdepth=7, has_target_integration_time=1, target_integration_time=3.0)
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_ISO, 4, 20, s, 80, seed=10)


def test_short_schedule_full_adaptation_cycle(L, orc):
    """num_tune = 30: early window (switches every 10 draws), first mass-matrix change + step-size re-search, main window,
    final step-size window, last tuning draw (best-guess step) and 15 post-tuning draws all inside the exact-agreement prefix."""
    s = _settings(L, num_tune=30, maxdepth=5)
    stats, _ = compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_ISO, 8, 10, s, 45, seed=21, strict=12, min_common=45, loose=1e-4)
    assert stats["tuning"][:30].all() and not stats["tuning"][30:].any()


def test_draw_variance_estimator_option(L, orc):
    s = _settings(L, num_tune=80, maxdepth=6)
    s.adapt_options.mass_matrix_options.use_grad_based_estimate = 0
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, 4, 30, s, 120, seed=11, strict=12, min_common=40)


def test_bad_initial_points_are_reported_and_skipped(L, orc):
    N, d = 4, 10
    x0 = np.random.default_rng(0).normal(size=(N, d))
    x0[1, 3] = np.nan
    x0[2, :] = 3.0  # exactly at the mode: zero gradient => BadInitGrad (reference transformed_hamiltonian.rs:314)
    s = _settings(L, num_tune=20, maxdepth=4)
    math = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=3.0)
    samp = L.Sampler(math, s, seed=1)
    st = samp.set_position(x0)
    om = orc.Model(_abi.NUTS_LOGP_GAUSS_ISO, d, mu=3.0)
    osamp = orc.Sampler(om, s, seed=1, nchains=N)
    np.testing.assert_array_equal(st, osamp.set_position(x0))
    np.testing.assert_array_equal(st, [0, 3, 3, 0])
    draws, stats = samp.draw(10)
    odraws, _ = osamp.draw(10)
    assert np.isnan(draws[:, 1]).all() and np.isnan(draws[:, 2]).all()
    np.testing.assert_allclose(draws[:, [0, 3]], odraws[:, [0, 3]], rtol=1e-9)
    samp.close()
    math.close()


def test_chain_offset_invariance_and_split_draw_calls(L, orc):
    """Streams are keyed by the GLOBAL chain id (reference set_stream(chain_id+1), src/sampler.rs:1106): a shard that starts
    at chain_id_offset=k reproduces chains k.. of the unsharded run; and draw(a)+draw(b) == draw(a+b)."""
    N, d = 6, 20
    s = _settings(L, num_tune=30, maxdepth=5)
    x0 = np.random.default_rng(1).normal(size=(N, d))
    kw = _model_kwargs(_abi.NUTS_LOGP_GAUSS_DIAG, d)
    m_all = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, **kw)
    s_all = L.Sampler(m_all, s, seed=77)
    s_all.set_position(x0)
    d_all, st_all = s_all.draw(50)
    m_sh = L.CudaMath(3, d, _abi.NUTS_LOGP_GAUSS_DIAG, **kw)
    s_sh = L.Sampler(m_sh, s, seed=77, chain_id_offset=3)
    s_sh.set_position(x0[3:])
    d1, st1 = s_sh.draw(20)
    d2, st2 = s_sh.draw(30)
    np.testing.assert_array_equal(np.concatenate([d1, d2]), d_all[:, 3:])
    np.testing.assert_array_equal(np.concatenate([st1["n_steps"], st2["n_steps"]]), st_all["n_steps"][:, 3:])
    for x in (s_all, s_sh):
        x.close()
    for x in (m_all, m_sh):
        x.close()


def test_draws_written_directly_into_pinned_host_memory(L):
    """nuts_draw picks the output path from where draws_out lives (include/nuts_b200.h): page-locked host memory is written
    by the draw kernel itself, pageable memory is staged.  Both must deliver bit-identical draws (≙ Chain::draw's position,
    src/chain.rs:169-171) and a dead chain's missing draws must read NaN on both."""
    N, d, n = 24, 100, 12
    x0 = np.random.default_rng(3).normal(size=(N, d))
    x0[5, 7] = np.inf  # BadInitGrad for chain 5 (transformed_hamiltonian.rs:678-681): it never produces a draw
    outs = []
    for pinned in (False, True):
        m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, **_model_kwargs(_abi.NUTS_LOGP_GAUSS_DIAG, d))
        s = L.Sampler(m, _settings(L, num_tune=6, maxdepth=6), seed=11)
        st = s.set_position(x0)
        assert st[5] == 3 and (np.delete(st, 5) == 0).all()
        buf = L.HostBuffer((n, N, d)) if pinned else None
        draws, stats = s.draw(n, out=buf.array if pinned else None)
        assert s.last_draw_direct() == pinned
        outs.append((draws.copy(), stats))
        s.close()
        m.close()
        if buf is not None:
            buf.close()
    (a, sa), (b, sb) = outs
    assert np.isnan(a[:, 5]).all() and np.isnan(b[:, 5]).all()
    live = np.delete(np.arange(N), 5)
    assert np.array_equal(a[:, live], b[:, live])
    for k in sa:
        assert np.array_equal(sa[k][:, live], sb[k][:, live]), k


def test_draw_trace_schema_on_device(L):
    """Sampler.draw_trace: the engine's output in the reference's trace layout ([chain, draw], src/storage/core.rs:12-77),
    `draw` continuing across calls and `chain` carrying the global chain id of a shard (src/sampler.rs:165-174)."""
    N, d = 6, 20
    m = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=3.0)
    s = L.Sampler(m, _settings(L, num_tune=5, maxdepth=5), seed=3, chain_id_offset=100)
    assert (s.set_position(np.full((N, d), 3.5)) == 0).all()
    a = s.draw_trace(7)
    b = s.draw_trace(4)
    s.close()
    m.close()
    assert a["posterior"]["unconstrained_draw"].shape == (N, 7, d)
    assert (a["sample_stats"]["chain"][:, 0] == np.arange(100, 100 + N)).all()
    assert (a["sample_stats"]["draw"][0] == np.arange(7)).all() and (b["sample_stats"]["draw"][0] == np.arange(7, 11)).all()
    assert a["sample_stats"]["tuning"][:, :5].all() and not a["sample_stats"]["tuning"][:, 5:].any() and not b["sample_stats"]["tuning"].any()
    assert (a["sample_stats"]["n_steps"] >= 1).all() and np.isfinite(a["sample_stats"]["logp"]).all()


def test_bad_initial_points_are_retried_like_the_reference_sampler(L, orc):
    """src/sampler.rs:1133-1143: a chain whose set_position fails asks for a fresh init_position (up to 500 times).  The chains
    that were fine keep their state bit for bit; a retried chain starts over as a new chain on its own random stream."""
    N, d = 6, 10
    s = _settings(L, num_tune=20, maxdepth=4)
    rng = np.random.default_rng(4)
    good = rng.normal(size=(N, d))
    calls = []

    def init_position(chain_ids):
        calls.append(np.array(chain_ids))
        x = good[chain_ids].copy()
        if len(calls) == 1:
            x[1, 3] = np.nan        # chain 1: non-finite start
            x[4, :] = 3.0           # chain 4: exactly at the mode (zero gradient => BadInitGrad)
        elif len(calls) == 2:
            x[list(chain_ids).index(4), :] = 3.0  # chain 4 fails a second time
        return x

    math = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=3.0)
    samp = L.Sampler(math, s, seed=1)
    status, tries = samp.set_position_with_retries(init_position)
    assert (status == 0).all() and tries == 3
    assert [list(c) for c in calls] == [[0, 1, 2, 3, 4, 5], [1, 4], [4]]
    draws, stats = samp.draw(30)
    assert np.isfinite(draws).all()
    # chains 0, 2, 3, 5 never failed: identical to a run without any bad point
    m2 = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_ISO, mu=3.0)
    s2 = L.Sampler(m2, s, seed=1)
    assert (s2.set_position(good) == 0).all()
    d2, _ = s2.draw(30)
    ok = [0, 2, 3, 5]
    assert np.array_equal(draws[:, ok], d2[:, ok])
    # a chain that exhausts its tries stays dead (status 3) and produces NaN draws
    samp2 = L.Sampler(m2, s, seed=1)
    status, tries = samp2.set_position_with_retries(lambda ids: np.where(np.asarray(ids)[:, None] == 2, 3.0, good[ids]), max_tries=4)
    assert tries == 4 and status[2] == 3 and (np.delete(status, 2) == 0).all()
    d3, _ = samp2.draw(5)
    assert np.isnan(d3[:, 2]).all() and np.isfinite(np.delete(d3, 2, axis=1)).all()
    for x in (samp, s2, samp2):
        x.close()
    for x in (math, m2):
        x.close()


LARGE_DIMS = [(5000, 2, 16), (9000, 2, 12), (10240, 2, 10)]


@pytest.mark.parametrize("d,N,draws", LARGE_DIMS)
@pytest.mark.parametrize("kind", [_abi.NUTS_LOGP_GAUSS_RANK1, _abi.NUTS_LOGP_FUNNEL])
def test_large_dim_cluster_engine_non_elementwise_targets(L, orc, kind, d, N, draws):
    """dim ~ 10^4 for the targets that need a team-wide reduction inside every leapfrog (rank-1 Gaussian, funnel): by default the
    cluster engine (the team = the 4 CTAs of a thread-block cluster, reductions through distributed shared memory); fixed step
    size and mass matrix, every draw within 1e-9 of the oracle with identical tree shapes."""
    s = _no_adapt(L, maxdepth=5)
    if kind == _abi.NUTS_LOGP_FUNNEL:
        # the funnel's neck is chaotic: start every chain inside the bulk (v = 0) and require exact agreement on a prefix
        x0 = np.random.default_rng(d).normal(size=(N, d))
        x0[:, 0] = 0.1
        compare_run(L, orc, kind, N, d, s, draws, seed=d, x0=x0, strict=4, min_common=6, rtol=1e-8)
    else:
        compare_run(L, orc, kind, N, d, s, draws, seed=d, strict=6, min_common=8)


@pytest.mark.parametrize("d,N,draws", LARGE_DIMS[:2])
def test_large_dim_cluster_engine_diag_with_adaptation(L, orc, monkeypatch, d, N, draws):
    """The cluster engine on the elementwise target (selected explicitly; the decoupled engine is the default there), warm-up on."""
    monkeypatch.setenv("NUTS_B200_ENGINE", "1024,10,41")
    s = _settings(L, num_tune=draws // 2, maxdepth=6)
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, s, draws, seed=d, strict=1, min_common=5, loose=1e-6)
    s = _no_adapt(L, maxdepth=6)
    compare_run(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, s, draws, seed=d + 1)


@pytest.mark.parametrize("kind,d,N", [(_abi.NUTS_LOGP_GAUSS_ISO, 10, 6), (_abi.NUTS_LOGP_GAUSS_DIAG, 1000, 4), (_abi.NUTS_LOGP_GAUSS_RANK1, 100, 5)])
def test_adam_step_size_adaptation_matches_the_oracle(L, orc, kind, d, N):
    """StepSizeAdaptMethod::Adam (src/stepsize/adam.rs) instead of dual averaging, mass-matrix adaptation on: exact-agreement prefix
    like every adaptive run, step size and step_size_bar included; the optimizer state travels in the dual-averaging record of the
    chain state (log_step, m, v, t)."""
    s = _settings(L, num_tune=60, maxdepth=6, step={"adapt_options.method": _abi.NUTS_STEPSIZE_ADAM})
    stats, ostats = compare_run(L, orc, kind, N, d, s, 80, seed=21, strict=4, min_common=10, loose=1e-6)
    assert stats["tuning"][:60].all() and not stats["tuning"][60:].any()
    # after the warm-up the step size is exp(log_step) of the optimizer (jittered), not a dual-averaging average
    np.testing.assert_allclose(stats["step_size_bar"][:4], ostats["step_size_bar"][:4], rtol=1e-9)
