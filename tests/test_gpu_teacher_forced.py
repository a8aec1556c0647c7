"""GPU parity with adaptation ON, one draw ahead ("teacher forced").

The map draw -> next draw is chaotic once the step size and the mass matrix feed back on 1e-16 summation-order noise, so a
free-running comparison can only demand agreement on a prefix (tests/test_gpu_sampler.py).  Here the oracle is re-seeded
with the GPU's COMPLETE chain state (nuts_sampler_get_chain_state: point, mass matrix, step size + dual averaging, the four
running-variance estimators, window schedule, random-stream position) after every draw t and must then reproduce draw t+1:

  * tree depth, leapfrog count, divergence flag, index_in_trajectory, tuning flag, random numbers consumed: IDENTICAL
  * the draw and every float statistic: within 1e-9 relative (BASELINE.json north_star)
  * the state AFTER the draw (GlobalStrategy::adapt, reference src/adapt_strategy.rs:121-222: estimator updates
    src/transform/adapt/diagonal.rs:134-148, mass-matrix update src/transform/diagonal.rs:107-131, dual averaging
    src/stepsize/dual_avg.rs:55-63, step-size re-search src/stepsize/adapt.rs:91-199, window counters): counters and flags
    identical, floats within 1e-9 of their natural scale

for EVERY draw of the whole warm-up schedule (early windows, window switches, the first mass-matrix change with its step-size
re-search, the final step-size window, the last tuning draw) and a stretch of post-warm-up draws, on the BASELINE targets:
config 2 (d = 1000, the 64x16 exact tile), config 4 (d = 10^4, the large-dim engine), the funnel and the rank-1 Gaussian.
No amplification from draw to draw means a single wrong branch anywhere in the adaptation shows up as a hard failure."""
import numpy as np
import pytest

from nuts_rs_b200 import _abi

pytestmark = pytest.mark.gpu

RTOL = 1e-9
DISCRETE = ("depth", "n_steps", "diverging", "maxdepth_reached", "index_in_trajectory", "tuning")
FLOATS = ("logp", "energy", "energy_error", "step_size", "step_size_bar", "mean_tree_accept", "mean_tree_accept_sym", "max_energy_error",
          "fisher_distance")
STATE_EXACT = ("point_transform_id", "mass_matrix_id", "da_count", "foreground_count", "background_count", "tuning",
               "has_initial_mass_matrix", "last_update", "current_window_size", "draw_count", "rng_counter", "total_leapfrogs", "alive")
STATE_SCALARS = ("logp", "step_size", "da_log_step", "da_log_step_adapted", "da_hbar", "da_mu")
STATE_VECTORS = ("position", "gradient", "stds", "inv_stds", "mean", "draw_mean", "grad_mean", "draw_mean_bg", "grad_mean_bg")
STATE_VARIANCES = (("draw_var", "position"), ("grad_var", "gradient"), ("draw_var_bg", "position"), ("grad_var_bg", "gradient"))


@pytest.fixture(scope="module")
def L():
    from nuts_rs_b200 import lib

    assert lib.device_available(), lib.load().nuts_last_error()
    return lib


def _close(a, b, tol, scale):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    if not np.array_equal(np.isfinite(a), fin):
        return np.inf
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(a[fin] - b[fin]) / (tol * np.broadcast_to(scale, a.shape)[fin])))


def teacher_forced(L, orc, kind, N, d, settings, n_draws, seed, model_kwargs, rtol=RTOL, x0=None, max_knife_edge=0, oracle_chains=None,
                   skip_draws=0):
    """Returns the per-draw statistics of the GPU run.  `max_knife_edge`: (chain, draw) pairs that may differ in a discrete
    outcome - a U-turn product or a multinomial weight within rounding of its threshold; they are re-synchronised by the next
    draw's state injection and never allowed in the adaptation state itself.
    `oracle_chains`: the GPU runs all N chains (a BASELINE config at full size), the oracle follows the first `oracle_chains` of
    them (chains are independent and their random streams are keyed by the chain id).  `skip_draws`: GPU draws made in one call
    before the comparison starts (to reach a later part of the schedule at full size)."""
    if x0 is None:
        x0 = np.random.default_rng(seed).normal(size=(N, d))
    n_or = N if oracle_chains is None else oracle_chains
    math = L.CudaMath(N, d, kind, **model_kwargs)
    S = L.Sampler(math, settings, seed=seed)
    st = S.set_position(x0)
    om = orc.Model(kind, d, **model_kwargs)
    osamp = orc.Sampler(om, settings, seed=seed, nchains=n_or, nthreads=min(n_or, 8))
    np.testing.assert_array_equal(st[:n_or], osamp.set_position(x0[:n_or]))
    live = st[:n_or] == 0
    assert live.any()

    def gpu_state():
        return {k: v[:n_or] for k, v in S.chain_state().items()}

    state = gpu_state()
    # the state after Chain::set_position agrees before anything is injected (initial mass matrix, step-size search)
    o0 = osamp.chain_state()
    for k in STATE_EXACT:
        np.testing.assert_array_equal(state[k][live], o0[k][live], err_msg=k)
    np.testing.assert_allclose(state["step_size"][live], o0["step_size"][live], rtol=rtol)
    np.testing.assert_allclose(state["stds"][live], o0["stds"][live], rtol=rtol)
    knife_edge = []
    all_stats = {k: [] for k in DISCRETE + FLOATS}
    worst = 0.0
    if skip_draws:
        S.draw(skip_draws, want_draws=False, stats=False)
        state = gpu_state()
    for t in range(skip_draws, skip_draws + n_draws):
        gd, gs = S.draw(1)
        gd, gs = gd[:, :n_or], {k: v[:, :n_or] for k, v in gs.items()}
        osamp.set_chain_state(state)
        od, os_ = osamp.draw(1)
        post, opost = gpu_state(), osamp.chain_state()
        for k in all_stats:
            all_stats[k].append(gs[k][0].copy())
        for c in np.nonzero(live)[0]:
            same = all(gs[k][0, c] == os_[k][0, c] for k in DISCRETE) and post["rng_counter"][c] == opost["rng_counter"][c]
            if not same:
                knife_edge.append((t, int(c), {k: (int(gs[k][0, c]), int(os_[k][0, c])) for k in DISCRETE}))
                assert len(knife_edge) <= max_knife_edge, f"discrete outcome differs (draw, chain, gpu/oracle): {knife_edge}"
                continue
            where = f"draw {t} chain {c}"
            e = _close(gd[0, c], od[0, c], rtol, np.maximum(1.0, np.abs(od[0, c])))
            assert e <= 1.0, f"{where}: position off by {e * rtol:.3e} relative"
            worst = max(worst, e)
            escale = max(1.0, abs(os_["energy"][0, c]))
            # Energies are sums of d terms; GPU and oracle add them in different orders and agree to ~1e-13 of their magnitude
            # (1e-9 * 1e-4).  exp(E0 - E) turns that ABSOLUTE error into a relative one, so the acceptance statistics - and what
            # dual averaging derives from them - are only determined to 1e-13 |E|: visible in the first draws of config 4, where
            # |E| ~ 1e9 before the mass matrix has adapted (x0 ~ N(0,1) against sigma = 1e-3).
            ascale = max(1.0, 1e-4 * escale)
            for k in FLOATS:
                scale = max(1.0, abs(os_[k][0, c]))
                if k in ("energy_error", "max_energy_error"):
                    scale = 10 * escale  # differences of O(d) energies
                if k == "fisher_distance":
                    scale = 10 * max(1.0, abs(os_[k][0, c]), escale)
                if k in ("mean_tree_accept", "mean_tree_accept_sym", "step_size", "step_size_bar"):
                    scale = scale * ascale
                e = _close(gs[k][0, c], os_[k][0, c], rtol, scale)
                assert e <= 1.0, f"{where}: statistic {k}: {gs[k][0, c]!r} vs {os_[k][0, c]!r}"
            # ---- the state after the draw: every branch of GlobalStrategy::adapt
            for k in STATE_EXACT:
                assert post[k][c] == opost[k][c], f"{where}: state {k}: {post[k][c]} vs {opost[k][c]}"
            for k in STATE_SCALARS:
                e = _close(post[k][c], opost[k][c], rtol, max(1.0, abs(opost[k][c])) * (1.0 if k == "logp" else ascale))
                assert e <= 1.0, f"{where}: state {k}: {post[k][c]!r} vs {opost[k][c]!r}"
            for k in STATE_VECTORS:
                e = _close(post[k][c], opost[k][c], rtol, np.maximum(np.abs(opost[k][c]), 1e-300) if k in ("stds", "inv_stds")
                           else np.maximum(1.0, np.abs(opost[k][c])))
                assert e <= 1.0, f"{where}: state vector {k} off by {e * rtol:.3e}"
            for k, of in STATE_VARIANCES:
                # sum of squared deviations of `of`: scale by the squared magnitude of the samples that went in
                mag = np.maximum(1.0, np.abs(opost[of][c])) ** 2 * max(1.0, float(opost["foreground_count"][c]))
                e = _close(post[k][c], opost[k][c], rtol, np.maximum(mag, np.abs(opost[k][c])))
                assert e <= 1.0, f"{where}: state vector {k} off by {e * rtol:.3e}"
            # logdet = sum ln(inv_std): reduction order differs, scale by the sum of the magnitudes of the terms
            ld_scale = max(1.0, float(np.sum(np.abs(np.log(opost["inv_stds"][c])))))
            for k in ("point_logdet", "mass_matrix_logdet"):
                e = _close(post[k][c], opost[k][c], rtol, ld_scale)
                assert e <= 1.0, f"{where}: state {k}: {post[k][c]!r} vs {opost[k][c]!r}"
            if post["point_transform_id"][c] == post["mass_matrix_id"][c]:  # else both sides re-whiten at the next draw
                for k in ("transformed_position", "transformed_gradient"):
                    e = _close(post[k][c], opost[k][c], rtol, np.maximum(1.0, np.abs(opost[k][c])))
                    assert e <= 1.0, f"{where}: state vector {k} off by {e * rtol:.3e}"
        state = post
    S.close()
    math.close()
    out = {k: np.stack(v) for k, v in all_stats.items()}
    out["_worst_position_error"] = worst * rtol
    out["_knife_edge"] = knife_edge
    return out


def _schedule_checks(stats, num_tune):
    assert stats["tuning"][:num_tune].all() and not stats["tuning"][num_tune:].any()
    assert stats["n_steps"].min() >= 1


def test_config2_target_whole_warmup_schedule(L, orc):
    """BASELINE config 2 target (1000-dim diagonal Gaussian, sigma = exp(lin(-1, 1))), default settings (num_tune 400, maxdepth 10):
    all 400 tuning draws + 40 sampling draws, 4 chains, on the 64x16 exact tile the bench runs."""
    d, N, tune = 1000, 4, 400
    s = L.DiagNutsSettings(num_tune=tune)
    stats = teacher_forced(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, s, tune + 40, 42, dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, d))),
                           max_knife_edge=2)
    _schedule_checks(stats, tune)
    assert abs(stats["mean_tree_accept"][tune:].mean() - 0.8) < 0.1


def test_config4_target_large_dim_engine_whole_schedule(L, orc):
    """BASELINE config 4 target (10^4-dim Gaussian, sigma = 10^lin(-3, 3): condition number 10^12 before adaptation) on the
    large-dim engine, default settings: all 400 tuning draws + 20 sampling draws, 3 chains."""
    d, N, tune = 10000, 3, 400
    s = L.DiagNutsSettings(num_tune=tune)
    stats = teacher_forced(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, N, d, s, tune + 20, 7, dict(mu=0.0, sigma=10.0 ** np.linspace(-3, 3, d)),
                           max_knife_edge=2)
    _schedule_checks(stats, tune)


def test_funnel_whole_warmup_schedule(L, orc):
    """BASELINE config 3 target (Neal's funnel, d = 10): divergences, depth spread 1..10, 16 chains x (400 + 60) draws."""
    tune = 400
    s = L.DiagNutsSettings(num_tune=tune)
    stats = teacher_forced(L, orc, _abi.NUTS_LOGP_FUNNEL, 16, 10, s, tune + 60, 5, dict(funnel_scale=3.0), rtol=1e-8, max_knife_edge=8)
    _schedule_checks(stats, tune)
    assert stats["diverging"].sum() > 0 and stats["depth"].max() - stats["depth"].min() >= 5


def test_rank1_whole_warmup_schedule(L, orc):
    """BASELINE config 5 target (100-dim Gaussian, Sigma = I + 0.5 * 11^T, tests/sample_normal.rs:29-96): 8 chains x (400 + 40) draws."""
    tune = 400
    s = L.DiagNutsSettings(num_tune=tune)
    stats = teacher_forced(L, orc, _abi.NUTS_LOGP_GAUSS_RANK1, 8, 100, s, tune + 40, 3, dict(mu=0.0, rank1_scale=0.5), max_knife_edge=2)
    _schedule_checks(stats, tune)


def test_draw_variance_estimator_and_short_windows(L, orc):
    """The non-default estimator (update_diag_draw, src/transform/diagonal.rs:85-105) and a schedule with custom window settings."""
    s = L.DiagNutsSettings(num_tune=120, maxdepth=7)
    s.adapt_options.mass_matrix_options.use_grad_based_estimate = 0
    s.adapt_options.mass_matrix_switch_freq = 20
    s.adapt_options.early_mass_matrix_switch_freq = 5
    s.adapt_options.mass_matrix_window_growth = 1.2
    d = 33
    stats = teacher_forced(L, orc, _abi.NUTS_LOGP_GAUSS_DIAG, 6, d, s, 150, 11, dict(mu=0.5, sigma=np.exp(np.linspace(-1, 1, d))))
    _schedule_checks(stats, 120)


def test_checkpoint_resume_is_bit_identical(L):
    """nuts_sampler_set_chain_state: a fresh sampler restored from a checkpoint continues exactly like the original."""
    N, d = 24, 100
    sigma = np.exp(np.linspace(-1, 1, d))
    x0 = np.random.default_rng(2).normal(size=(N, d))
    s = L.DiagNutsSettings(num_tune=60, maxdepth=7)
    m1 = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=sigma)
    a = L.Sampler(m1, s, seed=9, chain_id_offset=5)
    a.set_position(x0)
    a.draw(25)  # checkpoint in the middle of the warm-up, after a window switch
    ckpt = a.chain_state()
    da, sa = a.draw(70)
    m2 = L.CudaMath(N, d, _abi.NUTS_LOGP_GAUSS_DIAG, mu=0.5, sigma=sigma)
    b = L.Sampler(m2, s, seed=9, chain_id_offset=5)
    b.set_chain_state(ckpt)
    assert b.counters()[1] == 25
    db, sb = b.draw(70)
    assert np.array_equal(da, db)
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    fa, fb = a.chain_state(), b.chain_state()
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k
    for x in (a, b):
        x.close()
    for x in (m1, m2):
        x.close()
