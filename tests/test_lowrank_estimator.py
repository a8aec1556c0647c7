"""The host-side low-rank estimator (nuts_rs_b200/lowrank.py) against the reference's known-answer tests
(src/transform/adapt/low_rank.rs:354-407) and its defining properties."""
import numpy as np

from nuts_rs_b200 import lowrank


def test_spd_mean_known_answer():
    # adapt/low_rank.rs:354-381: spd_mean(diag(1, 4, 8), diag(1, 1, 0.5)) = diag(1, 2, 4) to 1e-10
    out = lowrank.spd_mean(np.diag([1.0, 4.0, 8.0]), np.diag([1.0, 1.0, 0.5]))
    np.testing.assert_allclose(out, np.diag([1.0, 2.0, 4.0]), rtol=1e-10, atol=1e-10)


def test_spd_mean_is_the_geometric_mean():
    # X = mean(A, B) solves X B X = A (the geometric mean of A and B^-1); symmetric positive definite
    rng = np.random.default_rng(0)
    a, b = (lambda m: m @ m.T + np.eye(6))(rng.normal(size=(6, 6))), (lambda m: m @ m.T + np.eye(6))(rng.normal(size=(6, 6)))
    x = lowrank.spd_mean(a, b)
    np.testing.assert_allclose(x @ b @ x, a, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(x, x.T, atol=1e-10)
    assert (np.linalg.eigvalsh(x) > 0).all()


def test_estimate_mass_matrix_known_answer():
    # adapt/low_rank.rs:383-407: grads = -draws (a standard normal target) => all eigenvalues 1 to 1e-5
    rng = np.random.default_rng(1)
    draws = rng.normal(size=(20, 3))
    vals, vecs = lowrank.estimate_mass_matrix(draws, -draws, 0.0001)
    assert (vals > 0).all() and np.isfinite(vecs).all()
    np.testing.assert_allclose(vals, np.ones(20), rtol=1e-5, atol=1e-5)


def test_compute_update_recovers_a_correlated_gaussian_exactly():
    """The property behind tests/sample_normal.rs:320-356: for draws of N(m, Sigma) with exact scores alpha = -Sigma^-1 (x - m), once the
    window spans the space the transformation F(z) = sigma * ((I + U (sqrt(lambda) - 1) U^T) z + mu_lr) + mean whitens the target
    exactly: z + grad_z = 0 for every point (fisher distance 0)."""
    rng = np.random.default_rng(2)
    d, n = 10, 40
    cov = np.eye(d) + 0.5 * np.ones((d, d))
    m = rng.normal(size=d)
    prec = np.linalg.inv(cov)
    x = m + rng.multivariate_normal(np.zeros(d), cov, size=n)
    g = -(x - m) @ prec
    stds, mean, vals, vecs, mu = lowrank.compute_update(x, g, gamma=1e-5, eigval_cutoff=1.00001)
    assert np.isfinite(vals).all() and len(vals) >= 1
    a_fwd = np.eye(d) + vecs.T @ np.diag(np.sqrt(vals) - 1.0) @ vecs      # I + U (sqrt(lambda) - 1) U^T
    a_inv = np.eye(d) + vecs.T @ np.diag(1.0 / np.sqrt(vals) - 1.0) @ vecs
    for k in range(5):
        xt = m + rng.multivariate_normal(np.zeros(d), cov)
        gx = -prec @ (xt - m)
        z = a_inv @ ((xt - mean) / stds - mu)
        gz = a_fwd @ (gx * stds)
        assert np.sum((z + gz) ** 2) < 1e-10
    # eigenvectors are orthonormal and the filter dropped the eigenvalues next to 1
    np.testing.assert_allclose(vecs @ vecs.T, np.eye(len(vals)), atol=1e-8)
    assert ((vals > 1.00001) | (vals < 1 / 1.00001)).all()


def test_rescale_points():
    # adapt/low_rank.rs:150-208: sigma = (var x / var alpha)^(1/4), mu = mean x + sigma^2 mean alpha; outputs centred
    rng = np.random.default_rng(3)
    x, g = rng.normal(size=(4, 30)) * np.array([[1.0], [2.0], [0.5], [3.0]]) + 1.0, rng.normal(size=(4, 30))
    sigma, mu, dm, gm, x2, g2 = lowrank.rescale_points(x.copy(), g.copy())
    np.testing.assert_allclose(sigma, (x.var(axis=1) / g.var(axis=1)) ** 0.25, rtol=1e-12)
    np.testing.assert_allclose(mu, x.mean(axis=1) + sigma**2 * g.mean(axis=1), rtol=1e-12)
    np.testing.assert_allclose(x2.mean(axis=1), 0.0, atol=1e-12)
    np.testing.assert_allclose(g2.mean(axis=1), 0.0, atol=1e-12)
    np.testing.assert_allclose(x2 + dm[:, None], (x - mu[:, None]) / sigma[:, None], rtol=1e-12, atol=1e-12)


def test_window_schedule_matches_the_global_strategy(orc):
    """lowrank.schedule_step (the mass-matrix half of GlobalStrategy::adapt, src/adapt_strategy.rs:139-203, around the deque of
    LowRankMassMatrixStrategy) against the oracle's GlobalStrategy driving the diagonal strategy with the low-rank update frequency
    (20): both keep a foreground window that contains the background window, so after every draw the counts, the window size, the
    draw of the last update and whether the transformation changed must agree - on the funnel, where divergent draws are skipped."""
    from nuts_rs_b200 import _abi

    s = _abi.default_settings()
    s.num_tune, s.maxdepth = 400, 6
    s.adapt_options.mass_matrix_update_freq = 20
    ao = s.adapt_options
    early_end = int(ao.early_window * s.num_tune)
    final_window = s.num_tune - int(ao.step_size_window * s.num_tune)
    N, d = 4, 8
    m = orc.Model(_abi.NUTS_LOGP_FUNNEL, d, funnel_scale=3.0)
    smp = orc.Sampler(m, s, seed=3, nchains=N, nthreads=2)
    x0 = np.random.default_rng(0).normal(size=(N, d))
    x0[:, 0] = 0.1
    assert (smp.set_position(x0) == 0).all()
    windows = [lowrank._ChainWindow(int(ao.mass_matrix_switch_freq)) for _ in range(N)]
    for w in windows:
        w.add(None, None)  # LowRankMassMatrixStrategy::init adds the initial point
    prev = smp.chain_state()
    skipped = 0
    for t in range(final_window + 5):
        _, stats = smp.draw(1)
        cur = smp.chain_state()
        div, idx = stats["diverging"][0], stats["index_in_trajectory"][0]
        good = np.where(div != 0, np.abs(idx) > 4, idx != 0)
        skipped += int((~good).sum())
        for c in range(N):
            if t < final_window:
                due = lowrank.schedule_step(windows[c], t, bool(good[c]), None, None, early_end, final_window,
                                            int(ao.early_mass_matrix_switch_freq), float(ao.mass_matrix_window_growth), 20)
            else:
                due = False
            w = windows[c]
            assert due == (cur["mass_matrix_id"][c] != prev["mass_matrix_id"][c]), (t, c)
            assert w.background_count() == cur["background_count"][c] and len(w.draws) == cur["foreground_count"][c], (t, c)
            assert w.current_window_size == cur["current_window_size"][c] and w.last_update == cur["last_update"][c], (t, c)
        prev = cur
    assert skipped > 0  # the funnel produced draws that the estimators skip
