"""Low-rank mass-matrix adaptation around the SM_LOWRANK engines: the estimator and the window schedule of the reference on the host,
the leapfrogs on the GPU.

What is mirrored (reference `pymc-devs/nuts-rs`):
  * `LowRankMassMatrixStrategy` (src/transform/adapt/low_rank.rs:14-131, 289-349): a window of draws and gradients per chain (deque +
    `background_split`), `compute_update` = diagonal rescaling -> thin SVDs of draws and gradients -> pivoted QR of the joint subspace ->
    SPD geometric mean of the two regularised covariances (three symmetric eigendecompositions) -> eigenvalue filter -> back-projection.
  * the mass-matrix half of `GlobalStrategy::adapt` (src/adapt_strategy.rs:121-222): early / growing windows, `switch`, an update at every
    switch and every `mass_matrix_update_freq` draws (20 for LowRankNutsSettings, src/sampler.rs:636-642), per chain.
The dense linear algebra (faer in the reference) is numpy / scipy here: once per update and chain, O(dim * window^2) - not on the
north-star path, which is the leapfrog.  The device keeps adapting the step size (dual averaging) on its own; its diagonal mass-matrix
adaptation is switched off (update and switch frequencies beyond num_tune), so its `is_late` flag is always set and the step size
always learns from the symmetric acceptance statistic (the reference does that only in the last window: documented deviation).
The launch length follows the schedule: a launch ends exactly at the next draw after which some chain's update is due, so every chain
sees its new transformation at the same draw index as in the reference.
"""
import collections
import concurrent.futures
import math
import os

import numpy as np
import scipy.linalg

from . import lib as _lib


def _blas_single_threaded():
    try:
        from threadpoolctl import threadpool_limits

        return threadpool_limits(limits=1)
    except ImportError:  # without threadpoolctl the estimator still works, only slower
        import contextlib

        return contextlib.nullcontext()


# ------------------------------------------------------------------------------------------------ the estimator (host)
def rescale_points(draws, grads):
    """adapt/low_rank.rs:150-208.  draws, grads: [dim, n].  Returns (stds, mu, draw_mean, grad_mean, draws', grads') with the rescaled
    and centred copies: x -> (x - mu) / sigma, alpha -> alpha * sigma, sigma = (var x / var alpha)^(1/4), mu = mean x + sigma^2 mean alpha."""
    n = draws.shape[1]
    dmean = draws.sum(axis=1) / n
    gmean = grads.sum(axis=1) / n
    dvar = ((draws - dmean[:, None]) ** 2).sum(axis=1) / n
    gvar = ((grads - gmean[:, None]) ** 2).sum(axis=1) / n
    with np.errstate(divide="ignore", invalid="ignore"):
        sigma = np.sqrt(np.sqrt(dvar / gvar))
        mu = dmean + sigma * sigma * gmean
        d2 = (draws - mu[:, None]) * (1.0 / sigma)[:, None]
        g2 = grads * sigma[:, None]
    dm2 = d2.sum(axis=1) / n
    gm2 = g2.sum(axis=1) / n
    return sigma, mu, dm2, gm2, d2 - dm2[:, None], g2 - gm2[:, None]


def _sym_fun(a, fun):
    w, u = np.linalg.eigh(a)
    return (u * fun(w)) @ u.T


def spd_mean(cov_draws, cov_grads):
    """adapt/low_rank.rs:241-268: B^(-1/2) (B^(1/2) A B^(1/2))^(1/2) B^(-1/2) with A = cov_draws, B = cov_grads (the geometric mean
    of A and B^-1)."""
    w, u = np.linalg.eigh(cov_grads)
    b_sqrt = (u * np.sqrt(w)) @ u.T
    m = b_sqrt @ cov_draws @ b_sqrt
    m_sqrt = _sym_fun(m, np.sqrt)
    b_inv_sqrt = (u * (1.0 / np.sqrt(w))) @ u.T
    return b_inv_sqrt @ m_sqrt @ b_inv_sqrt


def estimate_mass_matrix(draws, grads, gamma):
    """adapt/low_rank.rs:210-239.  draws, grads: [k, n] (projected).  Returns (vals ascending, vecs [k, k]) or None."""
    cov_draws = draws @ draws.T / gamma + np.eye(draws.shape[0])
    cov_grads = grads @ grads.T / gamma + np.eye(grads.shape[0])
    try:
        mean = spd_mean(cov_draws, cov_grads)
        if not np.isfinite(mean).all():
            return None
        vals, vecs = np.linalg.eigh((mean + mean.T) / 2.0)
    except np.linalg.LinAlgError:
        return None
    return vals, vecs


def compute_update(draws, grads, gamma=1e-5, eigval_cutoff=2.0):
    """LowRankMassMatrixStrategy::compute_update (adapt/low_rank.rs:73-131).  draws, grads: [n, dim] (the window, oldest first).
    Returns (stds [dim], mean [dim], vals [r], vecs [r, dim], mean_low_rank [dim]) or None when a factorisation fails."""
    d_, g_ = np.ascontiguousarray(draws.T, dtype=np.float64), np.ascontiguousarray(grads.T, dtype=np.float64)
    stds, mean, draw_mean, grad_mean, d_, g_ = rescale_points(d_, g_)
    if not (np.isfinite(d_).all() and np.isfinite(g_).all()):
        return None  # (the reference's SVD fails on non-finite input: compute_update returns None, nothing changes)
    try:
        ud = np.linalg.svd(d_, full_matrices=False)[0]
        ug = np.linalg.svd(g_, full_matrices=False)[0]
        subspace = np.concatenate([ud, ug], axis=1)
        basis = scipy.linalg.qr(subspace, mode="economic", pivoting=True)[0]  # col_piv_qr().compute_thin_Q()
    except (np.linalg.LinAlgError, ValueError):
        return None
    est = estimate_mass_matrix(basis.T @ d_, basis.T @ g_, gamma)
    if est is None:
        return None
    vals, vecs = est
    keep = (vals > eigval_cutoff) | (vals < 1.0 / eigval_cutoff)
    vals = vals[keep]
    vecs = basis @ vecs[:, keep]  # [dim, r]
    b = vecs @ ((vals - 1.0) * (vecs.T @ grad_mean))
    mu = draw_mean + grad_mean + b
    return stds, mean, vals, np.ascontiguousarray(vecs.T), mu


# ------------------------------------------------------------------------------------------------ the schedule (host) + the sampler
class _ChainWindow:
    """LowRankMassMatrixStrategy of one chain (adapt/low_rank.rs:14-71, 289-349) + the bookkeeping GlobalStrategy keeps per chain."""

    def __init__(self, switch_freq):
        self.draws = collections.deque()
        self.grads = collections.deque()
        self.background_split = 0
        self.last_update = 0
        self.current_window_size = switch_freq

    def add(self, draw, grad):
        self.draws.append(draw)
        self.grads.append(grad)

    def switch(self):  # :318-326
        for _ in range(self.background_split):
            self.draws.popleft()
            self.grads.popleft()
        self.background_split = len(self.draws)

    def background_count(self):
        return len(self.draws) - self.background_split


def schedule_step(w, t, good, draw, grad, early_end, final_window, early_switch_freq, growth, update_freq):
    """One draw of the mass-matrix half of GlobalStrategy::adapt (src/adapt_strategy.rs:139-203) for the window `w` of one chain: feed the
    estimator (`good` = DrawGradCollector::is_good), switch windows when the background is full, and say whether an update of the
    transformation is due after draw index `t` (forced by a switch, or every `update_freq` draws; needs three draws in the window)."""
    is_early = t < early_end
    if (not is_early) and t == early_end:
        w.current_window_size = max(w.current_window_size, w.background_count())
    switch_freq = early_switch_freq if is_early else w.current_window_size
    if good:
        w.add(draw, grad)
    could_switch = w.background_count() >= switch_freq
    # (f64::round rounds half away from zero; Python's round() rounds half to even)
    nxt = early_switch_freq if is_early else max(w.current_window_size + 1, int(math.floor(w.current_window_size * growth + 0.5)))
    is_late = nxt + t > final_window
    force = False
    if could_switch and not is_late:
        w.switch()
        force = True
        if not is_early:
            w.current_window_size = nxt
    if force or (t - w.last_update >= update_freq):
        if len(w.draws) >= 3:  # LowRankMassMatrixStrategy::adapt (adapt/low_rank.rs:340-348)
            w.last_update = t
            return True
    return False


class LowRankSampler:
    """`LowRankNutsSettings` chains on one GPU: lib.Sampler on a low-rank engine + the host-side adaptation above.

    settings: lib.DiagNutsSettings (num_tune, maxdepth, step size options ... as for the diagonal sampler; the mass-matrix window
    options early_window / step_size_window / mass_matrix_switch_freq / early_mass_matrix_switch_freq / mass_matrix_window_growth are
    read here and frozen on the device).  gamma, eigval_cutoff: LowRankSettings (src/transform/low_rank.rs:199-209)."""

    def __init__(self, math, settings, seed, chain_id_offset=0, rank_max=16, gamma=1e-5, eigval_cutoff=2.0, update_freq=20):
        self.math, self.N, self.d = math, math.nchains, math.dim
        self.gamma, self.cutoff, self.update_freq, self.rank_max = gamma, eigval_cutoff, update_freq, rank_max
        ao = settings.adapt_options
        self.num_tune = int(settings.num_tune)
        self.early_end = int(ao.early_window * self.num_tune)  # adapt_strategy.rs:77-98
        step_size_window = int(ao.step_size_window * self.num_tune)
        self.final_window = self.num_tune - step_size_window if self.num_tune >= step_size_window else 0
        self.switch_freq, self.early_switch_freq = int(ao.mass_matrix_switch_freq), int(ao.early_mass_matrix_switch_freq)
        self.growth = float(ao.mass_matrix_window_growth)
        dev = type(settings).from_buffer_copy(settings)  # the device only adapts the step size
        big = 1 << 40
        dev.adapt_options.mass_matrix_update_freq = big
        dev.adapt_options.early_mass_matrix_switch_freq = big
        dev.adapt_options.mass_matrix_switch_freq = big
        self.sampler = _lib.Sampler(math, dev, seed, chain_id_offset, lowrank_rank_max=rank_max)
        self.windows = [_ChainWindow(self.switch_freq) for _ in range(self.N)]
        self.draw_index = 0
        self.updates = 0          # transformations installed so far (summed over chains)
        self.last_ranks = np.zeros(self.N, dtype=np.int64)
        self._gbuf = None
        nthreads = min(self.N, os.cpu_count() or 1, int(os.environ.get("NUTS_B200_ESTIMATOR_THREADS", "16")))
        self._pool = concurrent.futures.ThreadPoolExecutor(max_workers=nthreads) if nthreads > 1 else None

    def set_position(self, position):
        status = self.sampler.set_position(position)
        st = self.sampler.chain_state()
        for c in range(self.N):  # LowRankMassMatrixStrategy::init: add_draw(initial point) (adapt/low_rank.rs:289-305)
            if status[c] == 0:
                self.windows[c].add(st["position"][c].copy(), st["gradient"][c].copy())
        self.alive = status == 0
        return status

    # draws until (and including) the next draw after which some chain's update can be due
    def _next_launch(self, remaining):
        t = self.draw_index
        if t >= self.final_window:
            return remaining
        best = self.final_window - t
        for c in range(self.N):
            if not self.alive[c]:
                continue
            w = self.windows[c]
            is_early = t < self.early_end
            switch_freq = self.early_switch_freq if is_early else w.current_window_size
            # (a window that is full but may not switch any more - `is_late` - never switches: only the update frequency counts)
            to_switch = max(switch_freq - w.background_count(), 1) if w.background_count() < switch_freq else remaining
            to_update = max(w.last_update + self.update_freq - t + 1, 1)
            best = min(best, to_switch, to_update)
        return max(1, min(best, remaining))

    def _adapt_chain(self, c, t, good, draw, grad):
        """the mass-matrix half of GlobalStrategy::adapt for draw index t of chain c; returns True when an update is due"""
        return schedule_step(self.windows[c], t, good, draw, grad, self.early_end, self.final_window, self.early_switch_freq, self.growth,
                             self.update_freq)

    def draw(self, n_draws):
        """n_draws x Chain::draw for every chain; returns (draws [n, N, dim], stats dict) like lib.Sampler.draw."""
        N, d = self.N, self.d
        out_draws, out_stats = [], []
        remaining = n_draws
        while remaining > 0:
            k = self._next_launch(remaining)
            tuning_launch = self.draw_index < self.final_window
            if tuning_launch:
                if self._gbuf is None or self._gbuf.array.shape[0] < k:
                    if self._gbuf is not None:
                        self.sampler.set_grads_out(None)
                        self._gbuf.close()
                    self._gbuf = _lib.HostBuffer((max(k, 32), N, d))
                self.sampler.set_grads_out(self._gbuf)
            else:
                self.sampler.set_grads_out(None)
            draws, stats = self.sampler.draw(k)
            out_draws.append(draws)
            out_stats.append(stats)
            if tuning_launch:
                grads = self._gbuf.array[:k]
                due = np.zeros(N, dtype=bool)
                for j in range(k):
                    t = self.draw_index + j
                    if t >= self.final_window:
                        break
                    div, idx = stats["diverging"][j], stats["index_in_trajectory"][j]
                    good = np.where(div != 0, np.abs(idx) > 4, idx != 0)  # DrawGradCollector::register_draw (adapt/diagonal.rs:74-83)
                    for c in range(N):
                        if self.alive[c] and np.isfinite(draws[j, c]).all():
                            due[c] |= self._adapt_chain(c, t, bool(good[c]), draws[j, c].copy(), grads[j, c].copy())
                if due.any():
                    self._install(due)
            self.draw_index += k
            remaining -= k
        stats = {name: np.concatenate([s[name] for s in out_stats]) for name in out_stats[0]}
        return np.concatenate(out_draws), stats

    def _install(self, due):
        N, d = self.N, self.d
        cur = self.sampler.chain_state()
        stds, mean = cur["stds"].copy(), cur["mean"].copy()
        vals, vecs = np.ones((N, self.rank_max)), np.zeros((N, self.rank_max, d))
        mu, rank = np.zeros((N, d)), np.zeros(N, dtype=np.int32)
        # the estimators of the chains that are due are independent: one host thread each (LAPACK releases the GIL), like the
        # reference's one-chain-per-rayon-task warm-up (src/sampler.rs:1287-1326)
        todo = [int(c) for c in np.nonzero(due)[0]]
        work = lambda c: compute_update(np.array(self.windows[c].draws), np.array(self.windows[c].grads), self.gamma, self.cutoff)
        # (BLAS / LAPACK themselves run single-threaded meanwhile: on these small matrices - 2 * window <= a few hundred rows - their
        # own threading costs 5x more than it gains, and one thread per chain makes every chain's result independent of the others)
        with _blas_single_threaded():
            if len(todo) > 1 and self._pool is not None:
                updates = dict(zip(todo, self._pool.map(work, todo)))
            else:
                updates = {c: work(c) for c in todo}
        for c in todo:
            upd = updates[c]
            if upd is None:
                stds[c, 0] = np.nan  # "return" of LowRankMassMatrixStrategy::update: nothing changes for this chain
                continue
            s_, m_, va, ve, mu_ = upd
            if len(va) > self.rank_max:  # keep the eigenvalues farthest from 1 (the reference keeps all of them)
                order = np.argsort(-np.abs(np.log(va)))[: self.rank_max]
                va, ve = va[order], ve[order]
            r = len(va)
            stds[c], mean[c], mu[c], rank[c] = s_, m_, mu_, r
            vals[c, :r], vecs[c, :r] = va, ve
        # chains that are not due keep what they have: mask them with a non-finite value so the update is rejected for them
        for c in np.nonzero(~due)[0]:
            stds[c, 0] = np.nan
        ok = self.sampler.set_lowrank_transform(stds, mean, vals, vecs, mu, rank)
        self.updates += int(ok.sum())
        self.last_ranks = np.where(ok, rank, self.last_ranks)

    def close(self):
        if self._gbuf is not None:
            self.sampler.set_grads_out(None)
            self._gbuf.close()
            self._gbuf = None
        if self._pool is not None:
            self._pool.shutdown(wait=True)
            self._pool = None
        self.sampler.close()
